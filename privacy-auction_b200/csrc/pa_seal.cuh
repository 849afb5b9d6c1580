// Whole-auction runner for the SEAL protocol: every bidder of every auction of a
// batch advances in lock step, one batched kernel sequence per protocol phase,
// all state resident in HBM.  This is what the reference's main loop
// (SEAL/main.cpp:32-120) does one bidder, one EC_POINT_mul at a time.
//
//   commit    : Bidder::commitBid for all (bidder, bit)            SEAL/bidder.cpp:1109-1162
//   per step  : roundOne -> Y scan -> roundTwo -> roundThree        :1203-1236, 1271-1336, 1386-1421
//   verify    : every published proof is verified exactly once     :1171-1195, 1245-1262, 1346-1377
//               (the reference lets each of the n bidders repeat the same
//               deterministic checks; the verdicts are identical)
//
// Two schedules publish the same bytes: step-major (a batch of auctions in lock step, one kernel
// sequence per protocol step on four streams) and, for ONE auction, phase-major (keys, Y and both
// cryptogram candidates of every step in large launches, one thread block walks the steps, then
// all proofs of all steps in one batch per kind) - see "single-auction schedule" below.
//
// Partitioning: independent auctions need no exchange at
// all; ONE auction can be sharded by bidder slice, in which case published points
// (step-major: X_i and b_i once per step; phase-major: the X_i of all steps at once,
// then each rank's partial sum of cryptograms per step) are all-gathered through a
// caller-supplied callback (NCCL over NVLink in bench.py / tests, see
// INTEGRATION.md) into the device buffers d_recv.
#pragma once

// ---- runner kernels -----------------------------------------------------------------------
// commitment points straight from the draw array: slot s has draws
//   alpha, beta, v_A, v_B, r1, d1, d2   (7 x 32 B)
__global__ void __launch_bounds__(PA_BLOCK)
k_seal_commit_points(const unsigned char *rndc, const unsigned char *bits, const u32 *__restrict__ comb, u32 *jout, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  int which = t / n, i = t % n;
  sc a, b, k;
  ld_sc(a, rndc + 224 * (size_t)i);
  ld_sc(b, rndc + 224 * (size_t)i + 32);
  if (which == 0) {
    sc bit;
    sc_set_zero(bit);
    bit.v[0] = bits[i] ? 1u : 0u;
    sc_mul(k, a, b);
    sc_add(k, k, bit);
  } else {
    k = which == 1 ? a : b;
  }
  jac r;
  fixed_base_mul(r, k, comb);
  st_jac(jout + 24 * ((size_t)i * 3 + which), r);
}

// X = g^x, R = g^r from the round-one draws x, r, v_X, v_R (4 x 32 B per bidder)
__global__ void __launch_bounds__(PA_BLOCK)
k_seal_r1_points(const unsigned char *rnd1, const u32 *__restrict__ comb, u32 *jout, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  int which = t / n, i = t % n;
  sc k;
  ld_sc(k, rnd1 + 128 * (size_t)i + 32 * which);
  jac r;
  fixed_base_mul(r, k, comb);
  st_jac(jout + 24 * ((size_t)i * 2 + which), r);
}

// the cryptogram: b = R^x if the bidder vetoes, Y^x otherwise           SEAL/bidder.cpp:1301-1309
__global__ void __launch_bounds__(PA_BLOCK)
k_seal_encode(const u32 *act, const u32 *pauc, const unsigned char *bits, const u32 *boff, int step,
              const unsigned char *junc, const unsigned char *prevbit, const unsigned char *r1, const unsigned char *Y,
              const unsigned char *rnd1, unsigned char *ebit, u32 *jout, int n) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  u32 slot = act[p];
  int bit = bits[boff[slot] + step];
  int veto = bit && (!junc[pauc[p]] || prevbit[slot]);
  ebit[p] = (unsigned char)veto;
  jac P, r;
  sc x;
  ld_point_jac(P, veto ? r1 + 320 * (size_t)p + 64 : Y + 64 * (size_t)p);
  ld_sc(x, rnd1 + 128 * (size_t)p);
  var_base_mul(r, P, x);
  st_jac(jout + 24 * (size_t)p, r);
}

PA_D void cp64(unsigned char *d, const unsigned char *s) {
  const uint4 *a = reinterpret_cast<const uint4 *>(s);
  uint4 *b = reinterpret_cast<uint4 *>(d);
  b[0] = a[0]; b[1] = a[1]; b[2] = a[2]; b[3] = a[3];
}
PA_D void cp32(unsigned char *d, const unsigned char *s) {
  const uint4 *a = reinterpret_cast<const uint4 *>(s);
  uint4 *b = reinterpret_cast<uint4 *>(d);
  b[0] = a[0]; b[1] = a[1];
}

// statement / witness assembly for the round-two proofs of group members g[q] (positions in
// the active list).  Stage 1: (b, X, Y, R, c, A, B), extended secrets (x, alpha, r, beta).  Stage 2:
// (Bi, Xi, Ri, Bj, Xj, Rj, Ci, A, B, Yi, Yj), (xi, xj, alpha, ri, rj, beta), bi = encoded bit,
// bj = prevDecidingBit, cb = the committed bit (pa_proof.cuh, "prover with witnesses").
__global__ void k_seal_stmt(int stage, const u32 *g, const u32 *act, const u32 *boff, int step, const unsigned char *b,
                            const unsigned char *r1, const unsigned char *Y, const unsigned char *rnd1,
                            const unsigned char *crec, const unsigned char *rndc, const unsigned char *prevpts,
                            const unsigned char *prevx, const unsigned char *ebit, const unsigned char *prevbit,
                            const unsigned char *bits, unsigned char *stmt, unsigned char *sec, unsigned char *bi,
                            unsigned char *bj, unsigned char *cb, int n) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  u32 p = g[q], slot = act[p];
  size_t cs = (size_t)boff[slot] + step;  // commitment slot of this bidder's current bit
  const unsigned char *X = r1 + 320 * (size_t)p, *R = X + 64, *c = crec + 736 * cs;
  if (stage == 1) {
    unsigned char *o = stmt + 448 * (size_t)q;
    cp64(o, b + 64 * (size_t)p); cp64(o + 64, X); cp64(o + 128, Y + 64 * (size_t)p); cp64(o + 192, R);
    cp64(o + 256, c); cp64(o + 320, c + 64); cp64(o + 384, c + 128);
    unsigned char *w = sec + 128 * (size_t)q;
    cp32(w, rnd1 + 128 * (size_t)p); cp32(w + 32, rndc + 224 * cs);
    cp32(w + 64, rnd1 + 128 * (size_t)p + 32); cp32(w + 96, rndc + 224 * cs + 32);
    bi[q] = ebit[p];
  } else {
    const unsigned char *pp = prevpts + 256 * (size_t)slot;  // X, R, Y, b at the previous deciding step
    unsigned char *o = stmt + 704 * (size_t)q;
    cp64(o, b + 64 * (size_t)p); cp64(o + 64, X); cp64(o + 128, R);
    cp64(o + 192, pp + 192); cp64(o + 256, pp); cp64(o + 320, pp + 64);
    cp64(o + 384, c); cp64(o + 448, c + 64); cp64(o + 512, c + 128);
    cp64(o + 576, Y + 64 * (size_t)p); cp64(o + 640, pp + 128);
    unsigned char *w = sec + 192 * (size_t)q;
    cp32(w, rnd1 + 128 * (size_t)p); cp32(w + 32, prevx + 64 * (size_t)slot); cp32(w + 64, rndc + 224 * cs);
    cp32(w + 96, rnd1 + 128 * (size_t)p + 32); cp32(w + 128, prevx + 64 * (size_t)slot + 32); cp32(w + 160, rndc + 224 * cs + 32);
    bi[q] = ebit[p];
    bj[q] = prevbit[slot];
    cb[q] = bits[cs];
  }
}

// after round three: in a deciding step every bidder snapshots (X, R, Y, b, x) and folds its
// true bit into prevDecidingBit                                         SEAL/bidder.cpp:1397-1411
__global__ void k_seal_update(const u32 *act, const u32 *pseg, const u32 *pauc, const int *isinf, const unsigned char *bits,
                              const u32 *boff, int step, const unsigned char *r1, const unsigned char *Y,
                              const unsigned char *b, const unsigned char *rnd1, unsigned char *prevpts,
                              unsigned char *prevx, unsigned char *prevbit, unsigned char *junc, int n) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (isinf[pseg[p]]) return;
  u32 slot = act[p];
  unsigned char *pp = prevpts + 256 * (size_t)slot;
  cp64(pp, r1 + 320 * (size_t)p); cp64(pp + 64, r1 + 320 * (size_t)p + 64);
  cp64(pp + 128, Y + 64 * (size_t)p); cp64(pp + 192, b + 64 * (size_t)p);
  cp32(prevx + 64 * (size_t)slot, rnd1 + 128 * (size_t)p);  // x and r of the deciding step
  cp32(prevx + 64 * (size_t)slot + 32, rnd1 + 128 * (size_t)p + 32);
  prevbit[slot] &= bits[boff[slot] + step];
  junc[pauc[p]] = 1;
}

// o[i] = pair[2i] & pair[2i+1] (& c[i])
__global__ void k_seal_and_pairs(const unsigned char *pair, const unsigned char *c, unsigned char *o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = pair[2 * i] & pair[2 * i + 1] & (c ? c[i] : 1);
}
__global__ void k_seal_scatter_u8(const u32 *g, const unsigned char *src, unsigned char *dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[g[i]] = src[i];
}

// ---- single-auction schedule ("phase-major") ------------------------------------------------------
// One auction advanced step by step is a chain of lone-warp latencies: keys -> Y scan -> cryptogram ->
// round three cost ~2 ms per step whatever n is.  But the keys of ALL steps depend only on the draw
// counters, the Y of all steps only on the keys, and the cryptogram of a bidder is one of two values,
// Y^x or R^x, whichever the auction state selects.  So for one auction the runner computes keys, Y and
// BOTH candidates for every step in three large launches, lets one thread block walk through the
// steps (select, sum, is-infinity, update the state) without the host, and then proves and verifies
// the proofs of all steps in one batch per kind.  Item i = step * m + bidder.

// both candidates of every item: cand[2i] = Y^x (no veto), cand[2i+1] = R^x (veto)    SEAL/bidder.cpp:1301-1309
// The veto candidate is the bidder's own R = g^r raised to x, i.e. g^(r x): a fixed-base multiplication.
__global__ void __launch_bounds__(PA_BLOCK, PA_VAR_MINBLOCKS)
k_seal_candidates(const unsigned char *Y, const unsigned char *rnd1, const u32 *__restrict__ comb, u32 *jout, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  int which = t / n, i = t % n;
  jac r;
  sc x;
  ld_sc(x, rnd1 + 128 * (size_t)i);
  if (which) {
    sc rr, k;
    ld_sc(rr, rnd1 + 128 * (size_t)i + 32);
    sc_mul(k, rr, x);
    fixed_base_mul(r, k, comb);
  } else {
    jac P;
    ld_point_jac(P, Y + 64 * (size_t)i);
    var_base_mul(r, P, x);
  }
  st_jac(jout + 24 * ((size_t)i * 2 + which), r);
}

// statements and witnesses of the round-two proofs of every item; stage-1 items are the first n1
__global__ void k_seal_stmt_items(int m, int n1, const int *stage, const int *prevstep, const u32 *boff, const unsigned char *b,
                                  const unsigned char *r1, const unsigned char *Y, const unsigned char *rnd1,
                                  const unsigned char *crec, const unsigned char *rndc, const unsigned char *ebit,
                                  const unsigned char *bjv, const unsigned char *bits, unsigned char *stmt1, unsigned char *sec1,
                                  unsigned char *bi1, unsigned char *stmt2, unsigned char *sec2, unsigned char *bi2, unsigned char *bj2,
                                  unsigned char *cb2, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = i / m, p = i % m;
  size_t cs = (size_t)boff[p] + s;
  const unsigned char *X = r1 + 320 * (size_t)i, *R = X + 64, *c = crec + 736 * cs;
  if (stage[s] == 1) {
    size_t q = i;
    unsigned char *o = stmt1 + 448 * q;
    cp64(o, b + 64 * (size_t)i); cp64(o + 64, X); cp64(o + 128, Y + 64 * (size_t)i); cp64(o + 192, R);
    cp64(o + 256, c); cp64(o + 320, c + 64); cp64(o + 384, c + 128);
    unsigned char *w = sec1 + 128 * q;  // x, alpha, r, beta
    cp32(w, rnd1 + 128 * (size_t)i); cp32(w + 32, rndc + 224 * cs);
    cp32(w + 64, rnd1 + 128 * (size_t)i + 32); cp32(w + 96, rndc + 224 * cs + 32);
    bi1[q] = ebit[i];
  } else {
    size_t q = (size_t)i - n1, j = (size_t)prevstep[s] * m + p;  // the same bidder at the previous deciding step
    unsigned char *o = stmt2 + 704 * q;
    cp64(o, b + 64 * (size_t)i); cp64(o + 64, X); cp64(o + 128, R);
    cp64(o + 192, b + 64 * j); cp64(o + 256, r1 + 320 * j); cp64(o + 320, r1 + 320 * j + 64);
    cp64(o + 384, c); cp64(o + 448, c + 64); cp64(o + 512, c + 128);
    cp64(o + 576, Y + 64 * (size_t)i); cp64(o + 640, Y + 64 * j);
    unsigned char *w = sec2 + 192 * q;  // xi, xj, alpha, ri, rj, beta
    cp32(w, rnd1 + 128 * (size_t)i); cp32(w + 32, rnd1 + 128 * j); cp32(w + 64, rndc + 224 * cs);
    cp32(w + 96, rnd1 + 128 * (size_t)i + 32); cp32(w + 128, rnd1 + 128 * j + 32); cp32(w + 160, rndc + 224 * cs + 32);
    bi2[q] = ebit[i];
    bj2[q] = bjv[i];
    cb2[q] = bits[cs];
  }
}

// ---- the same schedule with the auction sharded by bidder slice ---------------------------------------
// dst[i] = src[idx[i]], 64-byte items
__global__ void k_seal_gather64(unsigned char *dst, const unsigned char *src, const u32 *idx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cp64(dst + 64 * (size_t)i, src + 64 * (size_t)idx[i]);
}
// send[(s * slice + p) * 64] = X of local item (s, p), steps [s0, c)
__global__ void k_seal_pack_x(const unsigned char *r1, unsigned char *send, int m, int slice, int s0, int n) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int s = s0 + k / m, p = k % m;
  cp64(send + 64 * ((size_t)s * slice + p), r1 + 320 * ((size_t)s * m + p));
}
// One step of the walk when every rank holds a slice: first fold the ranks' partial sums of step s - 1
// (Jacobian, 128 bytes apart in `recv`) into the auction state, then select this rank's cryptograms of
// step s and leave their sum in `send`.  s == limit only folds.
__global__ void __launch_bounds__(PA_SCAN_T)
k_seal_decide_shard(int s, int m, int limit, int world, const unsigned char *bits, const u32 *boff, const unsigned char *cand,
                    unsigned char *prevbit, int *state, unsigned char *ebit, unsigned char *bj, unsigned char *b, int *stage,
                    int *prevstep, int *r3, const u32 *recv, u32 *send) {
  __shared__ __align__(16) u32 part[PA_SCAN_T][24];
  __shared__ int s_junc, s_last, s_deciding;
  int t = threadIdx.x;
  if (t == 0) s_junc = state[0], s_last = state[1], s_deciding = 0;
  if (s > 0) {  // fold the ranks' partial sums of step s - 1 (tree over the first `world` threads)
    jac a;
    if (t < world) ld_jac(a, recv + 32 * (size_t)t); else jac_set_inf(a);
    st_jac(part[t], a);
    __syncthreads();
    for (int d = PA_SCAN_T / 2; d > 0; d >>= 1) {
      if (t < d && t + d < world) {
        jac c;
        ld_jac(a, part[t]);
        ld_jac(c, part[t + d]);
        jac_add(a, a, c);
        st_jac(part[t], a);
      }
      __syncthreads();
    }
    if (t == 0) {
      ld_jac(a, part[0]);
      s_deciding = jac_is_inf(a) ? 0 : 1;
      r3[s - 1] = s_deciding;
      if (s_deciding) s_junc = 1, s_last = s - 1;
    }
  }
  __syncthreads();
  if (s_deciding)
    for (int p = t; p < m; p += PA_SCAN_T) prevbit[p] &= bits[boff[p] + s - 1];
  if (s < limit) {
    int junc = s_junc;
    jac acc;
    jac_set_inf(acc);
    for (int p = t; p < m; p += PA_SCAN_T) {
      size_t i = (size_t)s * m + p;
      int bit = bits[boff[p] + s];
      int pb = prevbit[p];
      int veto = bit && (!junc || pb);
      ebit[i] = (unsigned char)veto;
      bj[i] = (unsigned char)pb;
      const unsigned char *src = cand + 64 * (2 * i + veto);
      cp64(b + 64 * i, src);
      aff x;
      ld_aff(x, src);
      jac_madd(acc, acc, x);
    }
    st_jac(part[t], acc);
    __syncthreads();
    for (int d = PA_SCAN_T / 2; d > 0; d >>= 1) {
      if (t < d) {
        jac a, c;
        ld_jac(a, part[t]);
        ld_jac(c, part[t + d]);
        jac_add(a, a, c);
        st_jac(part[t], a);
      }
      __syncthreads();
    }
    if (t < 32) send[t] = t < 24 ? part[0][t] : 0u;
    if (t == 0) stage[s] = junc ? 2 : 1, prevstep[s] = s_last;
  }
  if (t == 0) state[0] = s_junc, state[1] = s_last, state[2] = s;
}

// ---- peer exchange: the ranks' kernels talk through each other's HBM ---------------------------------
// Every rank owns a window (pa_xchg_create) that all ranks have mapped (pa_xchg_connect, CUDA IPC over
// NVLink).  A value is PUT into the same slot of every rank's window: 24 payload words, a system-scope
// fence, then a tag word; readers poll the tag in their OWN memory and then read the payload.  No host
// round trip and no collective call: the step walk of a sharded auction is one kernel per pass.
//   slot (kind, parity, step, rank): 128 bytes = 24 payload words + tag at word 24
//     kind 0  a rank's sum of public keys of a step     (Y reconstruction, SEAL/bidder.cpp:1286-1299)
//     kind 1  a rank's sum of cryptograms of a step     (round three, SEAL/bidder.cpp:1393-1397)
//     kind 2  end of run: (draws clean, verdict, max bid)
//   tag = epoch << 8 | pass: epoch counts sharded runs (the same on every rank, parity = epoch & 1 picks one of
//   two slot sets), pass counts the passes of a run, so a slot written twice in a run (a step redone after
//   the junction) is never mistaken for its first value.
//   bulk area: two buffers for a plain all-gather of up to PA_XCHG_BULK bytes (the step-major schedule's X
//   and b), tagged with a sequence number per (buffer, rank).
#define PA_XCHG_STEPS 256  // real steps (<= 64) and the virtual steps of the junction planes
#define PA_XCHG_SLOTS (3 * 2 * PA_XCHG_STEPS * PA_XCHG_MAX_WORLD)
#define PA_XCHG_BULK_OFF ((size_t)PA_XCHG_SLOTS * 128 + 4096)
#define PA_XCHG_BULK ((PA_XCHG_BYTES - PA_XCHG_BULK_OFF) / 2)
#define PA_XCHG_TIMEOUT_NS 10000000000ull  // a peer that does not show up in 10 s is an error, not a hang

PA_D unsigned char *xchg_slot(unsigned char *win, int kind, int par, int step, int rank) {
  return win + 128 * (size_t)((((kind * 2 + par) * PA_XCHG_STEPS) + step) * PA_XCHG_MAX_WORLD + rank);
}
PA_D unsigned long long pa_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// block-cooperative: put val[0..24) into `slot` of every rank's window (warp w serves ranks w, w + nwarps, ...)
PA_D void xchg_put(unsigned char *const *peers, int world, int kind, int par, int step, int rank, const u32 *val, u32 tag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < world; r += nw) {
    volatile u32 *dst = (volatile u32 *)xchg_slot(peers[r], kind, par, step, rank);
    if (lane < 24) dst[lane] = val[lane];
    __threadfence_system();
    __syncwarp();
    if (lane == 0) dst[24] = tag;
  }
}
// block-cooperative: wait until every rank's value for `slot` has arrived in THIS rank's window, copy the
// payloads to shared memory.  Returns false (for the whole block) when a peer timed out; *err is set.
PA_D bool xchg_get(unsigned char *win, int world, int kind, int par, int step, u32 tag, u32 (*dst)[24], int *err) {
  const int t = threadIdx.x;
  int bad = 0;
  if (t < world) {
    volatile u32 *src = (volatile u32 *)xchg_slot(win, kind, par, step, t);
    const unsigned long long t0 = pa_globaltimer();
    while (src[24] != tag) {
      if (pa_globaltimer() - t0 > PA_XCHG_TIMEOUT_NS) {
        bad = 1;
        atomicExch(err, 1);
        break;
      }
      __nanosleep(64);
    }
  }
  if (__syncthreads_or(bad)) return false;
  __threadfence_system();
  for (int k = t; k < world * 24; k += blockDim.x)
    dst[k / 24][k % 24] = ((volatile u32 *)xchg_slot(win, kind, par, step, k / 24))[k % 24];
  __syncthreads();
  return true;
}

// Plain all-gather through the windows: every rank's `bytes` at send go to offset rank * bytes of bulk
// buffer `buf` in every window.  Grid = world blocks: block r copies to rank r, then block 0 waits for all tags.
__global__ void k_xchg_allgather(const unsigned char *send, size_t bytes, unsigned char *const *peers, int world, int rank,
                                 int buf, u32 seq, int *err) {
  const int r = blockIdx.x;
  unsigned char *bulk = peers[r] + PA_XCHG_BULK_OFF + (size_t)buf * PA_XCHG_BULK;
  const uint4 *src = reinterpret_cast<const uint4 *>(send);
  uint4 *dst = reinterpret_cast<uint4 *>(bulk + (size_t)rank * bytes);
  for (size_t k = threadIdx.x; k < bytes / 16; k += blockDim.x) dst[k] = src[k];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *(volatile u32 *)(peers[r] + (size_t)PA_XCHG_SLOTS * 128 + 4 * (buf * PA_XCHG_MAX_WORLD + rank)) = seq;
  if (r != rank) return;
  // the block addressed to ourselves also waits until everybody's part has landed here
  if ((int)threadIdx.x < world) {
    volatile u32 *tagp = (volatile u32 *)(peers[rank] + (size_t)PA_XCHG_SLOTS * 128 + 4 * (buf * PA_XCHG_MAX_WORLD + threadIdx.x));
    const unsigned long long t0 = pa_globaltimer();
    while (*tagp != seq) {
      if (pa_globaltimer() - t0 > PA_XCHG_TIMEOUT_NS) {
        atomicExch(err, 1);
        break;
      }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

// Y reconstruction of steps [s0, s0 + gridDim.x) when every rank holds a slice of the bidders: block b scans
// this rank's m public keys of step s0 + b, PUTS their sum, GETS every rank's sum, and finishes
// Y_id = E_id + P_id - T with E_id = (sum of the ranks before this one) + (local exclusive prefix).
// Nobody needs anybody else's X: 96 bytes per rank and step cross the link instead of 64 bytes per bidder.
__global__ void __launch_bounds__(PA_SCAN_T)
k_y_scan_shard(const unsigned char *X, size_t xstride, int m, int s0, unsigned char *const *peers, int world, int rank, int par,
               u32 tag, int *err, u32 *jout) {
  __shared__ __align__(16) u32 part[2][PA_SCAN_T][24];
  __shared__ __align__(16) u32 tot[PA_XCHG_MAX_WORLD][24];
  __shared__ __align__(16) u32 bt[2][24];  // base = sum of the ranks before this one; minus the total
  const int b = blockIdx.x, s = s0 + b, lo = b * m, t = threadIdx.x;
  int chunk = (m + PA_SCAN_T - 1) / PA_SCAN_T;
  int c0 = min(m, t * chunk), c1 = min(m, c0 + chunk);
  jac acc;
  jac_set_inf(acc);
  for (int id = c0; id < c1; ++id) {
    aff x;
    ld_aff(x, X + xstride * (size_t)(lo + id));
    jac_madd(acc, acc, x);
  }
  st_jac(part[0][t], acc);
  __syncthreads();
  int cur = 0;
  for (int d = 1; d < PA_SCAN_T; d <<= 1) {  // Hillis-Steele inclusive scan of the chunk sums
    jac a;
    ld_jac(a, part[cur][t]);
    if (t >= d) {
      jac c;
      ld_jac(c, part[cur][t - d]);
      jac_add(a, a, c);
    }
    st_jac(part[cur ^ 1][t], a);
    __syncthreads();
    cur ^= 1;
  }
  xchg_put(peers, world, 0, par, s, rank, part[cur][PA_SCAN_T - 1], tag);
  if (!xchg_get(peers[rank], world, 0, par, s, tag, tot, err)) return;
  if (t == 0) {
    jac all, base, v;
    jac_set_inf(all);
    jac_set_inf(base);
    for (int r = 0; r < world; ++r) {
      if (r == rank) base = all;
      ld_jac(v, tot[r]);
      jac_add(all, all, v);
    }
    jac_neg(all, all);
    st_jac(bt[0], base);
    st_jac(bt[1], all);
  }
  __syncthreads();
  jac total, e, base;
  ld_jac(base, bt[0]);
  ld_jac(total, bt[1]);
  if (t == 0) jac_set_inf(e); else ld_jac(e, part[cur][t - 1]);
  jac_add(e, e, base);
  for (int id = c0; id < c1; ++id) {
    aff x;
    ld_aff(x, X + xstride * (size_t)(lo + id));
    jac p, y;
    jac_madd(p, e, x);     // inclusive prefix
    jac_add(y, e, p);      // E + P
    jac_add(y, y, total);  // - T
    st_jac(jout + 24 * (size_t)(lo + id), y);
    e = p;
  }
}

// ---- the walk, restated so that its sequential part shrinks with the race ---------------------------------
// Round three asks whether S(s) = sum_p b_p(s) is infinity (SEAL/bidder.cpp:1393-1397), where b_p(s) is the veto
// candidate R^x for the bidders that veto - those still in the race (prevDecidingBit = 1, initially everybody)
// whose bit at step s is 1 - and the other candidate Y^x for everybody else (:1301-1309).  Hence
//     S(s) = SY(s) + sum_{p in race, bit_p(s) = 1} D_p(s),   SY(s) = sum_p Y_p^x,   D_p(s) = R_p^x - Y_p^x.
// SY(s) does not depend on the state of the auction: k_seal_sum_y adds it up for every step of a pass at once,
// one block per step.  What is left for the sequential walk is the sum over the vetoing bidders, a tree whose
// depth is log2 of the number of bidders still in the race; that number halves (random bids) at every deciding
// step, so after ten deciding steps of a 1000-bidder auction a step costs a couple of additions.  The same
// group element is tested as before, so the decisions are the same.  The walk only keeps the race (a compacted
// list) and notes at which step each bidder left it (`elim`); what each bidder publishes at each step follows from
// that and is written by k_seal_select, in parallel, afterwards.
#define PA_WALK_T 256
__global__ void __launch_bounds__(PA_SCAN_T)
k_seal_sum_y(int m, int s0, const unsigned char *cand, u32 *sumy) {
  __shared__ __align__(16) u32 part[PA_SCAN_T][24];
  const int s = s0 + blockIdx.x, t = threadIdx.x;
  jac acc;
  jac_set_inf(acc);
  for (int p = t; p < m; p += PA_SCAN_T) {
    aff x;
    ld_aff(x, cand + 128 * ((size_t)s * m + p));
    jac_madd(acc, acc, x);
  }
  st_jac(part[t], acc);
  __syncthreads();
  for (int d = PA_SCAN_T / 2; d > 0; d >>= 1) {
    if (t < d) {
      jac a, c;
      ld_jac(a, part[t]);
      ld_jac(c, part[t + d]);
      jac_add(a, a, c);
      st_jac(part[t], a);
    }
    __syncthreads();
  }
  if (t < 24) sumy[24 * (size_t)s + t] = part[0][t];
}

// block-wide: race1 = the entries p of race0[0..R) with keep(p); returns the new count (the same in every thread).
// Dropped entries get elim[p] = s.  Order is preserved.
PA_D int walk_compact(const u32 *race0, u32 *race1, int R, int s, const unsigned char *bits, const u32 *boff, u32 *elim,
                      int *s_warp, int *s_total) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = PA_WALK_T / 32;
  int base = 0;
  for (int k0 = 0; k0 < R; k0 += PA_WALK_T) {
    const int k = k0 + t;
    u32 p = 0;
    int keep = 0;
    if (k < R) {
      p = race0[k];
      keep = bits[boff[p] + s] ? 1 : 0;
      if (!keep) elim[p] = (u32)s;
    }
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (keep) race1[off + __popc(mask & ((1u << lane) - 1u))] = p;
    int tot = 0;
    for (int w = 0; w < nw; ++w) tot += s_warp[w];
    base += tot;
    __syncthreads();
  }
  if (t == 0) *s_total = base;
  __syncthreads();
  return *s_total;
}

// state[0] = junction flag, state[1] = previous deciding step (-1: none), state[2] = next step to run,
// state[3] = the first deciding step.  Walks steps [state[2], limit); while `speculative` it stops one step after the
// first deciding step (the keys of later steps were drawn on counters that are no longer right).
// SHARDED: this rank holds m of the bidders; per step the ranks' sums (24 words) are PUT into every peer's window and
// folded, so that all ranks take the same decision.
// race: 2 * m words of scratch; elim[p]: the deciding step at which bidder p left the race (0xFFFFFFFF: still in).
template <bool SHARDED>
__global__ void __launch_bounds__(PA_WALK_T)
k_seal_walk(int m, int limit, int speculative, const unsigned char *bits, const u32 *boff, const unsigned char *cand,
            const u32 *sumy, unsigned char *prevbit, u32 *elim, u32 *race, int *state, int *stage, int *prevstep, int *r3,
            unsigned char *const *peers, int world, int rank, int par, u32 tag, int *err) {
  __shared__ __align__(16) u32 part[PA_WALK_T][24];
  __shared__ __align__(16) u32 tot[PA_XCHG_MAX_WORLD][24];
  __shared__ int s_warp[PA_WALK_T / 32], s_total, s_junc, s_last, s_deciding;
  const int t = threadIdx.x;
  if (t == 0) s_junc = state[0], s_last = state[1];
  // the race as a compact list, from the flags the previous pass left
  u32 *r0 = race, *r1 = race + m;
  {
    const int lane = t & 31, warp = t >> 5, nw = PA_WALK_T / 32;
    int base = 0;
    for (int k0 = 0; k0 < m; k0 += PA_WALK_T) {
      const int p = k0 + t;
      const int keep = (p < m && prevbit[p]) ? 1 : 0;
      const unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
      if (lane == 0) s_warp[warp] = __popc(mask);
      __syncthreads();
      int off = base;
      for (int w = 0; w < warp; ++w) off += s_warp[w];
      if (keep) r0[off + __popc(mask & ((1u << lane) - 1u))] = (u32)p;
      for (int w = 0; w < nw; ++w) base += s_warp[w];
      __syncthreads();
    }
    if (t == 0) s_total = base;
    __syncthreads();
  }
  int R = s_total;
  int s = state[2];
  for (; s < limit; ++s) {
    const int junc = s_junc;
    // this rank's sum over its vetoing bidders
    jac acc;
    jac_set_inf(acc);
    for (int k = t; k < R; k += PA_WALK_T) {
      const u32 p = r0[k];
      if (bits[boff[p] + s]) {
        const unsigned char *c2 = cand + 128 * ((size_t)s * m + p);
        aff y, v;
        ld_aff(y, c2);
        ld_aff(v, c2 + 64);
        aff_neg(y, y);
        jac d;
        jac_from_aff(d, v);
        jac_madd(d, d, y);  // D = R^x - Y^x
        jac_add(acc, acc, d);
      }
    }
    int width = 1;
    while (width < R && width < PA_WALK_T) width <<= 1;
    if (t < width) st_jac(part[t], acc);
    __syncthreads();
    for (int d = width / 2; d > 0; d >>= 1) {
      if (t < d) {
        jac a, c;
        ld_jac(a, part[t]);
        ld_jac(c, part[t + d]);
        jac_add(a, a, c);
        st_jac(part[t], a);
      }
      __syncthreads();
    }
    if (t == 0) {  // + the state-independent part of this step
      jac a, c;
      ld_jac(a, part[0]);
      ld_jac(c, sumy + 24 * (size_t)s);
      jac_add(a, a, c);
      st_jac(part[0], a);
    }
    __syncthreads();
    if (SHARDED) {
      xchg_put(peers, world, 1, par, s, rank, part[0], tag);
      if (!xchg_get(peers[rank], world, 1, par, s, tag, tot, err)) {
        if (t == 0) state[0] = s_junc, state[1] = s_last, state[2] = s;
        return;
      }
      for (int d = PA_XCHG_MAX_WORLD / 2; d > 0; d >>= 1) {  // fold the ranks' sums
        if (t < d && t + d < world) {
          jac a, c;
          ld_jac(a, tot[t]);
          ld_jac(c, tot[t + d]);
          jac_add(a, a, c);
          st_jac(tot[t], a);
        }
        __syncthreads();
      }
    }
    if (t == 0) {
      jac a;
      ld_jac(a, SHARDED ? tot[0] : part[0]);
      s_deciding = jac_is_inf(a) ? 0 : 1;
      stage[s] = junc ? 2 : 1;
      prevstep[s] = s_last;
      r3[s] = s_deciding;
    }
    __syncthreads();
    bool stop = false;
    if (s_deciding) {
      R = walk_compact(r0, r1, R, s, bits, boff, elim, s_warp, &s_total);  // who vetoed stays in the race
      u32 *sw = r0;
      r0 = r1;
      r1 = sw;
      if (speculative && !junc) stop = true;  // first deciding step: one more step is still valid
      if (t == 0) {
        if (!junc) state[3] = s;  // the junction
        s_junc = 1, s_last = s;
      }
    }
    __syncthreads();
    if (stop && limit > s + 2) limit = s + 2;
  }
  // the flags for the next pass
  for (int p = t; p < m; p += PA_WALK_T) prevbit[p] = elim[p] == 0xFFFFFFFFu ? 1 : 0;
  if (t == 0) state[0] = s_junc, state[1] = s_last, state[2] = s;
}

// what every bidder publishes at the steps [s0, min(s1, state[2])) the walk has been through: it vetoes iff it is still
// in the race (it left at step elim[p], i.e. it was in for every step <= elim[p]) and its bit is 1     SEAL/bidder.cpp:1301-1309
__global__ void k_seal_select(int m, int s0, int s1, const int *state, const unsigned char *bits, const u32 *boff, const u32 *elim,
                              const unsigned char *cand, unsigned char *ebit, unsigned char *bj, unsigned char *b) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = s0 + k / m, p = k % m;
  if (s >= s1 || s >= state[2]) return;
  const size_t i = (size_t)s * m + p;
  const int pb = elim[p] >= (u32)s ? 1 : 0;
  const int veto = (bits[boff[p] + s] && pb) ? 1 : 0;
  ebit[i] = (unsigned char)veto;
  bj[i] = (unsigned char)pb;
  cp64(b + 64 * i, cand + 64 * (2 * i + veto));
}

// end of a sharded run: every rank PUTS (draws clean, verdict, max bid) and reads everybody's; out = the AND
// of the first two and the OR of the max bids, so that all ranks take the same decision about a rerun.
__global__ void k_xchg_final(u32 clean, u32 ok, u64 maxbid, unsigned char *const *peers, int world, int rank, int par, u32 tag,
                             int *err, u64 *out) {
  __shared__ __align__(16) u32 val[24];
  __shared__ __align__(16) u32 tot[PA_XCHG_MAX_WORLD][24];
  if (threadIdx.x < 24) val[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) val[0] = clean, val[1] = ok, val[2] = (u32)maxbid, val[3] = (u32)(maxbid >> 32);
  __syncthreads();
  xchg_put(peers, world, 2, par, 0, rank, val, tag);
  if (!xchg_get(peers[rank], world, 2, par, 0, tag, tot, err)) return;
  if (threadIdx.x == 0) {
    u64 c = 1, o = 1, mb = 0;
    for (int r = 0; r < world; ++r) c &= tot[r][0], o &= tot[r][1], mb |= (u64)tot[r][2] | ((u64)tot[r][3] << 32);
    out[0] = c, out[1] = o, out[2] = mb;
  }
}

// TEST HOOK (pa_debug_set PA_DBG_CORRUPT): flip one byte of a published record
__global__ void k_dbg_flip(unsigned char *p) { *p ^= 0x01; }

// ---- host side ------------------------------------------------------------------------------------
// repeat a verification launch `vreps` times (see pa_seal_job.verify)
#define PA_VREP(...)                           \
  for (int rep_ = 0; rep_ < vreps; ++rep_) {  \
    int rcv_ = (__VA_ARGS__);                 \
    if (rcv_) return rcv_;                    \
  }

namespace {

// Everything the runner needs on the device is carved out of ONE grow-only arena owned by the
// context (no cudaMalloc / cudaFree per run once it has reached its high-water mark).
struct DevPool {
  pa_ctx *ctx;
  size_t off = 0;
  bool sizing;  // first pass: only add up sizes
  explicit DevPool(pa_ctx *c, bool sizing_pass) : ctx(c), sizing(sizing_pass) {}
  template <class T> T *alloc(size_t count, bool zero = false) {
    size_t bytes = align_up((count ? count : 1) * sizeof(T), 256);
    unsigned char *p = sizing ? (unsigned char *)0x100 : ctx->d_pool + off;
    off += bytes;
    if (!sizing && zero) cudaMemsetAsync(p, 0, bytes, ctx->stream);
    return (T *)p;
  }
};

template <class T> int up(pa_ctx *ctx, T *d, const std::vector<T> &h) {
  if (!h.empty()) PA_CUDA(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return PA_OK;
}

// A lane = a stream with its own work arena.  Within a step only a short chain is truly
// sequential (keys -> Y scan -> cryptogram -> round three); the Schnorr proofs of round one, the
// OR proofs of round two and all verification hang off that chain.  They run in three side lanes
// so that kernels of ~10^3..10^4 threads from different phases and steps share the 148 SMs.
// While a LaneScope is alive, ctx->stream / ctx->d_work ARE the lane's.
struct LaneScope {
  pa_ctx *ctx;
  pa_ctx::Lane *lane;
  cudaStream_t s0;
  unsigned char *w0;
  size_t b0;
  LaneScope(pa_ctx *c, pa_ctx::Lane *l) : ctx(c), lane(l), s0(c->stream), w0(c->d_work), b0(c->work_bytes) {
    ctx->stream = lane->stream;
    ctx->d_work = lane->work;
    ctx->work_bytes = lane->work_bytes;
  }
  ~LaneScope() {
    lane->work = ctx->d_work;  // may have grown
    lane->work_bytes = ctx->work_bytes;
    ctx->stream = s0;
    ctx->d_work = w0;
    ctx->work_bytes = b0;
  }
};

int lanes_init(pa_ctx *ctx) {
  if (ctx->lanes[0].stream) return PA_OK;
  for (int i = 0; i < 3; ++i) PA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lanes[i].stream, cudaStreamNonBlocking));
  for (int i = 0; i < 10; ++i) PA_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lane_ev[i], cudaEventDisableTiming));
  return PA_OK;
}

struct StepBufs {  // per-step device buffers; two sets, used alternately
  u32 *act, *pauc, *pseg, *soff, *g, *gslot;
  u64 *pid, *gid;
  unsigned char *rnd1, *r1, *pokv, *Y, *b, *ebit, *stmt, *sec, *bi, *bj, *cb, *rnd2, *proof, *pv;
  int *isinf;
};

}  // namespace

static int seal_run_impl(pa_ctx *ctx, const pa_seal_job *job, bool *rerun_step_major);

// The phase-major schedule computes draw counters arithmetically, which is exact unless a draw is rejected
// (probability 2^-128 per draw).  When that happens (all ranks of a sharded auction learn it together) the
// auction is simply run again step-major, where every counter is carried sequentially.
extern "C" int pa_seal_run(pa_ctx *ctx, const pa_seal_job *job) {
  PA_ENTER(ctx);
  // ONE exchange epoch per call, whether or not the auction is run twice: a rank that owns no bidder calls pa_xchg_skip once
  // per auction and must stay in step with the others.  (The second, step-major run uses other slots and tags than the first:
  // bulk all-gathers, which the phase-major schedule does not use, and the end-of-run tag 254 instead of 255.)
  if (job && job->use_xchg) ctx->xchg.epoch++;
  bool rerun = false;
  int rc = seal_run_impl(ctx, job, &rerun);
  if (rc == PA_OK && rerun) {
    pa_seal_job again = *job;
    again.schedule = PA_SEAL_STEP_MAJOR;
    ctx->reruns++;
    rc = seal_run_impl(ctx, &again, &rerun);
  }
  return rc;
}
extern "C" uint64_t pa_ctx_reruns(pa_ctx *ctx) { return ctx ? ctx->reruns : 0; }
extern "C" int pa_xchg_skip(pa_ctx *ctx) {  // a rank that owns no bidder of a sharded auction: keep the run counter in step
  PA_ENTER(ctx);
  ctx->xchg.epoch++;
  return PA_OK;
}

static int seal_run_impl(pa_ctx *ctx, const pa_seal_job *job, bool *rerun_step_major) {
  PA_ARGCHECK(ctx, ctx && job && job->n_auctions >= 1 && job->n && job->c && job->bids);
  const size_t A = job->n_auctions;
  const bool p2p = job->use_xchg != 0;  // exchanges go through the peer window, inside kernels
  const bool sharded = job->allgather != nullptr || p2p;
  PA_ARGCHECK(ctx, !sharded || (A == 1 && job->hi > job->lo && job->hi <= job->n[0] && job->slice >= job->hi - job->lo));
  PA_ARGCHECK(ctx, !sharded || p2p || (job->d_send && job->d_recv));
  const int xworld = sharded ? (int)((job->n[0] + job->slice - 1) / job->slice) : 1;  // ranks that own bidders
  const int xrank = sharded ? (int)(job->lo / job->slice) : 0;
  PA_ARGCHECK(ctx, !p2p || (ctx->xchg.world >= xworld && ctx->xchg.rank == xrank && job->lo == (uint32_t)xrank * job->slice &&
                            (size_t)job->slice * 64 * xworld <= PA_XCHG_BULK));
  u32 xpass = 0, xseq = 0;  // passes / plain all-gathers of this run
  const u32 xepoch = ctx->xchg.epoch;  // advanced by pa_seal_run / pa_xchg_skip
  const int xpar = (int)(xepoch & 1u);
  unsigned char *const *xpeers = ctx->xchg.d_peers;
  int *xerr = ctx->xchg.d_err;
  bool draws_clean = true;
  // stream-ordered all-gather of `bytes` per rank through the windows; *recv = where the concatenation lands
  auto xchg_allgather = [&](const unsigned char *send, size_t bytes, const unsigned char **recv) -> int {
    const int buf = (int)(xseq & 1u);
    const u32 seq = (xepoch << 12) | (++xseq);
    PA_LAUNCH(ctx, PA_K_ENCODE, (k_xchg_allgather<<<xworld, 256, 0, ctx->stream>>>(send, bytes, xpeers, xworld, xrank, buf, seq, xerr)));
    *recv = ctx->xchg.local + PA_XCHG_BULK_OFF + (size_t)buf * PA_XCHG_BULK;
    return PA_OK;
  };
  auto xchg_check = [&]() -> int {  // after a host sync: did a wait on a peer time out?
    int e = 0;
    PA_CUDA(ctx, cudaMemcpyAsync(&e, xerr, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (e) {
      cudaMemsetAsync(xerr, 0, sizeof(int), ctx->stream);
      return pa_fail(ctx, PA_ECUDA, "pa_seal_run: a peer did not reach the exchange within 10 s (peer exchange window)");
    }
    return PA_OK;
  };
  // TEST HOOK: flip one byte of a published record between proving and verifying
  auto dbg_corrupt = [&](int section, unsigned char *rec) -> int {
    if (ctx->corrupt.section == section) PA_LAUNCH(ctx, PA_K_VERDICT, (k_dbg_flip<<<1, 1, 0, ctx->stream>>>(rec + ctx->corrupt.offset)));
    return PA_OK;
  };
  const bool verify = job->verify != 0;
  // verify = k > 1: every proof is verified k times (k = n - 1 is the work of the reference, where each of the n
  // bidders repeats the same deterministic checks on everybody else); the verdicts do not change
  const int vreps = job->verify > 1 ? job->verify : 1;
  // one unsharded auction: phase-major schedule unless the caller asks for the step-major one
  // (sharded: the exchange buffers must hold the X of all steps, and a Jacobian partial sum)
  bool phased = A == 1 && job->c[0] >= 1 && job->schedule != PA_SEAL_STEP_MAJOR &&
                (!sharded || p2p || (job->xchg_bytes >= (size_t)job->c[0] * job->slice * 64 && job->xchg_bytes >= 128));
  PA_ARGCHECK(ctx, job->schedule != PA_SEAL_PHASE_MAJOR || phased);
  const bool want_r1 = job->out_r1 != nullptr, want_b = job->out_r2_b != nullptr, want_proof = job->out_r2_proof != nullptr;
  int rc;
  if ((rc = lanes_init(ctx))) return rc;
  pa_ctx::Lane *L_pok = &ctx->lanes[0], *L_prove = &ctx->lanes[1], *L_verify = &ctx->lanes[2];
  cudaEvent_t *ev_r1 = &ctx->lane_ev[0], *ev_pok = &ctx->lane_ev[2], *ev_enc = &ctx->lane_ev[4], *ev_proved = &ctx->lane_ev[6],
              *ev_verified = &ctx->lane_ev[8];

  // ---- host-side index: local bidder slots, auction-major in id order ------------------
  std::vector<u32> auc, boff(1, 0), aoff(A + 1, 0);  // aoff: first local slot of auction a
  std::vector<u64> ids, streams;
  std::vector<unsigned char> bits;
  size_t cmax = 0;
  for (size_t a = 0; a < A; ++a) {
    u32 n = job->n[a], c = job->c[a];
    PA_ARGCHECK(ctx, c <= 64 && n >= 1);
    cmax = c > cmax ? c : cmax;
    u32 lo = sharded ? job->lo : 0, hi = sharded ? job->hi : n;
    u64 aid = job->auction_ids ? job->auction_ids[a] : a;
    for (u32 j = lo; j < hi; ++j) {
      u64 bid = job->bids[ids.size()];
      auc.push_back((u32)a);
      ids.push_back(j);
      streams.push_back((aid << 32) | j);
      for (u32 i = 0; i < c; ++i) bits.push_back((unsigned char)((bid >> (c - 1 - i)) & 1));  // MSB first, SEAL/bidder.cpp:31
      boff.push_back((u32)bits.size());
    }
    aoff[a + 1] = (u32)ids.size();
  }
  const size_t m = ids.size(), Mb = bits.size();
  std::vector<u64> cid(Mb);
  for (size_t s = 0; s < m; ++s)
    for (u32 k = boff[s]; k < boff[s + 1]; ++k) cid[k] = ids[s];

  // ---- device state --------------------------------------------------------------------------
  unsigned char *d_bits, *d_rndc, *d_crec, *d_cv, *d_junc, *d_prevbit, *d_prevpts, *d_prevx, *d_r1ok, *d_r2ok;
  u32 *d_boff;
  u64 *d_ids, *d_streams, *d_ctr, *d_cid, *d_sstream, *d_sctr;
  StepBufs SB[2];
  struct Phased {  // arrays over all items i = step * m + bidder of the single-auction schedule
    u64 *istream, *ictr, *pid;
    u32 *soff;
    int *stage, *prevstep, *r3, *state;
    unsigned char *rnd1, *r1, *pokv, *r1ok, *Y, *cand, *b, *ebit, *bj, *stmt, *sec, *bi, *bjp, *cbp, *rnd2, *proof, *r2ok;
    unsigned char *Xall, *Yall, *part;  // sharded: every bidder's X / Y of every step; the ranks' partial sums of a step
    u32 *gidx, *lidx, *soffN;
    u32 *sumy, *elim, *race;  // per step: sum of the no-veto candidates; per bidder: step at which it left the race; scratch
  } PH{};
  unsigned char *d_xsend = nullptr;  // peer window: staging of this rank's part of a plain all-gather
  u64 *d_xfinal = nullptr;
  const size_t nall = job->n[0];
  const size_t T = phased ? cmax * m : 0;
  // Junction planes (see the phase-major schedule below): the keys, Y and candidates of the steps after the junction for the
  // guesses J = 0..3, computed beside the first pass when that is cheap (a small slice: the GPU is far from full).
  // Plane q holds steps [q + 2, cmax) at virtual steps [vplane[q], vplane[q + 1]).
  std::vector<size_t> vplane;
  size_t VT = 0;  // virtual steps in total
  if (phased && (!sharded || p2p) && cmax >= 6) {
    size_t v = cmax;
    vplane.push_back(v);
    for (size_t g = 0; g < 4 && g + 2 < cmax; ++g) v += cmax - g - 2, vplane.push_back(v);
    VT = v - cmax;
    // the same decision on every rank: by the slice, not by this rank's share of it
    if (VT * (sharded ? (size_t)job->slice : m) > 40000 || cmax + VT > PA_XCHG_STEPS) vplane.clear(), VT = 0;  // not cheap: no planes
  }
  const size_t TX = T + VT * m;  // items incl. the planes
  const size_t nY = sharded ? nall : m;
  auto carve = [&](DevPool &pool) {
    if (phased) {
      PH.istream = pool.alloc<u64>(TX); PH.ictr = pool.alloc<u64>(TX); PH.pid = pool.alloc<u64>(T);
      PH.soff = pool.alloc<u32>(cmax + VT + 1);
      PH.stage = pool.alloc<int>(cmax); PH.prevstep = pool.alloc<int>(cmax); PH.r3 = pool.alloc<int>(cmax); PH.state = pool.alloc<int>(4);
      PH.rnd1 = pool.alloc<unsigned char>(TX * 128); PH.r1 = pool.alloc<unsigned char>(TX * 320);
      PH.pokv = pool.alloc<unsigned char>(T * 2); PH.r1ok = pool.alloc<unsigned char>(T);
      PH.Y = pool.alloc<unsigned char>(TX * 64); PH.cand = pool.alloc<unsigned char>(TX * 128); PH.b = pool.alloc<unsigned char>(T * 64);
      PH.ebit = pool.alloc<unsigned char>(T); PH.bj = pool.alloc<unsigned char>(T);
      PH.stmt = pool.alloc<unsigned char>(T * 704 + 256); PH.sec = pool.alloc<unsigned char>(T * 192 + 256);
      PH.bi = pool.alloc<unsigned char>(T + 256); PH.bjp = pool.alloc<unsigned char>(T + 256); PH.cbp = pool.alloc<unsigned char>(T + 256);
      PH.rnd2 = pool.alloc<unsigned char>(T * 352 + 256); PH.proof = pool.alloc<unsigned char>(T * 1344 + 256);
      PH.r2ok = pool.alloc<unsigned char>(T + 256);
      PH.sumy = pool.alloc<u32>(cmax * 24 + 8); PH.elim = pool.alloc<u32>(m); PH.race = pool.alloc<u32>(2 * m + 8);
      if (sharded && !p2p) {
        PH.Xall = pool.alloc<unsigned char>(cmax * nall * 64); PH.Yall = pool.alloc<unsigned char>(cmax * nall * 64);
        PH.gidx = pool.alloc<u32>(cmax * nall); PH.lidx = pool.alloc<u32>(T); PH.soffN = pool.alloc<u32>(cmax + 1);
        PH.part = pool.alloc<unsigned char>(((nall + job->slice - 1) / job->slice) * 128);
      }
    }
    if (p2p) {
      d_xsend = pool.alloc<unsigned char>((size_t)job->slice * 64);
      d_xfinal = pool.alloc<u64>(4);
    }
    d_bits = pool.alloc<unsigned char>(Mb);
    d_boff = pool.alloc<u32>(m + 1);
    d_ids = pool.alloc<u64>(m);
    d_streams = pool.alloc<u64>(m);
    d_ctr = pool.alloc<u64>(m, true);
    d_cid = pool.alloc<u64>(Mb);
    d_rndc = pool.alloc<unsigned char>(Mb * 224);
    d_crec = pool.alloc<unsigned char>(Mb * 736);
    d_cv = pool.alloc<unsigned char>(Mb * 4);
    d_junc = pool.alloc<unsigned char>(A, true);
    d_prevbit = pool.alloc<unsigned char>(m);
    d_prevpts = pool.alloc<unsigned char>(m * 256, true);
    d_prevx = pool.alloc<unsigned char>(m * 64, true);
    d_r1ok = pool.alloc<unsigned char>(cmax * m);
    d_r2ok = pool.alloc<unsigned char>(cmax * m);
    d_sstream = pool.alloc<u64>(Mb);
    d_sctr = pool.alloc<u64>(Mb);
    for (int k = 0; k < 2; ++k) {
      StepBufs &B = SB[k];
      B.act = pool.alloc<u32>(m); B.pauc = pool.alloc<u32>(m); B.pseg = pool.alloc<u32>(m); B.soff = pool.alloc<u32>(A + 1);
      B.g = pool.alloc<u32>(m); B.gslot = pool.alloc<u32>(m);
      B.pid = pool.alloc<u64>(m); B.gid = pool.alloc<u64>(m);
      B.rnd1 = pool.alloc<unsigned char>(m * 128); B.r1 = pool.alloc<unsigned char>(m * 320); B.pokv = pool.alloc<unsigned char>(m * 2);
      B.Y = pool.alloc<unsigned char>(nY * 64); B.b = pool.alloc<unsigned char>(m * 64); B.ebit = pool.alloc<unsigned char>(m);
      // stage-1 members first, stage-2 members behind them (256-byte aligned)
      B.stmt = pool.alloc<unsigned char>(m * 704 + 256); B.sec = pool.alloc<unsigned char>(m * 192 + 256);
      B.bi = pool.alloc<unsigned char>(m + 256); B.bj = pool.alloc<unsigned char>(m + 256); B.cb = pool.alloc<unsigned char>(m + 256);
      B.rnd2 = pool.alloc<unsigned char>(m * 352 + 256); B.proof = pool.alloc<unsigned char>(m * 1344 + 256);
      B.pv = pool.alloc<unsigned char>(m + 256);
      B.isinf = pool.alloc<int>(A);
    }
  };
  {
    DevPool sizing(ctx, true);
    carve(sizing);
    if ((rc = ensure(ctx, &ctx->d_pool, &ctx->pool_bytes, sizing.off + 4096))) return rc;
    DevPool real(ctx, false);
    carve(real);
  }
  if ((rc = up(ctx, d_bits, bits)) || (rc = up(ctx, d_boff, boff)) || (rc = up(ctx, d_ids, ids)) ||
      (rc = up(ctx, d_streams, streams)) || (rc = up(ctx, d_cid, cid)))
    return rc;
  PA_CUDA(ctx, cudaMemsetAsync(d_prevbit, 1, m, ctx->stream));  // prevDecidingBit(1), SEAL/bidder.cpp:23
  if (phased) PA_CUDA(ctx, cudaMemsetAsync(PH.elim, 0xFF, m * sizeof(u32), ctx->stream));  // nobody has left the race
  PA_CUDA(ctx, cudaMemsetAsync(d_r1ok, 1, cmax * m, ctx->stream));
  PA_CUDA(ctx, cudaMemsetAsync(d_r2ok, 1, cmax * m, ctx->stream));

  // pinned staging for the optional section outputs (written asynchronously by the lanes)
  unsigned char *h_stage = nullptr;
  size_t off_r1 = 0, off_b = 0, off_pf = 0, stage_bytes = 0;
  if (want_r1) off_r1 = stage_bytes, stage_bytes += cmax * m * 320;
  if (want_b) off_b = stage_bytes, stage_bytes += cmax * m * 64;
  if (want_proof) off_pf = stage_bytes, stage_bytes += cmax * m * 1344;
  if (stage_bytes) PA_CUDA(ctx, cudaMallocHost((void **)&h_stage, stage_bytes));
  struct HostFree {
    unsigned char *p;
    ~HostFree() { if (p) cudaFreeHost(p); }
  } host_free{h_stage};
  // whatever happens, leave no lane running when we return
  struct LaneDrain {
    pa_ctx *c;
    ~LaneDrain() { for (int i = 0; i < 3; ++i) cudaStreamSynchronize(c->lanes[i].stream); cudaStreamSynchronize(c->stream); }
  } drain{ctx};

  std::vector<unsigned char> junction(A, 0), okv(A, 1);
  std::vector<u64> maxbid(A, 0);
  std::vector<std::vector<u32>> acts(cmax), g1s_pos(cmax), g2s_pos(cmax);  // per step: slot of position p; group members
  // strides of the records the proofs live in; LC2 / LR2: the two Schnorr proofs of a record in one batch
  const pa_lay LC{736, 736, 224, 224, 1, 0, 0, 0, 0}, LC2{736, 736, 224, 224, 2, 96, 64, 32, 32};
  const pa_lay LR2{320, 320, 128, 128, 2, 96, 64, 32, 32};
  // round-two proofs with witnesses: packed records, extended secrets (x, alpha, r, beta) / (xi, xj, alpha, ri, rj, beta)
  pa_lay LW1 = pa_lay_packed<PA_S1>(), LW2 = pa_lay_packed<PA_S2>();
  LW1.secret = 128;
  LW2.secret = 192;

  // verdicts and records of the commit phase to the host (after their verification has been queued)
  auto collect_commitments = [&]() -> int {
    std::vector<unsigned char> cv(Mb);
    PA_CUDA(ctx, cudaMemcpyAsync(cv.data(), d_cv + 3 * Mb, Mb, cudaMemcpyDeviceToHost, ctx->stream));
    if (job->out_commit) PA_CUDA(ctx, cudaMemcpyAsync(job->out_commit, d_crec, Mb * 736, cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (job->out_commit_ok) memcpy(job->out_commit_ok, cv.data(), Mb);
    for (size_t s = 0; s < m; ++s)
      for (u32 k = boff[s]; k < boff[s + 1]; ++k) okv[auc[s]] &= cv[k];
    return PA_OK;
  };

  // ================= commit phase ===================================================
  std::vector<u64> sstream(Mb), sctr(Mb), cafter(Mb);  // commit-phase draw streams, counters before and after
  std::function<int()> commit_on_lane;
  {
    // 7 draws per (bidder, bit): slot k of bidder s starts at counter 7 * (k - boff[s]) of the bidder's
    // stream.  That arithmetic is exact unless a draw is rejected (probability 2^-128 per draw); if the
    // counters show a rejection, that bidder's draws are regenerated sequentially (step-major) or the whole
    // auction is run again step-major (phase-major, which looks at the counters at its first synchronisation).
    for (size_t s = 0; s < m; ++s)
      for (u32 k = boff[s]; k < boff[s + 1]; ++k) sstream[k] = streams[s], sctr[k] = 7ull * (k - boff[s]);
    if ((rc = up(ctx, d_sstream, sstream)) || (rc = up(ctx, d_sctr, sctr))) return rc;
    PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(Mb), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_sstream, d_sctr, nullptr, 7, d_rndc, (int)Mb)));
    std::vector<u64> &after = cafter;
    PA_CUDA(ctx, cudaMemcpyAsync(after.data(), d_sctr, Mb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (phased) PA_CUDA(ctx, cudaEventRecord(ev_r1[0], ctx->stream));  // the commit draws are there
    if (!phased) {
      PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      // every bidder's stream counter after the commit phase, in ONE upload (a copy per bidder cost 1.5 ms at n = 1000);
      // bidders with a rejected draw are redone one by one afterwards
      std::vector<u64> ctr_after(m);
      std::vector<size_t> redo;
      for (size_t s = 0; s < m; ++s) {
        bool clean = true;
        for (u32 k = boff[s]; k < boff[s + 1]; ++k) clean &= after[k] == sctr[k] + 7;
        ctr_after[s] = clean ? 7ull * (boff[s + 1] - boff[s]) : 0;
        if (!clean) draws_clean = false, redo.push_back(s);  // a rejected draw (probability 2^-128): counters are no longer arithmetic
      }
      if ((rc = up(ctx, d_ctr, ctr_after))) return rc;
      for (size_t s : redo) {
        u64 cnt = 7ull * (boff[s + 1] - boff[s]);
        PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<1, 1, 0, ctx->stream>>>(job->seed, d_streams + s, d_ctr + s, nullptr, (int)cnt, d_rndc + 224 * (size_t)boff[s], 1)));
      }
      PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // ctr_after goes out of scope
    }
    // commitment points, their Schnorr proofs (A: alpha, v_A; B: beta, v_B; 2 per record, one batch), the OR proof,
    // and the verification of all of it
    auto commit_work = [&]() -> int {
      int rc2;
      if ((rc2 = work_reserve(ctx, 3 * Mb))) return rc2;
      PA_LAUNCH(ctx, PA_K_COMMIT, (k_seal_commit_points<<<grid_for(3 * Mb), PA_BLOCK, 0, ctx->stream>>>(d_rndc, d_bits, ctx->d_comb, work_jac(ctx), (int)Mb)));
      if ((rc2 = normalize_to(ctx, d_crec, 3 * Mb, 3, 736))) return rc2;
      PA_CUDA(ctx, cudaEventRecord(ev_enc[1], ctx->stream));  // phi, A, B are in the records
      if ((rc2 = prove_dev<PA_POK>(ctx, d_crec + 64, d_rndc, nullptr, nullptr, d_cid, d_rndc + 64, d_crec + 192, 2 * Mb, LC2))) return rc2;
      if ((rc2 = prove_dev<PA_COM>(ctx, d_crec, d_rndc, d_bits, nullptr, d_cid, d_rndc + 128, d_crec + 384, Mb, LC, true))) return rc2;
      if (ctx->corrupt.bidder < m && ctx->corrupt.step < boff[ctx->corrupt.bidder + 1] - boff[ctx->corrupt.bidder] &&
          (rc2 = dbg_corrupt(1, d_crec + 736 * ((size_t)boff[ctx->corrupt.bidder] + ctx->corrupt.step))))
        return rc2;
      if (!verify) {
        PA_CUDA(ctx, cudaMemsetAsync(d_cv + 3 * Mb, 1, Mb, ctx->stream));
        return PA_OK;
      }
      PA_VREP(verify_dev<PA_POK, 1>(ctx, d_crec + 192, d_crec + 64, d_cid, d_cv, 2 * Mb, LC2));  // verdicts interleaved A, B
      PA_VREP(verify_dev<PA_COM, 4>(ctx, d_crec + 384, d_crec, d_cid, d_cv + 2 * Mb, Mb, LC));
      PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_and_pairs<<<grid_for(Mb), PA_BLOCK, 0, ctx->stream>>>(d_cv, d_cv + 2 * Mb, d_cv + 3 * Mb, (int)Mb)));
      return PA_OK;
    };
    if (phased) {
      // on a side lane: the steps do not need the commitments until the round-two statements are assembled.  It is queued
      // at once, without waiting for the check of its draw counters (the phase-major schedule looks at them at its first
      // synchronisation): measured on 8 GPUs, this bulk work must START first - it then runs beside the passes, while the
      // GPU is nearly idle, instead of beside the stage-2 group (queued behind the first pass: 7.5 -> 8.0 ms).
      commit_on_lane = [&, commit_work]() -> int {
        LaneScope ls(ctx, L_verify);
        PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_r1[0], 0));
        int rc2 = commit_work();
        if (rc2) return rc2;
        PA_CUDA(ctx, cudaEventRecord(ev_verified[0], ctx->stream));
        return PA_OK;
      };
      if ((rc = commit_on_lane())) return rc;
    } else if ((rc = commit_work())) {
      return rc;
    }
    if (!phased) {
      if ((rc = collect_commitments())) return rc;
    }
  }

  // ================= single auction, phase-major schedule ==============================================
  if (phased) {
    const size_t c = cmax;
    // Draw counters are arithmetic: 7 per committed bit, then per step 4 (x, r, v_X, v_R) and 5 (stage 1)
    // or 11 (stage 2).  Stage 2 starts the step after the first deciding step J.
    const u64 base = 7ull * c;
    auto key_ctr = [&](size_t s, long J) -> u64 {  // J < 0: no deciding step before s
      if (J < 0 || s <= (size_t)J + 1) return base + 9ull * s;
      return base + 9ull * ((size_t)J + 1) + 15ull * (s - (size_t)J - 1);
    };
    {
      std::vector<u32> soff(c + VT + 1);
      for (size_t k = 0; k <= c + VT; ++k) soff[k] = (u32)(k * m);
      std::vector<u64> pid(T);
      for (size_t i = 0; i < T; ++i) pid[i] = ids[i % m];
      int st0[4] = {0, -1, 0, 0};
      if ((rc = up(ctx, PH.soff, soff)) || (rc = up(ctx, PH.pid, pid))) return rc;
      PA_CUDA(ctx, cudaMemcpyAsync(PH.state, st0, sizeof st0, cudaMemcpyHostToDevice, ctx->stream));
      PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // st0 / vectors go out of scope
    }
    std::vector<u64> istream(TX), ictr(TX), after(TX);
    // the Schnorr proofs of X and R of every item, one batch (and their verification)
    auto pok_all = [&]() -> int {
      int rc2;
      if ((rc2 = prove_dev<PA_POK>(ctx, PH.r1, PH.rnd1, nullptr, nullptr, PH.pid, PH.rnd1 + 64, PH.r1 + 128, 2 * T, LR2))) return rc2;
      if (ctx->corrupt.step < c && ctx->corrupt.bidder < m && (rc2 = dbg_corrupt(2, PH.r1 + 320 * (ctx->corrupt.step * m + ctx->corrupt.bidder)))) return rc2;
      if (verify) {
        PA_VREP(verify_dev<PA_POK, 1>(ctx, PH.r1 + 128, PH.r1, PH.pid, PH.pokv, 2 * T, LR2));
        PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_and_pairs<<<grid_for(T), PA_BLOCK, 0, ctx->stream>>>(PH.pokv, nullptr, PH.r1ok, (int)T)));
      } else {
        PA_CUDA(ctx, cudaMemsetAsync(PH.r1ok, 1, T, ctx->stream));
      }
      return PA_OK;
    };
    bool pok_on_lane = false;
    // keys, Y and both cryptogram candidates of steps [s0, s1), then the walk through those steps
    // the proofs of X and R go to a side lane as soon as every key is final (they write the proof fields of the round-one
    // records; the passes read the point fields)
    // keys_final() marks the point of the main stream at which every X and R is final; start_pok() queues the proofs on
    // their lane behind that point - called only after the rest of the pass has been queued on the main stream, which is
    // the critical path (the host issues launches one at a time).
    bool keys_marked = false;
    auto keys_final = [&]() -> int {
      PA_CUDA(ctx, cudaEventRecord(ev_r1[1], ctx->stream));
      keys_marked = true;
      return PA_OK;
    };
    auto start_pok = [&]() -> int {
      int rc2;
      if (!keys_marked || pok_on_lane) return PA_OK;
      LaneScope ls(ctx, L_pok);
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_r1[1], 0));
      if ((rc2 = pok_all())) return rc2;
      PA_CUDA(ctx, cudaEventRecord(ev_pok[0], ctx->stream));
      pok_on_lane = true;
      return PA_OK;
    };
    // the check of the commit phase's draw counters, at the first synchronisation
    bool commit_checked = false;
    auto check_commit_draws = [&]() {  // after a synchronisation of the main stream
      if (commit_checked) return;
      commit_checked = true;
      for (size_t k = 0; k < Mb; ++k)
        if (cafter[k] != sctr[k] + 7) draws_clean = false;  // a rejected draw: the auction is run again step-major
    };
    // keys, Y and both cryptogram candidates of the (real or virtual) steps [s0, s1); istream / ictr of these items are set
    auto run_keys = [&](size_t s0, size_t s1, bool final_keys) -> int {
      const size_t i0 = s0 * m, cnt = (s1 - s0) * m;
      PA_CUDA(ctx, cudaMemcpyAsync(PH.istream + i0, istream.data() + i0, cnt * 8, cudaMemcpyHostToDevice, ctx->stream));
      PA_CUDA(ctx, cudaMemcpyAsync(PH.ictr + i0, ictr.data() + i0, cnt * 8, cudaMemcpyHostToDevice, ctx->stream));
      PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(cnt), PA_BLOCK, 0, ctx->stream>>>(job->seed, PH.istream + i0, PH.ictr + i0, nullptr, 4, PH.rnd1 + 128 * i0, (int)cnt)));
      PA_CUDA(ctx, cudaMemcpyAsync(after.data() + i0, PH.ictr + i0, cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
      int rc2;
      if ((rc2 = work_reserve(ctx, 2 * cnt))) return rc2;
      PA_LAUNCH(ctx, PA_K_FIXED, (k_seal_r1_points<<<grid_for(2 * cnt), PA_BLOCK, 0, ctx->stream>>>(PH.rnd1 + 128 * i0, ctx->d_comb, work_jac(ctx), (int)cnt)));
      if ((rc2 = normalize_to(ctx, PH.r1 + 320 * i0, 2 * cnt, 2, 320))) return rc2;
      if (final_keys && !keys_marked && (rc2 = keys_final())) return rc2;  // these are the last keys: every X and R is final
      if (!sharded) {
        if ((rc2 = work_reserve(ctx, cnt))) return rc2;
        PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<(unsigned)(s1 - s0), PA_SCAN_T, 0, ctx->stream>>>(PH.r1 + 320 * i0, 320, PH.soff, (int)cnt, work_jac(ctx))));
        if ((rc2 = normalize_to(ctx, PH.Y + 64 * i0, cnt))) return rc2;
      } else if (p2p) {
        // every rank scans its own slice; only the ranks' sums of public keys (96 bytes per rank and step) are exchanged,
        // kernel to kernel through the peer windows
        if ((rc2 = work_reserve(ctx, cnt))) return rc2;
        PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan_shard<<<(unsigned)(s1 - s0), PA_SCAN_T, 0, ctx->stream>>>(PH.r1 + 320 * i0, 320, (int)m, (int)s0, xpeers, xworld, xrank, xpar, (xepoch << 8) | xpass, xerr, work_jac(ctx))));
        if ((rc2 = normalize_to(ctx, PH.Y + 64 * i0, cnt))) return rc2;
      } else {
        // one exchange for the X of all remaining steps: rank r's block is [step][position in slice]
        const size_t cn = (s1 - s0) * nall;
        PA_CUDA(ctx, cudaMemsetAsync(job->d_send, 0, c * (size_t)job->slice * 64, ctx->stream));
        PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_pack_x<<<grid_for(cnt), PA_BLOCK, 0, ctx->stream>>>(PH.r1, job->d_send, (int)m, (int)job->slice, (int)s0, (int)cnt)));
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (job->allgather(job->user, 2) != 0) return pa_fail(ctx, PA_EINVAL, "pa_seal_run: all-gather callback failed (X of all steps)");
        PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_gather64<<<grid_for(cn), PA_BLOCK, 0, ctx->stream>>>(PH.Xall + 64 * s0 * nall, job->d_recv, PH.gidx + s0 * nall, (int)cn)));
        if ((rc2 = work_reserve(ctx, cn))) return rc2;
        PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<(unsigned)(s1 - s0), PA_SCAN_T, 0, ctx->stream>>>(PH.Xall + 64 * s0 * nall, 64, PH.soffN, (int)cn, work_jac(ctx))));
        if ((rc2 = normalize_to(ctx, PH.Yall + 64 * s0 * nall, cn))) return rc2;
        PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_gather64<<<grid_for(cnt), PA_BLOCK, 0, ctx->stream>>>(PH.Y + 64 * i0, PH.Yall, PH.lidx + i0, (int)cnt)));
      }
      if ((rc2 = work_reserve(ctx, 2 * cnt))) return rc2;
      PA_LAUNCH(ctx, PA_K_VAR, (k_seal_candidates<<<grid_for(2 * cnt), PA_BLOCK, 0, ctx->stream>>>(PH.Y + 64 * i0, PH.rnd1 + 128 * i0, ctx->d_comb, work_jac(ctx), (int)cnt)));
      if ((rc2 = normalize_to(ctx, PH.cand + 128 * i0, 2 * cnt))) return rc2;
      return PA_OK;
    };
    // the walk through the steps [s0, s1) whose candidates are in place
    auto run_walk = [&](size_t s0, size_t s1, int speculative) -> int {
      const size_t cnt = (s1 - s0) * m;
      if (!sharded || p2p) {
        // the state-independent part of every step's sum at once, the walk over what depends on the race, and then what
        // each bidder publishes at the steps walked (see "the walk, restated")
        PA_LAUNCH(ctx, PA_K_YSCAN, (k_seal_sum_y<<<(unsigned)(s1 - s0), PA_SCAN_T, 0, ctx->stream>>>((int)m, (int)s0, PH.cand, PH.sumy)));
        if (!sharded)
          PA_LAUNCH(ctx, PA_K_SUMINF, (k_seal_walk<false><<<1, PA_WALK_T, 0, ctx->stream>>>((int)m, (int)s1, speculative, d_bits, d_boff, PH.cand, PH.sumy, d_prevbit, PH.elim, PH.race, PH.state, PH.stage, PH.prevstep, PH.r3, nullptr, 1, 0, 0, 0u, nullptr)));
        else  // the walk and the per-step exchange of the ranks' cryptogram sums are one kernel
          PA_LAUNCH(ctx, PA_K_SUMINF, (k_seal_walk<true><<<1, PA_WALK_T, 0, ctx->stream>>>((int)m, (int)s1, speculative, d_bits, d_boff, PH.cand, PH.sumy, d_prevbit, PH.elim, PH.race, PH.state, PH.stage, PH.prevstep, PH.r3, xpeers, xworld, xrank, xpar, (xepoch << 8) | xpass, xerr)));
        PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_select<<<grid_for(cnt), PA_BLOCK, 0, ctx->stream>>>((int)m, (int)s0, (int)s1, PH.state, d_bits, d_boff, PH.elim, PH.cand, PH.ebit, PH.bj, PH.b)));
      }
      return PA_OK;
    };
    auto run_steps = [&](size_t s0, size_t s1, long J, int speculative) -> int {
      for (size_t i = s0 * m; i < s1 * m; ++i) istream[i] = streams[i % m], ictr[i] = key_ctr(i / m, J);
      int rc2;
      if ((rc2 = run_keys(s0, s1, !speculative && s1 == c)) || (rc2 = run_walk(s0, s1, speculative))) return rc2;
      return start_pok();  // bulk work, queued behind the chain
    };
    int st[4];
    bool clean = true;
    long J = -1;
    if (!sharded || p2p) {
      // Until the first deciding step the keys are speculative, so they are computed for a short window of
      // steps that doubles while no junction shows up (with random bids it is step 0 or 1; only an auction
      // of all-zero bids goes through every window); after it, one pass takes all remaining steps.
      size_t done = 0, win = 2;
      if (VT) {
        // Junction planes.  A small slice leaves the GPU far from full and every pass is a chain of lone-warp latencies,
        // so the second pass is taken out of the chain: beside the first window (steps 0 .. 4 on stage-1 counters) a
        // side lane computes keys, Y and candidates of the steps after the junction for each of the guesses J = 0 .. 3
        // (counters: key_ctr(., guess)).  If the walk finds the junction at one of them, that plane is copied into place
        // and only the walk is left to do; otherwise the planes are dropped and the windows go on as usual.
        const size_t nplanes = vplane.size() - 1, w1 = nplanes + 1 < c ? nplanes + 1 : c;  // first window: steps [0, w1), valid for J <= w1 - 2
        ++xpass;
        // the two chains the first synchronisation waits for: planes, then the window and its walk (they are equally long
        // and run side by side)
        for (size_t q = 0; q < nplanes; ++q)
          for (size_t v = vplane[q]; v < vplane[q + 1]; ++v)
            for (size_t k = 0; k < m; ++k) istream[v * m + k] = streams[k], ictr[v * m + k] = key_ctr(q + 2 + (v - vplane[q]), (long)q);
        {
          LaneScope ls(ctx, L_pok);  // needs nothing from the main stream
          if ((rc = run_keys(c, c + VT, false))) return rc;
          PA_CUDA(ctx, cudaEventRecord(ev_pok[1], ctx->stream));
        }
        for (size_t i = 0; i < w1 * m; ++i) istream[i] = streams[i % m], ictr[i] = key_ctr(i / m, -1);
        if ((rc = run_keys(0, w1, false)) || (rc = run_walk(0, w1, 1))) return rc;
        PA_CUDA(ctx, cudaMemcpyAsync(st, PH.state, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_pok[1], 0));  // the planes
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        check_commit_draws();
        if (p2p && (rc = xchg_check())) return rc;
        for (size_t i = 0; i < w1 * m; ++i) clean &= after[i] == ictr[i] + 4;
        if (st[0]) J = st[3];
        done = (size_t)st[2];
        win = 8;
        if (J >= 0 && (size_t)J < nplanes && done == (size_t)J + 2 && done < c) {
          const size_t q = (size_t)J, src = vplane[q] * m, dst = done * m, cnt = (c - done) * m;
          PA_CUDA(ctx, cudaMemcpyAsync(PH.rnd1 + 128 * dst, PH.rnd1 + 128 * src, 128 * cnt, cudaMemcpyDeviceToDevice, ctx->stream));
          PA_CUDA(ctx, cudaMemcpyAsync(PH.r1 + 320 * dst, PH.r1 + 320 * src, 320 * cnt, cudaMemcpyDeviceToDevice, ctx->stream));
          PA_CUDA(ctx, cudaMemcpyAsync(PH.Y + 64 * dst, PH.Y + 64 * src, 64 * cnt, cudaMemcpyDeviceToDevice, ctx->stream));
          PA_CUDA(ctx, cudaMemcpyAsync(PH.cand + 128 * dst, PH.cand + 128 * src, 128 * cnt, cudaMemcpyDeviceToDevice, ctx->stream));
          for (size_t i = 0; i < cnt; ++i) {
            clean &= after[src + i] == ictr[src + i] + 4;
            istream[dst + i] = istream[src + i], ictr[dst + i] = ictr[src + i], after[dst + i] = after[src + i];
          }
          if (!keys_marked && (rc = keys_final())) return rc;  // every X and R is final now
          ++xpass;
          if ((rc = run_walk(done, c, 0))) return rc;
          PA_CUDA(ctx, cudaMemcpyAsync(st, PH.state, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
          if ((rc = start_pok())) return rc;
          PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
          if (p2p && (rc = xchg_check())) return rc;
          done = (size_t)st[2];
        }
      }
      while (done < c) {
        const size_t s1 = J < 0 ? (done + win < c ? done + win : c) : c;
        ++xpass;
        if ((rc = run_steps(done, s1, J, J < 0 ? 1 : 0))) return rc;
        PA_CUDA(ctx, cudaMemcpyAsync(st, PH.state, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        check_commit_draws();
        if (p2p && (rc = xchg_check())) return rc;
        for (size_t i = done * m; i < s1 * m; ++i) clean &= after[i] == ictr[i] + 4;
        if (J < 0 && st[0]) J = st[3];
        done = (size_t)st[2];
        win *= 2;
      }
    } else {
      // every rank walks the steps with its slice; per step one exchange of the ranks' partial sums
      const int world = (int)((nall + job->slice - 1) / job->slice);
      {
        std::vector<u32> gidx(c * nall), lidx(T), soffN(c + 1);
        for (size_t s2 = 0; s2 < c; ++s2) {
          for (size_t j = 0; j < nall; ++j) gidx[s2 * nall + j] = (u32)(((j / job->slice) * c + s2) * job->slice + j % job->slice);
          for (size_t q = 0; q < m; ++q) lidx[s2 * m + q] = (u32)(s2 * nall + job->lo + q);
        }
        for (size_t k = 0; k <= c; ++k) soffN[k] = (u32)(k * nall);
        if ((rc = up(ctx, PH.gidx, gidx)) || (rc = up(ctx, PH.lidx, lidx)) || (rc = up(ctx, PH.soffN, soffN))) return rc;
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      }
      if ((rc = run_steps(0, c, -1, 1))) return rc;
      size_t redo_from = 0;
      for (size_t s2 = 0; s2 <= c; ++s2) {
        if (J >= 0 && s2 == (size_t)J + 2 && s2 < c) {  // the keys drawn on stage-1 counters end here
          if ((rc = run_steps(s2, c, J, 0))) return rc;
          redo_from = s2;
        }
        PA_LAUNCH(ctx, PA_K_SUMINF, (k_seal_decide_shard<<<1, PA_SCAN_T, 0, ctx->stream>>>((int)s2, (int)m, (int)c, world, d_bits, d_boff, PH.cand, d_prevbit, PH.state, PH.ebit, PH.bj, PH.b, PH.stage, PH.prevstep, PH.r3, (const u32 *)PH.part, (u32 *)job->d_send)));
        PA_CUDA(ctx, cudaMemcpyAsync(st, PH.state, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (J < 0 && st[0]) J = st[1];
        if (s2 < c) {
          if (job->allgather(job->user, 3) != 0) return pa_fail(ctx, PA_EINVAL, "pa_seal_run: all-gather callback failed (partial sums)");
          PA_CUDA(ctx, cudaMemcpyAsync(PH.part, job->d_recv, (size_t)world * 128, cudaMemcpyDeviceToDevice, ctx->stream));  // d_recv is reused by the X exchange
        }
      }
      for (size_t i = 0; i < T; ++i) clean &= after[i] == ictr[i] + 4;
      (void)redo_from;
    }
    std::vector<int> stage(c), r3(c);
    PA_CUDA(ctx, cudaMemcpyAsync(stage.data(), PH.stage, c * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaMemcpyAsync(r3.data(), PH.r3, c * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (!keys_marked && (rc = keys_final())) return rc;          // no second pass was needed: the keys are final as they are
    if ((rc = start_pok())) return rc;  // (already queued on every path that made a second pass)
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    check_commit_draws();
    if (J < 0)
      for (size_t s2 = 0; s2 < c; ++s2)
        if (r3[s2]) { J = (long)s2; break; }  // a junction in the last two steps never triggers the second pass
    size_t c1 = c;  // stage-1 steps: 0 .. J
    for (size_t s2 = 0; s2 < c; ++s2)
      if (stage[s2] == 2) { c1 = s2; break; }
    const size_t n1 = c1 * m, n2 = T - n1;
    const size_t o_stmt = align_up(n1 * 448, 256), o_sec = align_up(n1 * 128, 256), o_b = align_up(n1, 256),
                 o_rnd = align_up(n1 * 160, 256), o_proof = align_up(n1 * 672, 256);
    // The round-one proofs (their lane) write the proof fields of the round-one records; what follows reads the point
    // fields only, so it does not wait for them - the results do.  (Test hook: a corrupted KEY must reach the statements
    // deterministically, so then the wait stays here.)
    const bool pok_wait_early = ctx->corrupt.section == 2 && ctx->corrupt.offset < 128;
    if (!pok_on_lane) {
      if ((rc = pok_all())) return rc;
    } else if (pok_wait_early) {
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_pok[0], 0));
    }
    // ---- round-two proofs: statements, draws (right after the four key draws), prove, verify -----------------
    PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_enc[1], 0));  // the commitment points (side lane)
    PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_stmt_items<<<grid_for(T), PA_BLOCK, 0, ctx->stream>>>((int)m, (int)n1, PH.stage, PH.prevstep, d_boff, PH.b, PH.r1, PH.Y, PH.rnd1, d_crec, d_rndc, PH.ebit, PH.bj, d_bits, PH.stmt, PH.sec, PH.bi, PH.stmt + o_stmt, PH.sec + o_sec, PH.bi + o_b, PH.bjp + o_b, PH.cbp + o_b, (int)T)));
    for (size_t i = 0; i < T; ++i) ictr[i] = key_ctr(i / m, J) + 4, istream[i] = streams[i % m];
    PA_CUDA(ctx, cudaMemcpyAsync(PH.istream, istream.data(), T * 8, cudaMemcpyHostToDevice, ctx->stream));
    PA_CUDA(ctx, cudaMemcpyAsync(PH.ictr, ictr.data(), T * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n1) PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n1), PA_BLOCK, 0, ctx->stream>>>(job->seed, PH.istream, PH.ictr, nullptr, 5, PH.rnd2, (int)n1)));
    if (n2) PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n2), PA_BLOCK, 0, ctx->stream>>>(job->seed, PH.istream + n1, PH.ictr + n1, nullptr, 11, PH.rnd2 + o_rnd, (int)n2)));
    PA_CUDA(ctx, cudaMemcpyAsync(after.data(), PH.ictr, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (!verify) PA_CUDA(ctx, cudaMemsetAsync(PH.r2ok, 1, T, ctx->stream));
    // the two groups are independent: the (small, latency-bound) stage-1 group runs on a lane beside stage 2
    const bool s1_on_lane = n1 && n2;
    const size_t cor_item = ctx->corrupt.bidder < m ? ctx->corrupt.step * m + ctx->corrupt.bidder : T;
    if (s1_on_lane) {
      PA_CUDA(ctx, cudaEventRecord(ev_enc[0], ctx->stream));
      LaneScope ls(ctx, L_prove);
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_enc[0], 0));
      if ((rc = prove_dev<PA_S1>(ctx, PH.stmt, PH.sec, PH.bi, nullptr, PH.pid, PH.rnd2, PH.proof, n1, LW1, true))) return rc;
      if (cor_item < n1 && (rc = dbg_corrupt(3, PH.proof + 672 * cor_item))) return rc;
      if (verify) PA_VREP(verify_dev<PA_S1, 8>(ctx, PH.proof, PH.stmt, PH.pid, PH.r2ok, n1));
      PA_CUDA(ctx, cudaEventRecord(ev_proved[0], ctx->stream));
    } else if (n1) {
      if ((rc = prove_dev<PA_S1>(ctx, PH.stmt, PH.sec, PH.bi, nullptr, PH.pid, PH.rnd2, PH.proof, n1, LW1, true))) return rc;
      if (cor_item < n1 && (rc = dbg_corrupt(3, PH.proof + 672 * cor_item))) return rc;
      if (verify) PA_VREP(verify_dev<PA_S1, 8>(ctx, PH.proof, PH.stmt, PH.pid, PH.r2ok, n1));
    }
    if (n2) {
      if ((rc = prove_dev<PA_S2>(ctx, PH.stmt + o_stmt, PH.sec + o_sec, PH.bi + o_b, PH.bjp + o_b, PH.pid + n1, PH.rnd2 + o_rnd, PH.proof + o_proof, n2, LW2, true, PH.cbp + o_b))) return rc;
      if (cor_item >= n1 && cor_item < T && (rc = dbg_corrupt(3, PH.proof + o_proof + 1344 * (cor_item - n1)))) return rc;
      if (verify) PA_VREP(verify_dev<PA_S2, 16>(ctx, PH.proof + o_proof, PH.stmt + o_stmt, PH.pid + n1, PH.r2ok + n1, n2));
    }
    if (s1_on_lane) PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_proved[0], 0));
    // ---- results ---------------------------------------------------------------------------------------------
    PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_verified[0], 0));  // the commit phase (side lane)
    if (pok_on_lane && !pok_wait_early) PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_pok[0], 0));  // the round-one proofs
    if ((rc = collect_commitments())) return rc;
    std::vector<unsigned char> r1ok(T), r2ok(T);
    PA_CUDA(ctx, cudaMemcpyAsync(r1ok.data(), PH.r1ok, T, cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaMemcpyAsync(r2ok.data(), PH.r2ok, T, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_r1) PA_CUDA(ctx, cudaMemcpyAsync(job->out_r1, PH.r1, T * 320, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_b) PA_CUDA(ctx, cudaMemcpyAsync(job->out_r2_b, PH.b, T * 64, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_proof) {
      if (n1) PA_CUDA(ctx, cudaMemcpy2DAsync(job->out_r2_proof, 1344, PH.proof, 672, 672, n1, cudaMemcpyDeviceToHost, ctx->stream));
      if (n2) PA_CUDA(ctx, cudaMemcpyAsync(job->out_r2_proof + n1 * 1344, PH.proof + o_proof, n2 * 1344, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < T; ++i) clean &= after[i] == ictr[i] + (i < n1 ? 5 : 11);
    clean &= draws_clean;
    for (size_t i = 0; i < T; ++i) okv[0] &= r1ok[i] & r2ok[i];
    for (size_t s2 = 0; s2 < c; ++s2)
      if (r3[s2]) maxbid[0] |= (u64)1 << (c - s2 - 1);  // SEAL/bidder.cpp:1403, 64-bit shift
    if (sharded) {
      // all ranks must agree on whether a draw was rejected somewhere (then everybody runs the auction again step-major)
      if (p2p) {
        PA_LAUNCH(ctx, PA_K_VERDICT, (k_xchg_final<<<1, 32, 0, ctx->stream>>>(clean ? 1u : 0u, okv[0] ? 1u : 0u, maxbid[0], xpeers, xworld, xrank, xpar, (xepoch << 8) | 255u, xerr, d_xfinal)));
        u64 fin[3] = {0, 0, 0};
        PA_CUDA(ctx, cudaMemcpyAsync(fin, d_xfinal, sizeof fin, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = xchg_check())) return rc;
        clean = fin[0] != 0;
        if (job->ok_all) *job->ok_all = fin[1] != 0 ? 1 : 0;
      } else {
        u32 flag[32] = {clean ? 1u : 0u};
        PA_CUDA(ctx, cudaMemcpyAsync(job->d_send, flag, sizeof flag, cudaMemcpyHostToDevice, ctx->stream));
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (job->allgather(job->user, 3) != 0) return pa_fail(ctx, PA_EINVAL, "pa_seal_run: all-gather callback failed (draw flags)");
        std::vector<u32> all((size_t)xworld * 32);
        PA_CUDA(ctx, cudaMemcpyAsync(all.data(), job->d_recv, all.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < xworld; ++r) clean &= all[(size_t)r * 32] != 0;
      }
    }
    if (!clean) {
      *rerun_step_major = true;
      return PA_OK;
    }
    if (job->out_r1_ok) memcpy(job->out_r1_ok, r1ok.data(), T);
    if (job->out_r2_ok) memcpy(job->out_r2_ok, r2ok.data(), T);
    for (size_t s2 = 0; s2 < c; ++s2) {
      if (job->out_r2_tag) memset(job->out_r2_tag + s2 * m, stage[s2], m);
      if (job->out_r3) job->out_r3[s2] = r3[s2] ? 1 : 0;
    }
    if (job->max_bid) job->max_bid[0] = maxbid[0];
    if (job->ok) job->ok[0] = okv[0];
    return PA_OK;
  }

  // ================= auction steps =====================================================
  size_t steps_run = 0;
  for (size_t step = 0; step < cmax; ++step) {
    const int par = (int)(step & 1);
    StepBufs &B = SB[par];
    // active bidders, their per-auction segments, and the two proof groups
    std::vector<u32> &act = acts[step], &g1 = g1s_pos[step], &g2 = g2s_pos[step];
    std::vector<u32> pauc, pseg, soff(1, 0), g, gslot, actauc;
    std::vector<u64> pid, gid;
    for (size_t a = 0; a < A; ++a) {
      if (job->c[a] <= step) continue;
      for (u32 s = aoff[a]; s < aoff[a + 1]; ++s) {
        u32 p = (u32)act.size();
        act.push_back(s), pauc.push_back((u32)a), pseg.push_back((u32)actauc.size()), pid.push_back(ids[s]);
        (junction[a] ? g2 : g1).push_back(p);
      }
      actauc.push_back((u32)a);
      soff.push_back((u32)act.size());
    }
    const size_t ma = act.size(), na = actauc.size(), n1 = g1.size(), n2 = g2.size();
    if (ma == 0) break;
    steps_run = step + 1;
    for (u32 p : g1) g.push_back(p), gslot.push_back(act[p]), gid.push_back(ids[act[p]]);
    for (u32 p : g2) g.push_back(p), gslot.push_back(act[p]), gid.push_back(ids[act[p]]);
    // offsets of the stage-2 group inside the shared per-step arrays
    const size_t o_stmt = align_up(n1 * 448, 256), o_sec = align_up(n1 * 128, 256), o_b = align_up(n1, 256),
                 o_rnd = align_up(n1 * 160, 256), o_proof = align_up(n1 * 672, 256);

    // this buffer set was last used two steps ago: its side lanes must have drained
    if (step >= 2) {
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_pok[par], 0));
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_proved[par], 0));
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_verified[par], 0));
    }
    if ((rc = up(ctx, B.act, act)) || (rc = up(ctx, B.pauc, pauc)) || (rc = up(ctx, B.pseg, pseg)) || (rc = up(ctx, B.pid, pid)) ||
        (rc = up(ctx, B.soff, soff)) || (rc = up(ctx, B.g, g)) || (rc = up(ctx, B.gslot, gslot)) || (rc = up(ctx, B.gid, gid)))
      return rc;

    // ---- main lane: x, r, X = g^x, R = g^r ------------------------------------------------
    PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_streams, d_ctr, B.act, 4, B.rnd1, (int)ma)));
    if ((rc = work_reserve(ctx, 2 * ma))) return rc;
    PA_LAUNCH(ctx, PA_K_FIXED, (k_seal_r1_points<<<grid_for(2 * ma), PA_BLOCK, 0, ctx->stream>>>(B.rnd1, ctx->d_comb, work_jac(ctx), (int)ma)));
    if ((rc = normalize_to(ctx, B.r1, 2 * ma, 2, 320))) return rc;
    PA_CUDA(ctx, cudaEventRecord(ev_r1[par], ctx->stream));

    // ---- lane 1: the Schnorr proofs of X (x, v_X) and R (r, v_R) and their verification -------
    {
      LaneScope ls(ctx, L_pok);
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_r1[par], 0));
      if ((rc = prove_dev<PA_POK>(ctx, B.r1, B.rnd1, nullptr, nullptr, B.pid, B.rnd1 + 64, B.r1 + 128, 2 * ma, LR2))) return rc;
      if (A == 1 && ctx->corrupt.step == step && ctx->corrupt.bidder < ma && (rc = dbg_corrupt(2, B.r1 + 320 * ctx->corrupt.bidder))) return rc;
      if (verify) {
        PA_VREP(verify_dev<PA_POK, 1>(ctx, B.r1 + 128, B.r1, B.pid, B.pokv, 2 * ma, LR2));
        PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_and_pairs<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(B.pokv, nullptr, d_r1ok + step * m, (int)ma)));
      }
      if (want_r1) PA_CUDA(ctx, cudaMemcpyAsync(h_stage + off_r1 + step * m * 320, B.r1, ma * 320, cudaMemcpyDeviceToHost, ctx->stream));
      PA_CUDA(ctx, cudaEventRecord(ev_pok[par], ctx->stream));
    }

    // ---- main lane: Y reconstruction ----------------------------------------------------------
    const unsigned char *Yloc = B.Y;
    if (!sharded) {
      if ((rc = work_reserve(ctx, ma))) return rc;
      PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<(unsigned)na, PA_SCAN_T, 0, ctx->stream>>>(B.r1, 320, B.soff, (int)ma, work_jac(ctx))));
      if ((rc = normalize_to(ctx, B.Y, ma))) return rc;
    } else {
      // all-gather the X_i of every slice, then every rank scans the whole auction
      const size_t nall = job->n[0];
      const unsigned char *gathered = job->d_recv;
      unsigned char *send = p2p ? d_xsend : job->d_send;
      PA_CUDA(ctx, cudaMemsetAsync(send, 0, (size_t)job->slice * 64, ctx->stream));
      PA_CUDA(ctx, cudaMemcpy2DAsync(send, 64, B.r1, 320, 64, ma, cudaMemcpyDeviceToDevice, ctx->stream));
      if (p2p) {
        if ((rc = xchg_allgather(send, (size_t)job->slice * 64, &gathered))) return rc;
      } else {
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (job->allgather(job->user, 0) != 0) return pa_fail(ctx, PA_EINVAL, "pa_seal_run: all-gather callback failed (X)");
      }
      if ((rc = work_reserve(ctx, nall))) return rc;
      PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<1, PA_SCAN_T, 0, ctx->stream>>>(gathered, 64, nullptr, (int)nall, work_jac(ctx))));
      if ((rc = normalize_to(ctx, B.Y, nall))) return rc;
      Yloc = B.Y + 64 * (size_t)job->lo;
    }

    // ---- main lane: cryptogram b, statements and draws of the OR proofs --------------------------
    if ((rc = work_reserve(ctx, ma))) return rc;
    PA_LAUNCH(ctx, PA_K_VAR, (k_seal_encode<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(B.act, B.pauc, d_bits, d_boff, (int)step, d_junc, d_prevbit, B.r1, Yloc, B.rnd1, B.ebit, work_jac(ctx), (int)ma)));
    if ((rc = normalize_to(ctx, B.b, ma))) return rc;
    if (n1) {
      PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_stmt<<<grid_for(n1), PA_BLOCK, 0, ctx->stream>>>(1, B.g, B.act, d_boff, (int)step, B.b, B.r1, Yloc, B.rnd1, d_crec, d_rndc, d_prevpts, d_prevx, B.ebit, d_prevbit, d_bits, B.stmt, B.sec, B.bi, B.bj, B.cb, (int)n1)));
      PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n1), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_streams, d_ctr, B.gslot, 5, B.rnd2, (int)n1)));
    }
    if (n2) {
      PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_stmt<<<grid_for(n2), PA_BLOCK, 0, ctx->stream>>>(2, B.g + n1, B.act, d_boff, (int)step, B.b, B.r1, Yloc, B.rnd1, d_crec, d_rndc, d_prevpts, d_prevx, B.ebit, d_prevbit, d_bits, B.stmt + o_stmt, B.sec + o_sec, B.bi + o_b, B.bj + o_b, B.cb + o_b, (int)n2)));
      PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n2), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_streams, d_ctr, B.gslot + n1, 11, B.rnd2 + o_rnd, (int)n2)));
    }
    if (want_b) PA_CUDA(ctx, cudaMemcpyAsync(h_stage + off_b + step * m * 64, B.b, ma * 64, cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaEventRecord(ev_enc[par], ctx->stream));

    // ---- lane 2: the OR proofs -----------------------------------------------------------------------
    {
      LaneScope ls(ctx, L_prove);
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_enc[par], 0));
      if (n1 && (rc = prove_dev<PA_S1>(ctx, B.stmt, B.sec, B.bi, nullptr, B.gid, B.rnd2, B.proof, n1, LW1, true))) return rc;
      if (n2 && (rc = prove_dev<PA_S2>(ctx, B.stmt + o_stmt, B.sec + o_sec, B.bi + o_b, B.bj + o_b, B.gid + n1, B.rnd2 + o_rnd, B.proof + o_proof, n2, LW2, true, B.cb + o_b))) return rc;
      if (A == 1 && ctx->corrupt.step == step && ctx->corrupt.bidder < ma) {  // one auction: position = bidder, one group per step
        unsigned char *rec = n1 ? B.proof + 672 * ctx->corrupt.bidder : B.proof + o_proof + 1344 * ctx->corrupt.bidder;
        if ((rc = dbg_corrupt(3, rec))) return rc;
      }
      if (want_proof) {  // group-compact: stage-1 members first, then stage-2 members
        if (n1) PA_CUDA(ctx, cudaMemcpyAsync(h_stage + off_pf + step * m * 1344, B.proof, n1 * 672, cudaMemcpyDeviceToHost, ctx->stream));
        if (n2) PA_CUDA(ctx, cudaMemcpyAsync(h_stage + off_pf + step * m * 1344 + o_proof, B.proof + o_proof, n2 * 1344, cudaMemcpyDeviceToHost, ctx->stream));
      }
      PA_CUDA(ctx, cudaEventRecord(ev_proved[par], ctx->stream));
    }
    // ---- lane 3: their verification -----------------------------------------------------------------
    {
      LaneScope ls(ctx, L_verify);
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_proved[par], 0));
      if (verify) {
        if (n1) {
          PA_VREP(verify_dev<PA_S1, 8>(ctx, B.proof, B.stmt, B.gid, B.pv, n1));
          PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_scatter_u8<<<grid_for(n1), PA_BLOCK, 0, ctx->stream>>>(B.g, B.pv, d_r2ok + step * m, (int)n1)));
        }
        if (n2) {
          PA_VREP(verify_dev<PA_S2, 16>(ctx, B.proof + o_proof, B.stmt + o_stmt, B.gid + n1, B.pv + o_b, n2));
          PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_scatter_u8<<<grid_for(n2), PA_BLOCK, 0, ctx->stream>>>(B.g + n1, B.pv + o_b, d_r2ok + step * m, (int)n2)));
        }
      }
      PA_CUDA(ctx, cudaEventRecord(ev_verified[par], ctx->stream));
    }

    // ---- main lane, round three: is the sum of the cryptograms the point at infinity? ------------------
    if (!sharded) {
      PA_LAUNCH(ctx, PA_K_SUMINF, (k_point_sum_is_inf<<<(unsigned)na, PA_SCAN_T, 0, ctx->stream>>>(B.b, 64, B.soff, (int)ma, B.isinf)));
    } else {
      const unsigned char *gathered = job->d_recv;
      unsigned char *send = p2p ? d_xsend : job->d_send;
      PA_CUDA(ctx, cudaMemsetAsync(send, 0, (size_t)job->slice * 64, ctx->stream));
      PA_CUDA(ctx, cudaMemcpyAsync(send, B.b, ma * 64, cudaMemcpyDeviceToDevice, ctx->stream));
      if (p2p) {
        if ((rc = xchg_allgather(send, (size_t)job->slice * 64, &gathered))) return rc;
      } else {
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (job->allgather(job->user, 1) != 0) return pa_fail(ctx, PA_EINVAL, "pa_seal_run: all-gather callback failed (b)");
      }
      PA_LAUNCH(ctx, PA_K_SUMINF, (k_point_sum_is_inf<<<1, PA_SCAN_T, 0, ctx->stream>>>(gathered, 64, nullptr, (int)job->n[0], B.isinf)));
    }
    PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_update<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(B.act, B.pseg, B.pauc, B.isinf, d_bits, d_boff, (int)step, B.r1, Yloc, B.b, B.rnd1, d_prevpts, d_prevx, d_prevbit, d_junc, (int)ma)));
    std::vector<int> isinf(na);
    PA_CUDA(ctx, cudaMemcpyAsync(isinf.data(), B.isinf, na * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the only per-step host wait: the junction decides the next step's proof kind
    for (size_t p = 0; p < ma; ++p)
      if (job->out_r2_tag) job->out_r2_tag[step * m + act[p]] = junction[pauc[p]] ? 2 : 1;
    for (size_t k = 0; k < na; ++k) {
      u32 a = actauc[k];
      bool deciding = !isinf[k];
      if (job->out_r3) job->out_r3[step * A + a] = deciding ? 1 : 0;
      if (deciding) {
        junction[a] = 1;                                              // SEAL/bidder.cpp:1400
        maxbid[a] |= (u64)1 << (job->c[a] - step - 1);                 // :1403, 64-bit shift
      }
    }
  }

  // ================= drain the lanes, collect verdicts and sections ====================================
  for (int i = 0; i < 3; ++i) PA_CUDA(ctx, cudaStreamSynchronize(ctx->lanes[i].stream));
  std::vector<unsigned char> r1ok(cmax * m), r2ok(cmax * m);
  PA_CUDA(ctx, cudaMemcpyAsync(r1ok.data(), d_r1ok, cmax * m, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaMemcpyAsync(r2ok.data(), d_r2ok, cmax * m, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t step = 0; step < steps_run; ++step) {
    const std::vector<u32> &act = acts[step], &g1 = g1s_pos[step], &g2 = g2s_pos[step];
    const size_t n1 = g1.size(), o_proof = align_up(n1 * 672, 256);
    for (size_t p = 0; p < act.size(); ++p) {
      size_t o = step * m + act[p];
      okv[auc[act[p]]] &= r1ok[step * m + p] & r2ok[step * m + p];
      if (job->out_r1_ok) job->out_r1_ok[o] = r1ok[step * m + p];
      if (job->out_r2_ok) job->out_r2_ok[o] = r2ok[step * m + p];
      if (want_r1) memcpy(job->out_r1 + o * 320, h_stage + off_r1 + (step * m + p) * 320, 320);
      if (want_b) memcpy(job->out_r2_b + o * 64, h_stage + off_b + (step * m + p) * 64, 64);
    }
    if (want_proof) {
      const unsigned char *base = h_stage + off_pf + step * m * 1344;
      for (size_t q = 0; q < n1; ++q) memcpy(job->out_r2_proof + (step * m + act[g1[q]]) * 1344, base + 672 * q, 672);
      for (size_t q = 0; q < g2.size(); ++q) memcpy(job->out_r2_proof + (step * m + act[g2[q]]) * 1344, base + o_proof + 1344 * q, 1344);
    }
  }
  for (size_t a = 0; a < A; ++a) {
    if (job->max_bid) job->max_bid[a] = maxbid[a];
    if (job->ok) job->ok[a] = okv[a];
  }
  if (p2p) {  // the verdict of all ranks
    PA_LAUNCH(ctx, PA_K_VERDICT, (k_xchg_final<<<1, 32, 0, ctx->stream>>>(1u, okv[0] ? 1u : 0u, maxbid[0], xpeers, xworld, xrank, xpar, (xepoch << 8) | 254u, xerr, d_xfinal)));
    u64 fin[3] = {0, 0, 0};
    PA_CUDA(ctx, cudaMemcpyAsync(fin, d_xfinal, sizeof fin, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = xchg_check())) return rc;
    if (job->ok_all) *job->ok_all = fin[1] != 0 ? 1 : 0;
  }
  return PA_OK;
}
