// Whole-auction runner for the CCS22 protocol: every party of every auction of a batch advances
// in lock step; the curve work of each phase is one batched launch over all auctions.  This is
// what the reference's main (CCS22/main.cpp:16-130) does one party and one EC_POINT_mul at a time.
//
//   setup      Bidder::setupInner / Evaluator::setupInner     CCS22/bidder.cpp:48-89, evaluator.cpp:22-63
//   per step   BESEncode -> OTReceive1 -> OTSend -> OTReceive2 -> checkIfEnterDeciderRound
//                                                             bidder.cpp:118-212, evaluator.cpp:78-156
//
// Party i of auction a draws from PA stream (seed, (auction_id << 32) | i), the bulletin board of
// auction a (g1, h) from (seed, (auction_id << 32) | 0xFFFFFFFF), in the reference's draw order.
// Layout of a party's secret block (scalars): bidder x[c] r[c] s[c] t[c]; evaluator x[c] r[c]
// beta[c][n-1] — exactly the array the setup hash runs over (bidder.cpp:74-77, evaluator.cpp:43-50).
#pragma once

// ---- runner kernels -----------------------------------------------------------------------
// one thread per party: R, then the per-bit draws, written in hash order
__global__ void k_ccs22_setup_rng(u64 seed, const u64 *streams, u64 *ctrs, const u32 *soff, const u32 *cs, const u32 *ns,
                                  const unsigned char *is_eval, unsigned char *sec, unsigned char *R, int m) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  u64 ctr = ctrs[p];
  sc v;
  pa_stream_rand_range(v, seed, streams[p], ctr);  // constructor: R                 CCS22/bidder.cpp:21-27
  st_sc(R + 32 * (size_t)p, v);
  unsigned char *blk = sec + 32 * (size_t)soff[p];
  u32 c = cs[p], nb = ns[p] - 1;
  for (u32 i = 0; i < c; ++i) {
    if (is_eval[p]) {  // x_i, r_i, beta_{i,0..n-2}                                   evaluator.cpp:38-50
      pa_stream_rand_range(v, seed, streams[p], ctr);
      st_sc(blk + 32 * (size_t)i, v);
      pa_stream_rand_range(v, seed, streams[p], ctr);
      st_sc(blk + 32 * (size_t)(c + i), v);
      for (u32 j = 0; j < nb; ++j) {
        pa_stream_rand_range(v, seed, streams[p], ctr);
        st_sc(blk + 32 * ((size_t)2 * c + (size_t)i * nb + j), v);
      }
    } else {  // x_i, r_i, s_i, t_i                                                    bidder.cpp:63-66
      for (u32 k = 0; k < 4; ++k) {
        pa_stream_rand_range(v, seed, streams[p], ctr);
        st_sc(blk + 32 * ((size_t)k * c + i), v);
      }
    }
  }
  ctrs[p] = ctr;
}

// SHA256inSetup over a party's block (variable length per party)
__global__ void k_ccs22_party_hash(const unsigned char *sec, const u32 *soff, const u32 *scnt, unsigned char *out, int m) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  sha256_state s;
  sha256_init(s);
  bool failed = false;
  const unsigned char *blk = sec + 32 * (size_t)soff[p];
  for (u32 j = 0; j < scnt[p] && !failed; ++j) {
    const unsigned char *q = blk + 32 * (size_t)j;
    int lead = 0;
    while (lead < 32 && q[lead] == 0) ++lead;
    if (lead == 32) failed = true;
    for (int b = lead; b < 32; ++b) sha256_put(s, q[b]);
  }
  sc h;
  sc_set_zero(h);
  if (!failed) {
    u32 d[8];
    sha256_final(s, d);
    sc_from_digest(h, d);
  }
  st_sc(out + 32 * (size_t)p, h);
}

// dst[i] = src[idx[i]] for items of 32 or 64 bytes
__global__ void k_gather(unsigned char *dst, const unsigned char *src, const u32 *idx, int item, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 *s = reinterpret_cast<const uint4 *>(src + (size_t)item * idx[i]);
  uint4 *d = reinterpret_cast<uint4 *>(dst + (size_t)item * i);
  for (int k = 0; k < item / 16; ++k) d[k] = s[k];
}
// dst[dst_idx[i]] = src[src_idx[i]] (src_idx == NULL: i) for 64-byte items
__global__ void k_scatter64(unsigned char *dst, const u32 *dst_idx, const unsigned char *src, const u32 *src_idx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 *s = reinterpret_cast<const uint4 *>(src + 64 * (size_t)(src_idx ? src_idx[i] : i));
  uint4 *d = reinterpret_cast<uint4 *>(dst + 64 * (size_t)dst_idx[i]);
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}
// big-endian scalar from a small integer array (bids, 0/1 flags)
__global__ void k_scalar_from_u64(unsigned char *dst, const u64 *v, const u32 *idx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 x = v[idx ? idx[i] : i];
  unsigned char *o = dst + 32 * (size_t)i;
  for (int k = 0; k < 24; ++k) o[k] = 0;
  for (int k = 0; k < 8; ++k) o[31 - k] = (unsigned char)(x >> (8 * k));
}

// BESEncode: d = inRace && bit; B = Y^x (d = 0) or g^r (d = 1)                CCS22/bidder.cpp:118-147
__global__ void __launch_bounds__(PA_BLOCK)
k_ccs22_bes(const u32 *act, const unsigned char *bits, const u32 *boff, int step, const unsigned char *inrace,
            const unsigned char *Y, const unsigned char *sec, const u32 *soff, const u32 *cs, const u32 *__restrict__ comb,
            u64 *dflag, u32 *jout, int n) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  u32 p = act[q];
  int d = inrace[p] && bits[boff[p] + step];
  dflag[p] = (u64)d;
  const unsigned char *blk = sec + 32 * (size_t)soff[p];
  jac r;
  sc k;
  if (d) {
    ld_sc(k, blk + 32 * ((size_t)cs[p] + step));  // r_step
    fixed_base_mul(r, k, comb);
  } else {
    jac P;
    ld_point_jac(P, Y + 64 * (size_t)q);
    ld_sc(k, blk + 32 * (size_t)step);  // x_step
    var_base_mul(r, P, k);
  }
  st_jac(jout + 24 * (size_t)q, r);
}

// after OTReceive2: the published d of every active auction                 bidder.cpp:200-212, evaluator.cpp:117-156
//   newd[a] = d_e ? 1 : !isinf[a];  on newd: every party with d == 0 leaves the race
__global__ void k_ccs22_update(const u32 *act, const u32 *pseg, const u32 *aeval, const u64 *dflag, const int *isinf,
                               unsigned char *inrace, int *newd, int n) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  u32 p = act[q], seg = pseg[q];
  int de = (int)dflag[aeval[seg]];
  int nd = de ? 1 : (isinf[seg] ? 0 : 1);
  if (p == aeval[seg]) newd[seg] = nd;
  if (nd && dflag[p] == 0) inrace[p] = 0;
}

// ---- fused oblivious-transfer messages for the per-party class API --------------------------------------
// A party's OT message is 3-6 scalar multiplications that the reference runs one after the other.
// Here each of them gets its own WARP (role = warp index, 32 items per block), so a single
// message costs the latency of its slowest multiplication, not the sum; the warps meet in shared
// memory for the final additions.  (Roles must be warps, not lanes: lanes of one warp that take
// different code paths are serialised.)
//
// OTSend (CCS22/bidder.cpp:155-198): M1 = g^m, z = g^s h^t, C0 = G^s H^t B,
//   C1 = (G/g1)^s (H/T2)^t M1 = C0 / B / (g1^s T2^t) * M1
__global__ void __launch_bounds__(128)
k_ccs22_ot_send(const unsigned char *r1, const unsigned char *params, const unsigned char *B, const unsigned char *st,
                const unsigned char *mm, const u32 *__restrict__ comb, u32 *jout, int n) {
  __shared__ __align__(16) u32 res[4][32][24];
  int lane = threadIdx.x & 31, role = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
  bool live = i < n;
  if (live) {
    const unsigned char *rec = r1 + 192 * (size_t)i, *pp = params + 128 * (size_t)i;
    sc s, t;
    ld_sc(s, st + 64 * (size_t)i);
    ld_sc(t, st + 64 * (size_t)i + 32);
    jac r, P, Q;
    if (role == 0) {  // M1 = g^m
      sc m;
      ld_sc(m, mm + 32 * (size_t)i);
      fixed_base_mul(r, m, comb);
    } else if (role == 1) {  // z = g^s h^t
      jac v;
      ld_point_jac(P, pp + 64);
      fixed_base_mul(r, s, comb);
      var_base_mul(v, P, t);
      jac_add(r, r, v);
    } else if (role == 2) {  // G^s H^t
      ld_point_jac(P, rec + 64);
      ld_point_jac(Q, rec + 128);
      strauss<2>(r, P, s, Q, t);
    } else {  // g1^s T2^t
      ld_point_jac(P, pp);
      ld_point_jac(Q, rec);
      strauss<2>(r, P, s, Q, t);
    }
    st_jac(res[role][lane], r);
  }
  __syncthreads();
  if (!live) return;
  if (role == 0) {  // C0 = G^s H^t * B
    jac a;
    aff b;
    ld_jac(a, res[2][lane]);
    ld_aff(b, B + 64 * (size_t)i);
    jac_madd(a, a, b);
    st_jac(jout + 24 * ((size_t)i * 3 + 1), a);
  } else if (role == 1) {  // C1 = G^s H^t / (g1^s T2^t) * M1
    jac a, u, m1;
    ld_jac(a, res[2][lane]);
    ld_jac(u, res[3][lane]);
    ld_jac(m1, res[0][lane]);
    jac_neg(u, u);
    jac_add(a, a, u);
    jac_add(a, a, m1);
    st_jac(jout + 24 * ((size_t)i * 3 + 2), a);
  } else if (role == 2) {  // z
    jac z;
    ld_jac(z, res[1][lane]);
    st_jac(jout + 24 * ((size_t)i * 3), z);
  }
}

// OTReceive1 (CCS22/evaluator.cpp:91-111): T2 = g^k, G = g^beta g1^alpha, H = T2^alpha h^beta = g^(alpha k) h^beta
__global__ void __launch_bounds__(96)
k_ccs22_ot_recv1(const unsigned char *k, const unsigned char *beta, const unsigned char *alpha, const unsigned char *params,
                 const u32 *__restrict__ comb, u32 *jout, int n) {
  int lane = threadIdx.x & 31, role = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
  if (i >= n) return;
  const unsigned char *pp = params + 128 * (size_t)i;
  sc sk, sb, sa;
  ld_sc(sk, k + 32 * (size_t)i);
  ld_sc(sb, beta + 32 * (size_t)i);
  ld_sc(sa, alpha + 32 * (size_t)i);
  jac r, v, P;
  if (role == 0) {
    fixed_base_mul(r, sk, comb);
  } else if (role == 1) {
    ld_point_jac(P, pp);
    fixed_base_mul(r, sb, comb);
    var_base_mul(v, P, sa);
    jac_add(r, r, v);
  } else {
    sc ak;
    sc_mul(ak, sa, sk);
    ld_point_jac(P, pp + 64);
    fixed_base_mul(r, ak, comb);
    var_base_mul(v, P, sb);
    jac_add(r, r, v);
  }
  st_jac(jout + 24 * ((size_t)i * 3 + role), r);
}

// BESEncode (CCS22/bidder.cpp:118-147): B = Y_id^x when the party does not veto (d = 0), g^r when it does
__global__ void __launch_bounds__(PA_BLOCK, PA_OP_MINBLOCKS)
k_ccs22_bes_encode(const unsigned char *Y, const u64 *ids, const unsigned char *d, const unsigned char *x, const unsigned char *r,
                   const u32 *__restrict__ comb, u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac res;
  sc k;
  if (d[i]) {
    ld_sc(k, r + 32 * (size_t)i);
    fixed_base_mul(res, k, comb);
  } else {
    jac P;
    ld_point_jac(P, Y + 64 * (size_t)ids[i]);
    ld_sc(k, x + 32 * (size_t)i);
    var_base_mul(res, P, k);
  }
  st_jac(jout + 24 * (size_t)i, res);
}

// Com = g^bid * g1^H + h^R (CCS22/bidder.cpp:80-88)
__global__ void __launch_bounds__(PA_BLOCK, PA_OP_MINBLOCKS)
k_ccs22_commit(const unsigned char *bid, const unsigned char *H, const unsigned char *R, const unsigned char *params,
               const u32 *__restrict__ comb, u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac g1, h, res, f;
  sc a, b;
  ld_point_jac(g1, params + 128 * (size_t)i);
  ld_point_jac(h, params + 128 * (size_t)i + 64);
  ld_sc(a, H + 32 * (size_t)i);
  ld_sc(b, R + 32 * (size_t)i);
  strauss<2>(res, g1, a, h, b);
  ld_sc(a, bid + 32 * (size_t)i);
  fixed_base_mul(f, a, comb);
  jac_add(res, res, f);
  st_jac(jout + 24 * (size_t)i, res);
}

// OTReceive2 (CCS22/evaluator.cpp:130-140): M0_j = C0_j - beta_j * z_j
__global__ void __launch_bounds__(PA_BLOCK, PA_OP_MINBLOCKS)
k_ccs22_recv2_terms(const unsigned char *ots, const unsigned char *beta, u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac z, res;
  aff c0;
  sc k;
  ld_point_jac(z, ots + 192 * (size_t)i);
  ld_sc(k, beta + 32 * (size_t)i);
  var_base_mul(res, z, k);
  jac_neg(res, res);
  ld_aff(c0, ots + 192 * (size_t)i + 64);
  jac_madd(res, res, c0);
  st_jac(jout + 24 * (size_t)i, res);
}

// ---- one auction, phase-major ---------------------------------------------------------------------------
// Step by step, one CCS22 auction is ~25 dependent launches of a few hundred threads per step (6.5 ms).
// Nothing but the race flags depends on earlier steps, and they only SELECT: a party's B is Y^x or g^r, the
// evaluator's (G, H) is (g^beta, h^beta) or (g^beta g1, T2 h^beta), and with W_alpha = G_alpha^s H_alpha^t,
// U = g1^s T2^t:  C0 = W_alpha B,  C1 = W_alpha / U * M1,  and the evaluator's test is
// sum_j (W_0,j / z_j^beta_j) + sum_p B_p == infinity.  All draw counters are arithmetic (n - 1 per step for the
// evaluator, 1 per bidder).  So every candidate of every step is computed in a few large launches, one block
// walks the steps, and one kernel assembles the published records.  Item i = step * (n - 1) + slot,
// party item = step * n + party.
__global__ void __launch_bounds__(PA_BLOCK, PA_OP_MINBLOCKS)
k_c22p_bcand(const unsigned char *Yall, const unsigned char *sec, const u32 *soff, int c, int n, const u32 *__restrict__ comb,
             u32 *jout, int TP) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * TP) return;
  int which = t / TP, i = t % TP, s = i / n, p = i % n;
  const unsigned char *blk = sec + 32 * (size_t)soff[p];
  jac r;
  sc k;
  if (which) {
    ld_sc(k, blk + 32 * ((size_t)c + s));  // r_step
    fixed_base_mul(r, k, comb);
  } else {
    jac P;
    ld_point_jac(P, Yall + 64 * (size_t)i);
    ld_sc(k, blk + 32 * (size_t)s);  // x_step
    var_base_mul(r, P, k);
  }
  st_jac(jout + 24 * ((size_t)i * 2 + which), r);
}
// level 1, op-major: T2 = g^k, Gb = g^beta, Hb = h^beta, z = g^s h^t, M1 = g^m
__global__ void __launch_bounds__(PA_BLOCK, PA_OP_MINBLOCKS)
k_c22p_level1(const unsigned char *kk, const unsigned char *beta, const unsigned char *ssc, const unsigned char *tsc,
              const unsigned char *mm, const unsigned char *h, const u32 *__restrict__ comb, u32 *jout, int T) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 5 * T) return;
  int op = t / T, i = t % T;
  jac r, P, v;
  sc k;
  if (op == 0 || op == 1 || op == 4) {
    ld_sc(k, (op == 0 ? kk : op == 1 ? beta : mm) + 32 * (size_t)i);
    fixed_base_mul(r, k, comb);
  } else if (op == 2) {
    ld_point_jac(P, h);
    ld_sc(k, beta + 32 * (size_t)i);
    var_base_mul(r, P, k);
  } else {
    ld_point_jac(P, h);
    ld_sc(k, tsc + 32 * (size_t)i);
    var_base_mul(v, P, k);
    ld_sc(k, ssc + 32 * (size_t)i);
    fixed_base_mul(r, k, comb);
    jac_add(r, r, v);
  }
  st_jac(jout + 24 * (size_t)t, r);
}
// level 2, op-major, from the affine level-1 arrays L1[op][i]: W0 = Gb^s Hb^t, W1 = (Gb g1)^s (T2 Hb)^t,
// U = g1^s T2^t, bz = z^beta, G1 = Gb g1, H1 = T2 Hb
__global__ void __launch_bounds__(PA_BLOCK, PA_OP_MINBLOCKS)
k_c22p_level2(const unsigned char *L1, const unsigned char *ssc, const unsigned char *tsc, const unsigned char *beta,
              const unsigned char *g1, u32 *jout, int T) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * T) return;
  int op = t / T, i = t % T;
  const unsigned char *T2 = L1 + 64 * (size_t)i, *Gb = L1 + 64 * ((size_t)T + i), *Hb = L1 + 64 * (2 * (size_t)T + i),
                      *z = L1 + 64 * (3 * (size_t)T + i);
  jac r, P, Q;
  sc a, b;
  aff q;
  if (op <= 2) {
    ld_sc(a, ssc + 32 * (size_t)i);
    ld_sc(b, tsc + 32 * (size_t)i);
    if (op == 0) {
      ld_point_jac(P, Gb);
      ld_point_jac(Q, Hb);
    } else if (op == 1) {
      ld_point_jac(P, Gb);
      ld_aff(q, g1);
      jac_madd(P, P, q);
      ld_point_jac(Q, T2);
      ld_aff(q, Hb);
      jac_madd(Q, Q, q);
    } else {
      ld_point_jac(P, g1);
      ld_point_jac(Q, T2);
    }
    strauss<2>(r, P, a, Q, b);
  } else if (op == 3) {
    ld_point_jac(P, z);
    ld_sc(a, beta + 32 * (size_t)i);
    var_base_mul(r, P, a);
  } else if (op == 4) {
    ld_point_jac(r, Gb);
    ld_aff(q, g1);
    jac_madd(r, r, q);
  } else {
    ld_point_jac(r, T2);
    ld_aff(q, Hb);
    jac_madd(r, r, q);
  }
  st_jac(jout + 24 * (size_t)t, r);
}
// the walk through the steps: d_p = inRace_p && bit_p; alpha = d_e; newd = alpha ? 1 : (sum != infinity);
// on newd every party with d = 0 leaves the race       CCS22/bidder.cpp:118-147, 200-212; evaluator.cpp:117-156
__global__ void __launch_bounds__(PA_SCAN_T)
k_c22p_walk(int n, int c, int e, const unsigned char *bits, const u32 *boff, const unsigned char *Bcand, const unsigned char *V,
            unsigned char *inrace, unsigned char *dsel, int *alpha, int *newd) {
  __shared__ __align__(16) u32 part[PA_SCAN_T][24];
  __shared__ int s_de, s_nd;
  const int t = threadIdx.x, nb = n - 1;
  for (int s = 0; s < c; ++s) {
    for (int p = t; p < n; p += PA_SCAN_T) {
      int d = inrace[p] && bits[boff[p] + s];
      dsel[(size_t)s * n + p] = (unsigned char)d;
      if (p == e) s_de = d;
    }
    __syncthreads();
    if (!s_de) {
      jac acc;
      jac_set_inf(acc);
      aff x;
      for (int p = t; p < n; p += PA_SCAN_T) {
        size_t i = (size_t)s * n + p;
        ld_aff(x, Bcand + 64 * (2 * i + dsel[i]));
        jac_madd(acc, acc, x);
      }
      for (int j = t; j < nb; j += PA_SCAN_T) {
        ld_aff(x, V + 64 * ((size_t)s * nb + j));
        jac_madd(acc, acc, x);
      }
      st_jac(part[t], acc);
      __syncthreads();
      for (int d = PA_SCAN_T / 2; d > 0; d >>= 1) {
        if (t < d) {
          jac a, b;
          ld_jac(a, part[t]);
          ld_jac(b, part[t + d]);
          jac_add(a, a, b);
          st_jac(part[t], a);
        }
        __syncthreads();
      }
      if (t == 0) {
        jac a;
        ld_jac(a, part[0]);
        s_nd = jac_is_inf(a) ? 0 : 1;
      }
    } else if (t == 0) {
      s_nd = 1;
    }
    __syncthreads();
    if (t == 0) alpha[s] = s_de, newd[s] = s_nd;
    if (s_nd)
      for (int p = t; p < n; p += PA_SCAN_T)
        if (!dsel[(size_t)s * n + p]) inrace[p] = 0;
    __syncthreads();
  }
}
// the published records of every item: r1 = (T2, G_alpha, H_alpha); ots = (z, C0 = W_alpha B_d, C1 = W_alpha / U * M1)
__global__ void __launch_bounds__(PA_BLOCK, PA_OP_MINBLOCKS)
k_c22p_select(int T, int n, const u32 *slot_party, const int *alpha, const unsigned char *dsel, const unsigned char *L1,
              const unsigned char *L2, const unsigned char *Bcand, unsigned char *r1rec, unsigned char *otsrec, u32 *jout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const int nb = n - 1, s = i / nb, j = i % nb, a = alpha[s], p = (int)slot_party[j];
  const size_t pi = (size_t)s * n + p;
  auto l1 = [&](int op) { return L1 + 64 * ((size_t)op * T + i); };
  auto l2 = [&](int op) { return L2 + 64 * ((size_t)op * T + i); };
  cp64(r1rec + 192 * (size_t)i, l1(0));
  cp64(r1rec + 192 * (size_t)i + 64, a ? l2(4) : l1(1));
  cp64(r1rec + 192 * (size_t)i + 128, a ? l2(5) : l1(2));
  cp64(otsrec + 192 * (size_t)i, l1(3));
  jac W, r;
  aff q;
  ld_point_jac(W, a ? l2(1) : l2(0));
  ld_aff(q, Bcand + 64 * (2 * pi + dsel[pi]));
  jac_madd(r, W, q);
  st_jac(jout + 24 * ((size_t)i * 2), r);
  ld_aff(q, l2(2));
  aff_neg(q, q);
  jac_madd(r, W, q);
  ld_aff(q, l1(4));
  jac_madd(r, r, q);
  st_jac(jout + 24 * ((size_t)i * 2 + 1), r);
}

// ---- host side ------------------------------------------------------------------------------------
namespace {
int dev_point_add(pa_ctx *ctx, const unsigned char *p, const unsigned char *q, unsigned char *out, size_t n, int sub) {
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_POINT_ADD, k_point_add<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(p, q, work_jac(ctx), (int)n, sub));
  return normalize_to(ctx, out, n);
}
int dev_gather(pa_ctx *ctx, unsigned char *dst, const unsigned char *src, const u32 *idx, int item, size_t n) {
  if (n == 0) return PA_OK;
  PA_LAUNCH(ctx, PA_K_ENCODE, k_gather<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(dst, src, idx, item, (int)n));
  return PA_OK;
}
}  // namespace

extern "C" int pa_ccs22_ot_send_dev(pa_ctx *ctx, const uint8_t *r1, const uint8_t *params, const uint8_t *B, const uint8_t *st,
                                    const uint8_t *m, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (r1 && params && B && st && m && out)) && n < (1u << 26));
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, 3 * n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_LINCOMB2, (k_ccs22_ot_send<<<(unsigned)((n + 31) / 32), 128, 0, ctx->stream>>>(r1, params, B, st, m, ctx->d_comb, work_jac(ctx), (int)n)));
  return normalize_to(ctx, out, 3 * n);
}
extern "C" int pa_ccs22_ot_send(pa_ctx *ctx, const uint8_t *r1, const uint8_t *params, const uint8_t *B, const uint8_t *st,
                                const uint8_t *m, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (r1 && params && B && st && m && out)));
  HArg a[] = {{r1, 0, n * 192}, {params, 0, n * 128}, {B, 0, n * 64}, {st, 0, n * 64}, {m, 0, n * 32}, {0, out, n * 192}};
  return staged(ctx, a, 6, [&](unsigned char **d) { return pa_ccs22_ot_send_dev(ctx, d[0], d[1], d[2], d[3], d[4], d[5], n); });
}
extern "C" int pa_ccs22_ot_recv1_dev(pa_ctx *ctx, const uint8_t *k, const uint8_t *beta, const uint8_t *alpha, const uint8_t *params,
                                     uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (k && beta && alpha && params && out)) && n < (1u << 26));
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, 3 * n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_DOUBLE, (k_ccs22_ot_recv1<<<(unsigned)((n + 31) / 32), 96, 0, ctx->stream>>>(k, beta, alpha, params, ctx->d_comb, work_jac(ctx), (int)n)));
  return normalize_to(ctx, out, 3 * n);
}
extern "C" int pa_ccs22_ot_recv1(pa_ctx *ctx, const uint8_t *k, const uint8_t *beta, const uint8_t *alpha, const uint8_t *params,
                                 uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (k && beta && alpha && params && out)));
  HArg a[] = {{k, 0, n * 32}, {beta, 0, n * 32}, {alpha, 0, n * 32}, {params, 0, n * 128}, {0, out, n * 192}};
  return staged(ctx, a, 5, [&](unsigned char **d) { return pa_ccs22_ot_recv1_dev(ctx, d[0], d[1], d[2], d[3], d[4], n); });
}

extern "C" int pa_ccs22_bes_encode_dev(pa_ctx *ctx, const uint8_t *X, size_t n, const uint64_t *ids, const uint8_t *d, const uint8_t *x,
                                       const uint8_t *r, uint8_t *out, size_t m) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && n >= 1 && n < (1u << 26) && m < (1u << 26) && X && (m == 0 || (ids && d && x && r && out)));
  if (m == 0) return PA_OK;
  int rc = ensure(ctx, &ctx->d_aux, &ctx->aux_bytes, n * 64);
  if (rc) return rc;
  if ((rc = pa_y_scan_dev(ctx, X, ctx->d_aux, nullptr, 1, n))) return rc;
  if ((rc = work_reserve(ctx, m))) return rc;
  PA_LAUNCH(ctx, PA_K_VAR, (k_ccs22_bes_encode<<<grid_for(m), PA_BLOCK, 0, ctx->stream>>>(ctx->d_aux, ids, d, x, r, ctx->d_comb, work_jac(ctx), (int)m)));
  return normalize_to(ctx, out, m);
}
extern "C" int pa_ccs22_bes_encode(pa_ctx *ctx, const uint8_t *X, size_t n, const uint64_t *ids, const uint8_t *d, const uint8_t *x,
                                   const uint8_t *r, uint8_t *out, size_t m) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && n >= 1 && X && (m == 0 || (ids && d && x && r && out)));
  for (size_t i = 0; i < m; ++i) PA_ARGCHECK(ctx, ids[i] < n);
  HArg a[] = {{X, 0, n * 64}, {ids, 0, m * 8}, {d, 0, m}, {x, 0, m * 32}, {r, 0, m * 32}, {0, out, m * 64}};
  return staged(ctx, a, 6, [&](unsigned char **p) { return pa_ccs22_bes_encode_dev(ctx, p[0], n, (const uint64_t *)p[1], p[2], p[3], p[4], p[5], m); });
}
extern "C" int pa_ccs22_commit_dev(pa_ctx *ctx, const uint8_t *scalars, size_t k, const uint8_t *bid, const uint8_t *R, const uint8_t *params,
                                   uint8_t *out_H, uint8_t *out_com, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && k >= 1 && n < (1u << 26) && (n == 0 || (scalars && bid && R && params && out_H && out_com)));
  if (n == 0) return PA_OK;
  int rc = pa_ccs22_setup_hash_dev(ctx, scalars, k, out_H, n);
  if (rc) return rc;
  if ((rc = work_reserve(ctx, n))) return rc;
  PA_LAUNCH(ctx, PA_K_LINCOMB2, (k_ccs22_commit<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(bid, out_H, R, params, ctx->d_comb, work_jac(ctx), (int)n)));
  return normalize_to(ctx, out_com, n);
}
extern "C" int pa_ccs22_commit(pa_ctx *ctx, const uint8_t *scalars, size_t k, const uint8_t *bid, const uint8_t *R, const uint8_t *params,
                               uint8_t *out_H, uint8_t *out_com, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && k >= 1 && (n == 0 || (scalars && bid && R && params && out_H && out_com)));
  HArg a[] = {{scalars, 0, n * k * 32}, {bid, 0, n * 32}, {R, 0, n * 32}, {params, 0, n * 128}, {0, out_H, n * 32}, {0, out_com, n * 64}};
  return staged(ctx, a, 6, [&](unsigned char **p) { return pa_ccs22_commit_dev(ctx, p[0], k, p[1], p[2], p[3], p[4], p[5], n); });
}
extern "C" int pa_ccs22_ot_recv2_dev(pa_ctx *ctx, const uint8_t *ots, const uint8_t *beta, const uint8_t *B, size_t n, int32_t *d_is_inf) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && B && d_is_inf && n < (1u << 26) && (n == 0 || (ots && beta)));
  int rc = ensure(ctx, &ctx->d_aux, &ctx->aux_bytes, (n + 1) * 64);
  if (rc) return rc;
  if (n) {
    if ((rc = work_reserve(ctx, n))) return rc;
    PA_LAUNCH(ctx, PA_K_VAR, (k_ccs22_recv2_terms<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(ots, beta, work_jac(ctx), (int)n)));
    if ((rc = normalize_to(ctx, ctx->d_aux, n))) return rc;
  }
  PA_CUDA(ctx, cudaMemcpyAsync(ctx->d_aux + 64 * n, B, 64, cudaMemcpyDeviceToDevice, ctx->stream));
  return pa_point_sum_is_inf_dev(ctx, ctx->d_aux, nullptr, 1, n + 1, d_is_inf);
}
extern "C" int pa_ccs22_ot_recv2(pa_ctx *ctx, const uint8_t *ots, const uint8_t *beta, const uint8_t *B, size_t n, int *is_inf) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && B && is_inf && (n == 0 || (ots && beta)));
  int32_t flag = 0;
  HArg a[] = {{ots, 0, n * 192}, {beta, 0, n * 32}, {B, 0, 64}, {0, &flag, 4}};
  int rc = staged(ctx, a, 4, [&](unsigned char **p) { return pa_ccs22_ot_recv2_dev(ctx, p[0], p[1], p[2], n, (int32_t *)p[3]); });
  *is_inf = flag;
  return rc;
}

extern "C" int pa_ccs22_run(pa_ctx *ctx, const pa_ccs22_job *job) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && job && job->n_auctions >= 1 && job->n && job->c && job->bids && job->evaluator);
  const size_t A = job->n_auctions;
  int rc;
  // ---- host-side index ---------------------------------------------------------------------------
  std::vector<u32> pauc, boff(1, 0), soff(1, 0), aoff(A + 1, 0), cs, ns, scnt, aeval(A), slot_party, slot_auc, slot_j, sl_off(A + 1, 0);
  std::vector<u64> streams, bbstreams(A), bids64;
  std::vector<unsigned char> bits, is_eval;
  size_t cmax = 0;
  for (size_t a = 0; a < A; ++a) {
    u32 n = job->n[a], c = job->c[a], e = job->evaluator[a];
    PA_ARGCHECK(ctx, n >= 1 && c >= 1 && c <= 64 && e < n);
    cmax = c > cmax ? c : cmax;
    u64 aid = job->auction_ids ? job->auction_ids[a] : a;
    bbstreams[a] = (aid << 32) | 0xFFFFFFFFull;
    for (u32 i = 0; i < n; ++i) {
      u32 p = (u32)pauc.size();
      u64 bid = job->bids[p];
      pauc.push_back((u32)a);
      streams.push_back((aid << 32) | i);
      bids64.push_back(bid);
      cs.push_back(c), ns.push_back(n);
      is_eval.push_back(i == e);
      if (i == e) aeval[a] = p;
      u32 cnt = i == e ? (n + 1) * c : 4 * c;
      scnt.push_back(cnt);
      soff.push_back(soff.back() + cnt);
      for (u32 k = 0; k < c; ++k) bits.push_back((unsigned char)((bid >> (c - 1 - k)) & 1));
      boff.push_back((u32)bits.size());
      if (i != e) slot_party.push_back(p), slot_auc.push_back((u32)a), slot_j.push_back((u32)(slot_party.size() - 1 - sl_off[a]));
    }
    aoff[a + 1] = (u32)pauc.size();
    sl_off[a + 1] = (u32)slot_party.size();
  }
  const size_t m = pauc.size(), Mb = bits.size(), NS = soff.back(), ms = slot_party.size();

  // ---- device state ----------------------------------------------------------------------------------
  unsigned char *d_bits, *d_iseval, *d_sec, *d_R, *d_H, *d_params, *d_pub, *d_com, *d_inrace, *d_tmp32a, *d_tmp32b, *d_tmp32c,
      *d_p1, *d_p2, *d_p3, *d_p4, *d_p5, *d_p6, *d_Xs, *d_Y, *d_B, *d_T2, *d_G, *d_Hh, *d_z, *d_C0, *d_C1, *d_sum;
  u32 *d_soff, *d_scnt, *d_cs, *d_ns, *d_boff, *d_idx[10], *d_act, *d_pseg, *d_aeval, *d_segoff;
  u64 *d_streams, *d_ctr, *d_bbstreams, *d_bbctr, *d_bids, *d_dflag;
  int *d_isinf, *d_newd;
  const size_t mx = m > ms ? m : ms, SUM = ms + A;
  // one auction: phase-major unless the caller asks for the step-major schedule
  const bool phased = A == 1 && job->schedule != PA_SEAL_STEP_MAJOR;
  const size_t pn = job->n[0], pc = job->c[0], pnb = pn - 1, T = phased ? pc * pnb : 0, TP = phased ? pc * pn : 0;
  struct {
    u32 *ixp, *ib, *is, *it, *sp, *soffN;
    u64 *st64;
    int *alpha, *newd;
    unsigned char *Xall, *Yall, *Bcand, *kk, *mm, *beta, *ssc, *tsc, *L1, *L2, *V, *dsel, *r1rec, *otsrec;
  } PH{};
  auto carve = [&](DevPool &pool) {
    if (phased) {
      PH.ixp = pool.alloc<u32>(TP); PH.ib = pool.alloc<u32>(T); PH.is = pool.alloc<u32>(T); PH.it = pool.alloc<u32>(T);
      PH.sp = pool.alloc<u32>(pnb); PH.soffN = pool.alloc<u32>(pc + 1); PH.st64 = pool.alloc<u64>(4 * T);
      PH.alpha = pool.alloc<int>(pc); PH.newd = pool.alloc<int>(pc);
      PH.Xall = pool.alloc<unsigned char>(TP * 64); PH.Yall = pool.alloc<unsigned char>(TP * 64); PH.Bcand = pool.alloc<unsigned char>(TP * 128);
      PH.kk = pool.alloc<unsigned char>(T * 32); PH.mm = pool.alloc<unsigned char>(T * 32); PH.beta = pool.alloc<unsigned char>(T * 32);
      PH.ssc = pool.alloc<unsigned char>(T * 32); PH.tsc = pool.alloc<unsigned char>(T * 32);
      PH.L1 = pool.alloc<unsigned char>(5 * T * 64); PH.L2 = pool.alloc<unsigned char>(6 * T * 64); PH.V = pool.alloc<unsigned char>(T * 64);
      PH.dsel = pool.alloc<unsigned char>(TP); PH.r1rec = pool.alloc<unsigned char>(T * 192); PH.otsrec = pool.alloc<unsigned char>(T * 192);
    }
    d_bits = pool.alloc<unsigned char>(Mb); d_iseval = pool.alloc<unsigned char>(m);
    d_sec = pool.alloc<unsigned char>(NS * 32); d_R = pool.alloc<unsigned char>(m * 32); d_H = pool.alloc<unsigned char>(m * 32);
    d_params = pool.alloc<unsigned char>(A * 128); d_pub = pool.alloc<unsigned char>(Mb * 64); d_com = pool.alloc<unsigned char>(m * 64);
    d_inrace = pool.alloc<unsigned char>(m);
    d_tmp32a = pool.alloc<unsigned char>((Mb > mx ? Mb : mx) * 32); d_tmp32b = pool.alloc<unsigned char>(mx * 32); d_tmp32c = pool.alloc<unsigned char>(mx * 32);
    d_p1 = pool.alloc<unsigned char>(mx * 64); d_p2 = pool.alloc<unsigned char>(mx * 64); d_p3 = pool.alloc<unsigned char>(mx * 64);
    d_p4 = pool.alloc<unsigned char>(mx * 64); d_p5 = pool.alloc<unsigned char>(mx * 64); d_p6 = pool.alloc<unsigned char>(mx * 64);
    d_Xs = pool.alloc<unsigned char>(m * 64); d_Y = pool.alloc<unsigned char>(m * 64); d_B = pool.alloc<unsigned char>(m * 64);
    d_T2 = pool.alloc<unsigned char>(ms * 64); d_G = pool.alloc<unsigned char>(ms * 64); d_Hh = pool.alloc<unsigned char>(ms * 64);
    d_z = pool.alloc<unsigned char>(ms * 64); d_C0 = pool.alloc<unsigned char>(ms * 64); d_C1 = pool.alloc<unsigned char>(ms * 64);
    d_sum = pool.alloc<unsigned char>(SUM * 64);
    d_soff = pool.alloc<u32>(m + 1); d_scnt = pool.alloc<u32>(m); d_cs = pool.alloc<u32>(m); d_ns = pool.alloc<u32>(m); d_boff = pool.alloc<u32>(m + 1);
    for (int k = 0; k < 10; ++k) d_idx[k] = pool.alloc<u32>(SUM > Mb ? SUM : Mb);
    d_act = pool.alloc<u32>(m); d_pseg = pool.alloc<u32>(m); d_aeval = pool.alloc<u32>(A); d_segoff = pool.alloc<u32>(A + 1);
    d_streams = pool.alloc<u64>(m); d_ctr = pool.alloc<u64>(m, true); d_bbstreams = pool.alloc<u64>(A); d_bbctr = pool.alloc<u64>(A, true);
    d_bids = pool.alloc<u64>(m); d_dflag = pool.alloc<u64>(m, true);
    d_isinf = pool.alloc<int>(A); d_newd = pool.alloc<int>(A);
  };
  {
    DevPool sizing(ctx, true);
    carve(sizing);
    if ((rc = ensure(ctx, &ctx->d_pool, &ctx->pool_bytes, sizing.off + 4096))) return rc;
    DevPool real(ctx, false);
    carve(real);
  }
  if ((rc = up(ctx, d_bits, bits)) || (rc = up(ctx, d_iseval, is_eval)) || (rc = up(ctx, d_soff, soff)) || (rc = up(ctx, d_scnt, scnt)) ||
      (rc = up(ctx, d_cs, cs)) || (rc = up(ctx, d_ns, ns)) || (rc = up(ctx, d_boff, boff)) || (rc = up(ctx, d_streams, streams)) ||
      (rc = up(ctx, d_bbstreams, bbstreams)) || (rc = up(ctx, d_bids, bids64)))
    return rc;
  PA_CUDA(ctx, cudaMemsetAsync(d_inrace, 1, m, ctx->stream));
  auto upidx = [&](int k, const std::vector<u32> &v) { return up(ctx, d_idx[k], v); };
  auto d2h = [&](void *h, const void *d, size_t bytes) -> int {
    if (h && bytes) PA_CUDA(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PA_OK;
  };

  // ================= setup =====================================================================
  // public parameters g1 = g^rand256, h = g^rand256                          CCS22/bulletinBoard.cpp:28-51
  PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(A), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_bbstreams, d_bbctr, nullptr, 2, d_tmp32a, (int)A, 1)));
  if ((rc = pa_fixed_base_mul_dev(ctx, d_tmp32a, d_params, 2 * A))) return rc;
  // every party: R, per-bit secrets, X_i = g^x_i, H, Com = g^bid g1^H + h^R      bidder.cpp:48-89, evaluator.cpp:22-63
  PA_LAUNCH(ctx, PA_K_RNG, (k_ccs22_setup_rng<<<grid_for(m), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_streams, d_ctr, d_soff, d_cs, d_ns, d_iseval, d_sec, d_R, (int)m)));
  {
    std::vector<u32> ix(Mb), ig1(m), ih(m);
    for (size_t p = 0; p < m; ++p) {
      for (u32 k = boff[p]; k < boff[p + 1]; ++k) ix[k] = soff[p] + (k - boff[p]);  // x_i of the party's block
      ig1[p] = 2 * pauc[p], ih[p] = 2 * pauc[p] + 1;
    }
    if ((rc = upidx(0, ix)) || (rc = upidx(1, ig1)) || (rc = upidx(2, ih))) return rc;
    if ((rc = dev_gather(ctx, d_tmp32a, d_sec, d_idx[0], 32, Mb))) return rc;
    if ((rc = pa_fixed_base_mul_dev(ctx, d_tmp32a, d_pub, Mb))) return rc;
    PA_LAUNCH(ctx, PA_K_CHALLENGE, (k_ccs22_party_hash<<<grid_for(m), PA_BLOCK, 0, ctx->stream>>>(d_sec, d_soff, d_scnt, d_H, (int)m)));
    PA_LAUNCH(ctx, PA_K_ENCODE, (k_scalar_from_u64<<<grid_for(m), PA_BLOCK, 0, ctx->stream>>>(d_tmp32b, d_bids, nullptr, (int)m)));
    if ((rc = dev_gather(ctx, d_p1, d_params, d_idx[1], 64, m)) || (rc = dev_gather(ctx, d_p2, d_params, d_idx[2], 64, m))) return rc;
    if ((rc = pa_double_mul_dev(ctx, d_tmp32b, d_p1, d_H, d_p3, m))) return rc;   // g^bid * g1^H
    if ((rc = pa_var_base_mul_dev(ctx, d_p2, d_R, d_p4, m))) return rc;           // h^R
    if ((rc = dev_point_add(ctx, d_p3, d_p4, d_com, m, 0))) return rc;
  }
  if ((rc = d2h(job->out_params, d_params, A * 128)) || (rc = d2h(job->out_com, d_com, m * 64)) || (rc = d2h(job->out_pub, d_pub, Mb * 64))) return rc;
  std::vector<u64> hctr(m);  // stream counters after setup (they include any rejected range draws)
  PA_CUDA(ctx, cudaMemcpyAsync(hctr.data(), d_ctr, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

  // ================= one auction, phase-major (see k_c22p_*) ===================================================
  if (phased) {
    const u32 e = aeval[0];
    {
      std::vector<u32> ixp(TP), ib(T), is(T), it(T), sp(slot_party), soffN(pc + 1);
      std::vector<u64> st64(4 * T);
      for (size_t s = 0; s < pc; ++s) {
        for (size_t p = 0; p < pn; ++p) ixp[s * pn + p] = boff[p] + (u32)s;
        for (size_t j = 0; j < pnb; ++j) {
          const size_t i = s * pnb + j, p = slot_party[j];
          ib[i] = soff[e] + 2 * (u32)pc + (u32)(s * pnb + j);
          is[i] = soff[p] + 2 * (u32)pc + (u32)s;
          it[i] = soff[p] + 3 * (u32)pc + (u32)s;
          st64[i] = streams[e], st64[T + i] = hctr[e] + s * pnb + j;      // k of OTReceive1: n - 1 draws per step
          st64[2 * T + i] = streams[p], st64[3 * T + i] = hctr[p] + s;    // m of OTSend: one draw per step
        }
      }
      for (size_t k = 0; k <= pc; ++k) soffN[k] = (u32)(k * pn);
      if ((rc = up(ctx, PH.ixp, ixp)) || (rc = up(ctx, PH.ib, ib)) || (rc = up(ctx, PH.is, is)) || (rc = up(ctx, PH.it, it)) ||
          (rc = up(ctx, PH.sp, sp)) || (rc = up(ctx, PH.soffN, soffN)) || (rc = up(ctx, PH.st64, st64)))
        return rc;
      PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // every party's Y and both B candidates of every step: on a side lane, beside the OT candidates below
    if ((rc = lanes_init(ctx))) return rc;
    struct LaneDrain {
      pa_ctx *c;
      ~LaneDrain() { cudaStreamSynchronize(c->lanes[0].stream); }
    } drain{ctx};
    PA_CUDA(ctx, cudaEventRecord(ctx->lane_ev[0], ctx->stream));
    {
      LaneScope ls(ctx, &ctx->lanes[0]);
      PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_ev[0], 0));
      if ((rc = dev_gather(ctx, PH.Xall, d_pub, PH.ixp, 64, TP))) return rc;
      if ((rc = work_reserve(ctx, 2 * TP))) return rc;
      PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<(unsigned)pc, PA_SCAN_T, 0, ctx->stream>>>(PH.Xall, 64, PH.soffN, (int)TP, work_jac(ctx))));
      if ((rc = normalize_to(ctx, PH.Yall, TP))) return rc;
      PA_LAUNCH(ctx, PA_K_VAR, (k_c22p_bcand<<<grid_for(2 * TP), PA_BLOCK, 0, ctx->stream>>>(PH.Yall, d_sec, d_soff, (int)pc, (int)pn, ctx->d_comb, work_jac(ctx), (int)TP)));
      if ((rc = normalize_to(ctx, PH.Bcand, 2 * TP))) return rc;
      PA_CUDA(ctx, cudaEventRecord(ctx->lane_ev[1], ctx->stream));
    }
    if ((rc = work_reserve(ctx, 6 * T + 2))) return rc;
    if (T) {
      PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(T), PA_BLOCK, 0, ctx->stream>>>(job->seed, PH.st64, PH.st64 + T, nullptr, 1, PH.kk, (int)T, 1)));
      PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(T), PA_BLOCK, 0, ctx->stream>>>(job->seed, PH.st64 + 2 * T, PH.st64 + 3 * T, nullptr, 1, PH.mm, (int)T, 1)));
      if ((rc = dev_gather(ctx, PH.beta, d_sec, PH.ib, 32, T)) || (rc = dev_gather(ctx, PH.ssc, d_sec, PH.is, 32, T)) ||
          (rc = dev_gather(ctx, PH.tsc, d_sec, PH.it, 32, T)))
        return rc;
      PA_LAUNCH(ctx, PA_K_DOUBLE, (k_c22p_level1<<<grid_for(5 * T), PA_BLOCK, 0, ctx->stream>>>(PH.kk, PH.beta, PH.ssc, PH.tsc, PH.mm, d_params + 64, ctx->d_comb, work_jac(ctx), (int)T)));
      if ((rc = normalize_to(ctx, PH.L1, 5 * T))) return rc;
      PA_LAUNCH(ctx, PA_K_LINCOMB2, (k_c22p_level2<<<grid_for(6 * T), PA_BLOCK, 0, ctx->stream>>>(PH.L1, PH.ssc, PH.tsc, PH.beta, d_params, work_jac(ctx), (int)T)));
      if ((rc = normalize_to(ctx, PH.L2, 6 * T))) return rc;
      if ((rc = dev_point_add(ctx, PH.L2, PH.L2 + 64 * 3 * T, PH.V, T, 1))) return rc;   // V = W0 / z^beta
    }
    PA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_ev[1], 0));  // the B candidates (side lane)
    PA_LAUNCH(ctx, PA_K_SUMINF, (k_c22p_walk<<<1, PA_SCAN_T, 0, ctx->stream>>>((int)pn, (int)pc, (int)e, d_bits, d_boff, PH.Bcand, PH.V, d_inrace, PH.dsel, PH.alpha, PH.newd)));
    if (T) {
      PA_LAUNCH(ctx, PA_K_ENCODE, (k_c22p_select<<<grid_for(T), PA_BLOCK, 0, ctx->stream>>>((int)T, (int)pn, PH.sp, PH.alpha, PH.dsel, PH.L1, PH.L2, PH.Bcand, PH.r1rec, PH.otsrec, work_jac(ctx))));
      if ((rc = normalize_to(ctx, PH.otsrec + 64, 2 * T, 2, 192))) return rc;
    }
    std::vector<int> newd(pc);
    PA_CUDA(ctx, cudaMemcpyAsync(newd.data(), PH.newd, pc * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = d2h(job->out_r1, PH.r1rec, T * 192)) || (rc = d2h(job->out_ots, PH.otsrec, T * 192))) return rc;
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    u64 mb = 0;
    for (size_t s = 0; s < pc; ++s) {
      if (job->out_d) job->out_d[s] = (uint8_t)newd[s];
      if (newd[s]) mb |= (u64)1 << (pc - s - 1);  // bidder.cpp:210, evaluator.cpp:122, 147
    }
    for (size_t p = 0; p < m; ++p)
      if (job->max_bid) job->max_bid[p] = mb;
    return PA_OK;
  }

  std::vector<u64> maxbid(m, 0);
  // ================= computation phase ============================================================
  for (size_t step = 0; step < cmax; ++step) {
    std::vector<u32> act, pseg, segoff(1, 0), aev, actauc, sact, sseg, sumoff(1, 0);
    for (size_t a = 0; a < A; ++a) {
      if (job->c[a] <= step) continue;
      for (u32 p = aoff[a]; p < aoff[a + 1]; ++p) act.push_back(p), pseg.push_back((u32)actauc.size());
      for (u32 q = sl_off[a]; q < sl_off[a + 1]; ++q) sact.push_back(q), sseg.push_back((u32)actauc.size());
      aev.push_back(aeval[a]);
      actauc.push_back((u32)a);
      segoff.push_back((u32)act.size());
      sumoff.push_back((u32)(sact.size() + actauc.size()));
    }
    const size_t ma = act.size(), na = actauc.size(), sa = sact.size();
    if (ma == 0) break;
    // position of every active party inside the active list (for B of a slot / of an evaluator)
    std::vector<u32> pos_of(m, 0);
    for (size_t q = 0; q < ma; ++q) pos_of[act[q]] = (u32)q;
    if ((rc = up(ctx, d_act, act)) || (rc = up(ctx, d_pseg, pseg)) || (rc = up(ctx, d_segoff, segoff)) || (rc = up(ctx, d_aeval, aev))) return rc;

    // ---- BESEncode for every party ------------------------------------------------------------
    {
      std::vector<u32> ix(ma);
      for (size_t q = 0; q < ma; ++q) ix[q] = boff[act[q]] + (u32)step;  // public key of this step
      if ((rc = upidx(0, ix))) return rc;
      if ((rc = dev_gather(ctx, d_Xs, d_pub, d_idx[0], 64, ma))) return rc;
      if ((rc = work_reserve(ctx, ma))) return rc;
      PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<(unsigned)na, PA_SCAN_T, 0, ctx->stream>>>(d_Xs, 64, d_segoff, (int)ma, work_jac(ctx))));
      if ((rc = normalize_to(ctx, d_Y, ma))) return rc;
      PA_LAUNCH(ctx, PA_K_VAR, (k_ccs22_bes<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(d_act, d_bits, d_boff, (int)step, d_inrace, d_Y, d_sec, d_soff, d_cs, ctx->d_comb, d_dflag, work_jac(ctx), (int)ma)));
      if ((rc = normalize_to(ctx, d_B, ma))) return rc;
    }
    if (sa) {
      // per slot (= one non-evaluator bidder of an active auction): index arrays
      std::vector<u32> i_beta(sa), i_s(sa), i_t(sa), i_g1(sa), i_h(sa), i_alpha(sa), i_B(sa), i_bidder(sa), i_eval(sa);
      for (size_t k = 0; k < sa; ++k) {
        u32 q = sact[k], p = slot_party[q], a = slot_auc[q], e = aeval[a], c = job->c[a], nb = job->n[a] - 1;
        i_beta[k] = soff[e] + 2 * c + (u32)step * nb + slot_j[q];
        i_s[k] = soff[p] + 2 * c + (u32)step;
        i_t[k] = soff[p] + 3 * c + (u32)step;
        i_g1[k] = 2 * a, i_h[k] = 2 * a + 1;
        i_alpha[k] = e;  // d flag of the evaluator
        i_B[k] = pos_of[p];
        i_bidder[k] = p, i_eval[k] = e;
      }
      if ((rc = upidx(0, i_beta)) || (rc = upidx(1, i_s)) || (rc = upidx(2, i_t)) || (rc = upidx(3, i_g1)) || (rc = upidx(4, i_h)) ||
          (rc = upidx(5, i_alpha)) || (rc = upidx(6, i_B)) || (rc = upidx(7, i_bidder)) || (rc = upidx(8, i_eval)))
        return rc;
      // ---- OTReceive1: k <- rand256 (evaluator's stream, one per slot, in slot order), T2 = g^k,
      //      G = g^beta * g1^alpha, H = T2^alpha + h^beta                               evaluator.cpp:91-111
      // rand256 draws are never rejected, so the stream counters are tracked on the host: slot j of
      // evaluator e draws at counter ctr[e] + j; then every bidder draws once (M1 of OTSend)
      {
        std::vector<u64> sstream(sa), sctr(sa), bstream(sa), bctr(sa);
        for (size_t k = 0; k < sa; ++k) {
          u32 q = sact[k], e = aeval[slot_auc[q]];
          sstream[k] = streams[e], sctr[k] = hctr[e] + slot_j[q];
        }
        for (size_t k = 0; k < na; ++k) hctr[aev[k]] += job->n[actauc[k]] - 1;
        for (size_t k = 0; k < sa; ++k) {
          u32 p = slot_party[sact[k]];
          bstream[k] = streams[p], bctr[k] = hctr[p]++;
        }
        u64 *d_s64 = (u64 *)d_p6, *d_c64 = d_s64 + sa;  // scratch
        PA_CUDA(ctx, cudaMemcpyAsync(d_s64, sstream.data(), sa * 8, cudaMemcpyHostToDevice, ctx->stream));
        PA_CUDA(ctx, cudaMemcpyAsync(d_c64, sctr.data(), sa * 8, cudaMemcpyHostToDevice, ctx->stream));
        PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(sa), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_s64, d_c64, nullptr, 1, d_tmp32a, (int)sa, 1)));
        PA_CUDA(ctx, cudaMemcpyAsync(d_s64, bstream.data(), sa * 8, cudaMemcpyHostToDevice, ctx->stream));
        PA_CUDA(ctx, cudaMemcpyAsync(d_c64, bctr.data(), sa * 8, cudaMemcpyHostToDevice, ctx->stream));
        PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(sa), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_s64, d_c64, nullptr, 1, d_tmp32c, (int)sa, 1)));
      }
      if ((rc = pa_fixed_base_mul_dev(ctx, d_tmp32a, d_T2, sa))) return rc;                       // T2 = g^k
      if ((rc = dev_gather(ctx, d_tmp32a, d_sec, d_idx[0], 32, sa))) return rc;                   // beta
      PA_LAUNCH(ctx, PA_K_ENCODE, (k_scalar_from_u64<<<grid_for(sa), PA_BLOCK, 0, ctx->stream>>>(d_tmp32b, d_dflag, d_idx[5], (int)sa)));  // alpha = d_e
      if ((rc = dev_gather(ctx, d_p1, d_params, d_idx[3], 64, sa)) || (rc = dev_gather(ctx, d_p2, d_params, d_idx[4], 64, sa))) return rc;  // g1, h
      if ((rc = pa_double_mul_dev(ctx, d_tmp32a, d_p1, d_tmp32b, d_G, sa))) return rc;            // G = g^beta g1^alpha
      if ((rc = pa_lincomb2_dev(ctx, d_T2, d_tmp32b, d_p2, d_tmp32a, d_Hh, sa))) return rc;       // H = T2^alpha + h^beta
      // ---- OTSend: M1 = g^rand256, z = g^s h^t, C0 = G^s + H^t + B, C1 = (G - g1)^s + (H - T2)^t + M1   bidder.cpp:155-198
      if ((rc = pa_fixed_base_mul_dev(ctx, d_tmp32c, d_p3, sa))) return rc;                       // M1
      if ((rc = dev_gather(ctx, d_tmp32b, d_sec, d_idx[1], 32, sa)) || (rc = dev_gather(ctx, d_tmp32c, d_sec, d_idx[2], 32, sa))) return rc;  // s, t
      if ((rc = pa_double_mul_dev(ctx, d_tmp32b, d_p2, d_tmp32c, d_z, sa))) return rc;            // z
      if ((rc = pa_lincomb2_dev(ctx, d_G, d_tmp32b, d_Hh, d_tmp32c, d_p4, sa))) return rc;
      if ((rc = dev_gather(ctx, d_p5, d_B, d_idx[6], 64, sa))) return rc;                         // B of the bidder
      if ((rc = dev_point_add(ctx, d_p4, d_p5, d_C0, sa, 0))) return rc;
      if ((rc = dev_point_add(ctx, d_G, d_p1, d_p4, sa, 1)) || (rc = dev_point_add(ctx, d_Hh, d_T2, d_p5, sa, 1))) return rc;
      if ((rc = pa_lincomb2_dev(ctx, d_p4, d_tmp32b, d_p5, d_tmp32c, d_p6, sa))) return rc;
      if ((rc = dev_point_add(ctx, d_p6, d_p3, d_C1, sa, 0))) return rc;
      // ---- OTReceive2: sum_j (C0_j - beta_j z_j) + B_e                                evaluator.cpp:131-143
      if ((rc = pa_var_base_mul_dev(ctx, d_z, d_tmp32a, d_p4, sa))) return rc;                    // beta z
      if ((rc = dev_point_add(ctx, d_C0, d_p4, d_p5, sa, 1))) return rc;                          // M0
    }
    {
      // per active auction: its M0's (slot order) followed by the evaluator's own B, then one segmented sum
      std::vector<u32> where_m0(sa), where_b(na), src_b(na);
      size_t k = 0, w = 0;
      for (size_t sgi = 0; sgi < na; ++sgi) {
        while (k < sa && sseg[k] == sgi) where_m0[k++] = (u32)w++;
        src_b[sgi] = pos_of[aev[sgi]];
        where_b[sgi] = (u32)w++;
      }
      if ((rc = upidx(0, where_m0)) || (rc = upidx(1, where_b)) || (rc = upidx(2, src_b)) || (rc = up(ctx, d_segoff, sumoff))) return rc;
      if (sa) PA_LAUNCH(ctx, PA_K_ENCODE, (k_scatter64<<<grid_for(sa), PA_BLOCK, 0, ctx->stream>>>(d_sum, d_idx[0], d_p5, nullptr, (int)sa)));
      PA_LAUNCH(ctx, PA_K_ENCODE, (k_scatter64<<<grid_for(na), PA_BLOCK, 0, ctx->stream>>>(d_sum, d_idx[1], d_B, d_idx[2], (int)na)));
      PA_LAUNCH(ctx, PA_K_SUMINF, (k_point_sum_is_inf<<<(unsigned)na, PA_SCAN_T, 0, ctx->stream>>>(d_sum, 64, d_segoff, (int)(sa + na), d_isinf)));
      PA_LAUNCH(ctx, PA_K_VERDICT, (k_ccs22_update<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(d_act, d_pseg, d_aeval, d_dflag, d_isinf, d_inrace, d_newd, (int)ma)));
    }
    // ---- results of the step -----------------------------------------------------------------------
    std::vector<int> newd(na);
    PA_CUDA(ctx, cudaMemcpyAsync(newd.data(), d_newd, na * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<unsigned char> h_r1, h_ots;
    if (job->out_r1 && sa) {
      h_r1.resize(sa * 192);
      PA_CUDA(ctx, cudaMemcpy2DAsync(h_r1.data(), 192, d_T2, 64, 64, sa, cudaMemcpyDeviceToHost, ctx->stream));
      PA_CUDA(ctx, cudaMemcpy2DAsync(h_r1.data() + 64, 192, d_G, 64, 64, sa, cudaMemcpyDeviceToHost, ctx->stream));
      PA_CUDA(ctx, cudaMemcpy2DAsync(h_r1.data() + 128, 192, d_Hh, 64, 64, sa, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (job->out_ots && sa) {
      h_ots.resize(sa * 192);
      PA_CUDA(ctx, cudaMemcpy2DAsync(h_ots.data(), 192, d_z, 64, 64, sa, cudaMemcpyDeviceToHost, ctx->stream));
      PA_CUDA(ctx, cudaMemcpy2DAsync(h_ots.data() + 64, 192, d_C0, 64, 64, sa, cudaMemcpyDeviceToHost, ctx->stream));
      PA_CUDA(ctx, cudaMemcpy2DAsync(h_ots.data() + 128, 192, d_C1, 64, 64, sa, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < sa; ++k) {
      if (job->out_r1) memcpy(job->out_r1 + (step * ms + sact[k]) * 192, h_r1.data() + 192 * k, 192);
      if (job->out_ots) memcpy(job->out_ots + (step * ms + sact[k]) * 192, h_ots.data() + 192 * k, 192);
    }
    for (size_t s = 0; s < na; ++s) {
      u32 a = actauc[s];
      if (job->out_d) job->out_d[step * A + a] = (uint8_t)newd[s];
      if (newd[s])
        for (u32 p = aoff[a]; p < aoff[a + 1]; ++p) maxbid[p] |= (u64)1 << (job->c[a] - step - 1);  // bidder.cpp:210, evaluator.cpp:122, 147
    }
  }
  for (size_t p = 0; p < m; ++p)
    if (job->max_bid) job->max_bid[p] = maxbid[p];
  return PA_OK;
}
