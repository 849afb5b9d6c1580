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
// Partitioning: independent auctions need no exchange at
// all; ONE auction can be sharded by bidder slice, in which case the X_i of
// round one and the b_i of round two are all-gathered once per step through a
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
// the active list).  Stage 1: (b, X, Y, R, c, A, B), (x, alpha).  Stage 2: (Bi, Xi, Ri, Bj, Xj,
// Rj, Ci, A, B, Yi, Yj), (xi, xj, alpha), bi = encoded bit, bj = prevDecidingBit.
__global__ void k_seal_stmt(int stage, const u32 *g, const u32 *act, const u32 *boff, int step, const unsigned char *b,
                            const unsigned char *r1, const unsigned char *Y, const unsigned char *rnd1,
                            const unsigned char *crec, const unsigned char *rndc, const unsigned char *prevpts,
                            const unsigned char *prevx, const unsigned char *ebit, const unsigned char *prevbit,
                            unsigned char *stmt, unsigned char *sec, unsigned char *bi, unsigned char *bj, int n) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  u32 p = g[q], slot = act[p];
  size_t cs = (size_t)boff[slot] + step;  // commitment slot of this bidder's current bit
  const unsigned char *X = r1 + 320 * (size_t)p, *R = X + 64, *c = crec + 736 * cs;
  if (stage == 1) {
    unsigned char *o = stmt + 448 * (size_t)q;
    cp64(o, b + 64 * (size_t)p); cp64(o + 64, X); cp64(o + 128, Y + 64 * (size_t)p); cp64(o + 192, R);
    cp64(o + 256, c); cp64(o + 320, c + 64); cp64(o + 384, c + 128);
    cp32(sec + 64 * (size_t)q, rnd1 + 128 * (size_t)p);
    cp32(sec + 64 * (size_t)q + 32, rndc + 224 * cs);
    bi[q] = ebit[p];
  } else {
    const unsigned char *pp = prevpts + 256 * (size_t)slot;  // X, R, Y, b at the previous deciding step
    unsigned char *o = stmt + 704 * (size_t)q;
    cp64(o, b + 64 * (size_t)p); cp64(o + 64, X); cp64(o + 128, R);
    cp64(o + 192, pp + 192); cp64(o + 256, pp); cp64(o + 320, pp + 64);
    cp64(o + 384, c); cp64(o + 448, c + 64); cp64(o + 512, c + 128);
    cp64(o + 576, Y + 64 * (size_t)p); cp64(o + 640, pp + 128);
    cp32(sec + 96 * (size_t)q, rnd1 + 128 * (size_t)p);
    cp32(sec + 96 * (size_t)q + 32, prevx + 32 * (size_t)slot);
    cp32(sec + 96 * (size_t)q + 64, rndc + 224 * cs);
    bi[q] = ebit[p];
    bj[q] = prevbit[slot];
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
  cp32(prevx + 32 * (size_t)slot, rnd1 + 128 * (size_t)p);
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

// ---- host side ------------------------------------------------------------------------------------
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

}  // namespace

extern "C" int pa_seal_run(pa_ctx *ctx, const pa_seal_job *job) {
  PA_ARGCHECK(ctx, ctx && job && job->n_auctions >= 1 && job->n && job->c && job->bids);
  const size_t A = job->n_auctions;
  const bool sharded = job->allgather != nullptr;
  PA_ARGCHECK(ctx, !sharded || (A == 1 && job->hi > job->lo && job->hi <= job->n[0] && job->d_send && job->d_recv && job->slice >= job->hi - job->lo));
  const bool verify = job->verify != 0;

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

  unsigned char *d_bits, *d_rndc, *d_crec, *d_cv, *d_junc, *d_prevbit, *d_prevpts, *d_prevx, *d_rnd1, *d_r1, *d_r1v, *d_Y, *d_b,
      *d_ebit, *d_stmt, *d_sec, *d_bi, *d_bj, *d_rnd2, *d_proof, *d_pv, *d_r2v;
  u32 *d_boff, *d_act, *d_pauc, *d_pseg, *d_soff, *d_g, *d_gslot;
  u64 *d_ids, *d_streams, *d_ctr, *d_cid, *d_pid, *d_gid, *d_sstream, *d_sctr;
  int *d_isinf;
  auto carve = [&](DevPool &pool) {
#define PA_ALLOC(var, T, count, zero) var = pool.alloc<T>(count, zero)
  PA_ALLOC(d_bits, unsigned char, Mb, false);
  PA_ALLOC(d_boff, u32, m + 1, false);
  PA_ALLOC(d_ids, u64, m, false);
  PA_ALLOC(d_streams, u64, m, false);
  PA_ALLOC(d_ctr, u64, m, true);
  PA_ALLOC(d_cid, u64, Mb, false);
  PA_ALLOC(d_rndc, unsigned char, Mb * 224, false);
  PA_ALLOC(d_crec, unsigned char, Mb * 736, false);
  PA_ALLOC(d_cv, unsigned char, Mb * 4, false);
  PA_ALLOC(d_junc, unsigned char, A, true);
  PA_ALLOC(d_prevbit, unsigned char, m, false);
  PA_ALLOC(d_prevpts, unsigned char, m * 256, true);
  PA_ALLOC(d_prevx, unsigned char, m * 32, true);
  // per-step (sized for all local bidders)
  PA_ALLOC(d_act, u32, m, false);
  PA_ALLOC(d_pauc, u32, m, false);
  PA_ALLOC(d_pseg, u32, m, false);
  PA_ALLOC(d_pid, u64, m, false);
  PA_ALLOC(d_soff, u32, A + 1, false);
  PA_ALLOC(d_isinf, int, A, false);
  PA_ALLOC(d_rnd1, unsigned char, m * 128, false);
  PA_ALLOC(d_r1, unsigned char, m * 320, false);
  PA_ALLOC(d_r1v, unsigned char, m * 3, false);
  PA_ALLOC(d_Y, unsigned char, (sharded ? (size_t)job->n[0] : m) * 64, false);
  PA_ALLOC(d_b, unsigned char, m * 64, false);
  PA_ALLOC(d_ebit, unsigned char, m, false);
  PA_ALLOC(d_g, u32, 2 * m, false);       // group member positions: stage 1 first, then stage 2
  PA_ALLOC(d_gslot, u32, 2 * m, false);   // their bidder slots (for the draw counters)
  PA_ALLOC(d_gid, u64, 2 * m, false);
  PA_ALLOC(d_stmt, unsigned char, m * 704, false);
  PA_ALLOC(d_sec, unsigned char, m * 96, false);
  PA_ALLOC(d_bi, unsigned char, m, false);
  PA_ALLOC(d_bj, unsigned char, m, false);
  PA_ALLOC(d_rnd2, unsigned char, m * 352, false);
  PA_ALLOC(d_proof, unsigned char, m * 1344, false);
  PA_ALLOC(d_pv, unsigned char, m, false);
  PA_ALLOC(d_r2v, unsigned char, m, false);
  PA_ALLOC(d_sstream, u64, Mb, false);
  PA_ALLOC(d_sctr, u64, Mb, false);
#undef PA_ALLOC
  };
  {
    DevPool sizing(ctx, true);
    carve(sizing);
    int rc0 = ensure(ctx, &ctx->d_pool, &ctx->pool_bytes, sizing.off + 4096);
    if (rc0) return rc0;
    DevPool real(ctx, false);
    carve(real);
  }
  int rc;
  if ((rc = up(ctx, d_bits, bits)) || (rc = up(ctx, d_boff, boff)) || (rc = up(ctx, d_ids, ids)) ||
      (rc = up(ctx, d_streams, streams)) || (rc = up(ctx, d_cid, cid)))
    return rc;
  PA_CUDA(ctx, cudaMemsetAsync(d_prevbit, 1, m, ctx->stream));  // prevDecidingBit(1), SEAL/bidder.cpp:23

  std::vector<unsigned char> junction(A, 0), okv(A, 1);
  std::vector<u64> maxbid(A, 0);
  // strides of the records the proofs live in; LC2 / LR2: the two Schnorr proofs of a record in one batch
  const pa_lay LC{736, 736, 224, 224, 1, 0, 0, 0, 0}, LC2{736, 736, 224, 224, 2, 96, 64, 32, 32};
  const pa_lay LR2{320, 320, 128, 128, 2, 96, 64, 32, 32};
  auto d2h = [&](void *h, const void *d, size_t bytes) -> int {
    if (h && bytes) PA_CUDA(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PA_OK;
  };

  // ================= commit phase ===================================================
  {
    // 7 draws per bit, sequentially per bidder (thread per bidder walks its c bits)
    // draws of bidder s land at rndc + 224 * boff[s]: one launch per distinct c would need
    // compaction; instead draw bit-slot by bit-slot with per-slot (stream, counter = 7 * bit index).
    // The counter arithmetic is exact unless a draw is rejected (probability 2^-128 per draw);
    // rejections are handled by the sequential fallback below.
    std::vector<u64> sstream(Mb), sctr(Mb);
    for (size_t s = 0; s < m; ++s)
      for (u32 k = boff[s]; k < boff[s + 1]; ++k) sstream[k] = streams[s], sctr[k] = 7ull * (k - boff[s]);
    if ((rc = up(ctx, d_sstream, sstream)) || (rc = up(ctx, d_sctr, sctr))) return rc;
    PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(Mb), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_sstream, d_sctr, nullptr, 7, d_rndc, (int)Mb)));
    // every slot must have consumed exactly 7 counters; otherwise redo that bidder sequentially
    std::vector<u64> after(Mb);
    PA_CUDA(ctx, cudaMemcpyAsync(after.data(), d_sctr, Mb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t s = 0; s < m; ++s) {
      bool clean = true;
      for (u32 k = boff[s]; k < boff[s + 1]; ++k) clean &= after[k] == sctr[k] + 7;
      u64 cnt = 7ull * (boff[s + 1] - boff[s]);
      if (!clean) {  // a rejected draw shifted the stream: regenerate this bidder's draws in order
        u64 zero = 0;
        PA_CUDA(ctx, cudaMemcpyAsync(d_ctr + s, &zero, 8, cudaMemcpyHostToDevice, ctx->stream));
        PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<1, 1, 0, ctx->stream>>>(job->seed, d_streams + s, d_ctr + s, nullptr, (int)cnt, d_rndc + 224 * (size_t)boff[s], 1)));
      } else {
        PA_CUDA(ctx, cudaMemcpyAsync(d_ctr + s, &cnt, 8, cudaMemcpyHostToDevice, ctx->stream));
      }
    }
    if ((rc = work_reserve(ctx, 3 * Mb))) return rc;
    PA_LAUNCH(ctx, PA_K_COMMIT, (k_seal_commit_points<<<grid_for(3 * Mb), PA_BLOCK, 0, ctx->stream>>>(d_rndc, d_bits, ctx->d_comb, work_jac(ctx), (int)Mb)));
    if ((rc = normalize_to(ctx, d_crec, 3 * Mb, 3, 736))) return rc;
    // Schnorr proofs of A (alpha, v_A) and B (beta, v_B): 2 per record, one batch
    if ((rc = prove_dev<PA_POK>(ctx, d_crec + 64, d_rndc, nullptr, nullptr, d_cid, d_rndc + 64, d_crec + 192, 2 * Mb, LC2))) return rc;
    if ((rc = prove_dev<PA_COM>(ctx, d_crec, d_rndc, d_bits, nullptr, d_cid, d_rndc + 128, d_crec + 384, Mb, LC))) return rc;
    if (verify) {
      if ((rc = verify_dev<PA_POK, 1>(ctx, d_crec + 192, d_crec + 64, d_cid, d_cv, 2 * Mb, LC2))) return rc;  // verdicts interleaved A, B
      if ((rc = verify_dev<PA_COM, 4>(ctx, d_crec + 384, d_crec, d_cid, d_cv + 2 * Mb, Mb, LC))) return rc;
      PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_and_pairs<<<grid_for(Mb), PA_BLOCK, 0, ctx->stream>>>(d_cv, d_cv + 2 * Mb, d_cv + 3 * Mb, (int)Mb)));
    } else {
      PA_CUDA(ctx, cudaMemsetAsync(d_cv + 3 * Mb, 1, Mb, ctx->stream));
    }
    std::vector<unsigned char> cv(Mb);
    PA_CUDA(ctx, cudaMemcpyAsync(cv.data(), d_cv + 3 * Mb, Mb, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = d2h(job->out_commit, d_crec, Mb * 736))) return rc;
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (job->out_commit_ok) memcpy(job->out_commit_ok, cv.data(), Mb);
    for (size_t s = 0; s < m; ++s)
      for (u32 k = boff[s]; k < boff[s + 1]; ++k) okv[auc[s]] &= cv[k];
  }

  // ================= auction steps =====================================================
  for (size_t step = 0; step < cmax; ++step) {
    // active bidders and their per-auction segments
    std::vector<u32> act, pauc, pseg, soff(1, 0), g1, g2, g1s, g2s, actauc;
    std::vector<u64> pid, g1id, g2id;
    for (size_t a = 0; a < A; ++a) {
      if (job->c[a] <= step) continue;
      for (u32 s = aoff[a]; s < aoff[a + 1]; ++s) {
        u32 p = (u32)act.size();
        act.push_back(s), pauc.push_back((u32)a), pseg.push_back((u32)actauc.size()), pid.push_back(ids[s]);
        (junction[a] ? g2 : g1).push_back(p);
        (junction[a] ? g2s : g1s).push_back(s);
        (junction[a] ? g2id : g1id).push_back(ids[s]);
      }
      actauc.push_back((u32)a);
      soff.push_back((u32)act.size());
    }
    const size_t ma = act.size(), na = actauc.size(), n1 = g1.size(), n2 = g2.size();
    if (ma == 0) break;
    if ((rc = up(ctx, d_act, act)) || (rc = up(ctx, d_pauc, pauc)) || (rc = up(ctx, d_pseg, pseg)) || (rc = up(ctx, d_pid, pid)) ||
        (rc = up(ctx, d_soff, soff)) || (rc = up(ctx, d_g, g1)) || (rc = up(ctx, d_g + m, g2)) || (rc = up(ctx, d_gslot, g1s)) ||
        (rc = up(ctx, d_gslot + m, g2s)) || (rc = up(ctx, d_gid, g1id)) || (rc = up(ctx, d_gid + m, g2id)))
      return rc;

    // ---- round one: x, r, X = g^x, R = g^r, two Schnorr proofs ---------------------
    PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_streams, d_ctr, d_act, 4, d_rnd1, (int)ma)));
    if ((rc = work_reserve(ctx, 2 * ma))) return rc;
    PA_LAUNCH(ctx, PA_K_FIXED, (k_seal_r1_points<<<grid_for(2 * ma), PA_BLOCK, 0, ctx->stream>>>(d_rnd1, ctx->d_comb, work_jac(ctx), (int)ma)));
    if ((rc = normalize_to(ctx, d_r1, 2 * ma, 2, 320))) return rc;
    // Schnorr proofs of X (x, v_X) and R (r, v_R): 2 per record, one batch
    if ((rc = prove_dev<PA_POK>(ctx, d_r1, d_rnd1, nullptr, nullptr, d_pid, d_rnd1 + 64, d_r1 + 128, 2 * ma, LR2))) return rc;
    if (verify) {
      if ((rc = verify_dev<PA_POK, 1>(ctx, d_r1 + 128, d_r1, d_pid, d_r1v, 2 * ma, LR2))) return rc;
      PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_and_pairs<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(d_r1v, nullptr, d_r1v + 2 * ma, (int)ma)));
    } else {
      PA_CUDA(ctx, cudaMemsetAsync(d_r1v + 2 * ma, 1, ma, ctx->stream));
    }

    // ---- Y reconstruction -------------------------------------------------------------
    const unsigned char *Yloc = d_Y;
    if (!sharded) {
      if ((rc = work_reserve(ctx, ma))) return rc;
      PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<(unsigned)na, PA_SCAN_T, 0, ctx->stream>>>(d_r1, 320, d_soff, (int)ma, work_jac(ctx))));
      if ((rc = normalize_to(ctx, d_Y, ma))) return rc;
    } else {
      // all-gather the X_i of every slice, then every rank scans the whole auction
      const size_t nall = job->n[0];
      PA_CUDA(ctx, cudaMemsetAsync(job->d_send, 0, (size_t)job->slice * 64, ctx->stream));
      PA_CUDA(ctx, cudaMemcpy2DAsync(job->d_send, 64, d_r1, 320, 64, ma, cudaMemcpyDeviceToDevice, ctx->stream));
      PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (job->allgather(job->user, 0) != 0) return pa_fail(ctx, PA_EINVAL, "pa_seal_run: all-gather callback failed (X)");
      if ((rc = work_reserve(ctx, nall))) return rc;
      PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<1, PA_SCAN_T, 0, ctx->stream>>>(job->d_recv, 64, nullptr, (int)nall, work_jac(ctx))));
      if ((rc = normalize_to(ctx, d_Y, nall))) return rc;
      Yloc = d_Y + 64 * (size_t)job->lo;
    }

    // ---- round two: cryptogram b and its OR proof ---------------------------------------
    if ((rc = work_reserve(ctx, ma))) return rc;
    PA_LAUNCH(ctx, PA_K_VAR, (k_seal_encode<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(d_act, d_pauc, d_bits, d_boff, (int)step, d_junc, d_prevbit, d_r1, Yloc, d_rnd1, d_ebit, work_jac(ctx), (int)ma)));
    if ((rc = normalize_to(ctx, d_b, ma))) return rc;
    PA_CUDA(ctx, cudaMemsetAsync(d_r2v, 1, ma, ctx->stream));
    if (n1) {
      PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_stmt<<<grid_for(n1), PA_BLOCK, 0, ctx->stream>>>(1, d_g, d_act, d_boff, (int)step, d_b, d_r1, Yloc, d_rnd1, d_crec, d_rndc, d_prevpts, d_prevx, d_ebit, d_prevbit, d_stmt, d_sec, d_bi, d_bj, (int)n1)));
      PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n1), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_streams, d_ctr, d_gslot, 5, d_rnd2, (int)n1)));
      if ((rc = prove_dev<PA_S1>(ctx, d_stmt, d_sec, d_bi, nullptr, d_gid, d_rnd2, d_proof, n1))) return rc;
      if (verify) {
        if ((rc = verify_dev<PA_S1, 8>(ctx, d_proof, d_stmt, d_gid, d_pv, n1))) return rc;
        PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_scatter_u8<<<grid_for(n1), PA_BLOCK, 0, ctx->stream>>>(d_g, d_pv, d_r2v, (int)n1)));
      }
      // group-compact proofs go back to the host per group (stage-1 members first)
      if (job->out_r2_proof) {
        std::vector<unsigned char> tmp(n1 * 672);
        PA_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_proof, tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t q = 0; q < n1; ++q)
          memcpy(job->out_r2_proof + (step * m + act[g1[q]]) * 1344, tmp.data() + 672 * q, 672);
      }
    }
    if (n2) {
      PA_LAUNCH(ctx, PA_K_ENCODE, (k_seal_stmt<<<grid_for(n2), PA_BLOCK, 0, ctx->stream>>>(2, d_g + m, d_act, d_boff, (int)step, d_b, d_r1, Yloc, d_rnd1, d_crec, d_rndc, d_prevpts, d_prevx, d_ebit, d_prevbit, d_stmt, d_sec, d_bi, d_bj, (int)n2)));
      PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n2), PA_BLOCK, 0, ctx->stream>>>(job->seed, d_streams, d_ctr, d_gslot + m, 11, d_rnd2, (int)n2)));
      if ((rc = prove_dev<PA_S2>(ctx, d_stmt, d_sec, d_bi, d_bj, d_gid + m, d_rnd2, d_proof, n2))) return rc;
      if (verify) {
        if ((rc = verify_dev<PA_S2, 16>(ctx, d_proof, d_stmt, d_gid + m, d_pv, n2))) return rc;
        PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_scatter_u8<<<grid_for(n2), PA_BLOCK, 0, ctx->stream>>>(d_g + m, d_pv, d_r2v, (int)n2)));
      }
      if (job->out_r2_proof) {
        std::vector<unsigned char> tmp(n2 * 1344);
        PA_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_proof, tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
        PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t q = 0; q < n2; ++q)
          memcpy(job->out_r2_proof + (step * m + act[g2[q]]) * 1344, tmp.data() + 1344 * q, 1344);
      }
    }

    // ---- round three: is the sum of the cryptograms the point at infinity? ------------------
    if (!sharded) {
      PA_LAUNCH(ctx, PA_K_SUMINF, (k_point_sum_is_inf<<<(unsigned)na, PA_SCAN_T, 0, ctx->stream>>>(d_b, 64, d_soff, (int)ma, d_isinf)));
    } else {
      PA_CUDA(ctx, cudaMemsetAsync(job->d_send, 0, (size_t)job->slice * 64, ctx->stream));
      PA_CUDA(ctx, cudaMemcpyAsync(job->d_send, d_b, ma * 64, cudaMemcpyDeviceToDevice, ctx->stream));
      PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (job->allgather(job->user, 1) != 0) return pa_fail(ctx, PA_EINVAL, "pa_seal_run: all-gather callback failed (b)");
      PA_LAUNCH(ctx, PA_K_SUMINF, (k_point_sum_is_inf<<<1, PA_SCAN_T, 0, ctx->stream>>>(job->d_recv, 64, nullptr, (int)job->n[0], d_isinf)));
    }
    PA_LAUNCH(ctx, PA_K_VERDICT, (k_seal_update<<<grid_for(ma), PA_BLOCK, 0, ctx->stream>>>(d_act, d_pseg, d_pauc, d_isinf, d_bits, d_boff, (int)step, d_r1, Yloc, d_b, d_rnd1, d_prevpts, d_prevx, d_prevbit, d_junc, (int)ma)));

    // ---- results of the step back to the host ---------------------------------------------------
    std::vector<int> isinf(na);
    std::vector<unsigned char> r1v(ma), r2v(ma), ebit(ma);
    PA_CUDA(ctx, cudaMemcpyAsync(isinf.data(), d_isinf, na * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaMemcpyAsync(r1v.data(), d_r1v + 2 * ma, ma, cudaMemcpyDeviceToHost, ctx->stream));
    PA_CUDA(ctx, cudaMemcpyAsync(r2v.data(), d_r2v, ma, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<unsigned char> r1h, bh;
    if (job->out_r1) {
      r1h.resize(ma * 320);
      PA_CUDA(ctx, cudaMemcpyAsync(r1h.data(), d_r1, ma * 320, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (job->out_r2_b) {
      bh.resize(ma * 64);
      PA_CUDA(ctx, cudaMemcpyAsync(bh.data(), d_b, ma * 64, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t p = 0; p < ma; ++p) {
      size_t o = step * m + act[p];
      okv[pauc[p]] &= r1v[p] & r2v[p];
      if (job->out_r1) memcpy(job->out_r1 + o * 320, r1h.data() + 320 * p, 320);
      if (job->out_r2_b) memcpy(job->out_r2_b + o * 64, bh.data() + 64 * p, 64);
      if (job->out_r1_ok) job->out_r1_ok[o] = r1v[p];
      if (job->out_r2_ok) job->out_r2_ok[o] = r2v[p];
      if (job->out_r2_tag) job->out_r2_tag[o] = junction[pauc[p]] ? 2 : 1;
    }
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
  for (size_t a = 0; a < A; ++a) {
    if (job->max_bid) job->max_bid[a] = maxbid[a];
    if (job->ok) job->ok[a] = okv[a];
  }
  return PA_OK;
}
