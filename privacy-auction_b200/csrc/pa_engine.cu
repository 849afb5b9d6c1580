// Host side of the C ABI declared in include/pa_engine.h.  Nothing in this file
// computes on the CPU: every entry point stages buffers and launches the
// kernels of pa_kernels.cuh on the context's stream.
#include "../../include/pa_engine.h"
#include "pa_kernels.cuh"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <string>
#include <vector>

struct pa_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  u32 *d_comb = nullptr;  // PA_COMB_WORDS
  // grow-only device arenas
  unsigned char *d_work = nullptr;  // Jacobian scratch + prefix products
  size_t work_bytes = 0;
  unsigned char *d_stage = nullptr;  // staging for the host-buffer entry points
  size_t stage_bytes = 0;
  unsigned char *d_pool = nullptr;  // state of the whole-auction runner (pa_seal_run)
  size_t pool_bytes = 0;
  unsigned char *d_aux = nullptr;  // intermediate points of the composite CCS22 entry points
  size_t aux_bytes = 0;
  // side lanes of the whole-auction runner: a stream with its own work arena each (created on first use)
  struct Lane {
    cudaStream_t stream = nullptr;
    unsigned char *work = nullptr;
    size_t work_bytes = 0;
  } lanes[3];
  cudaEvent_t lane_ev[10] = {};
  // copy streams + events of the chunked host-buffer pipeline (created on first use)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaStream_t s_alt = nullptr;  // second compute stream of the pipeline, with its own work arena
  unsigned char *d_work_alt = nullptr;
  size_t work_alt_bytes = 0;
  cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_comp[3] = {nullptr, nullptr, nullptr}, ev_out[3] = {nullptr, nullptr, nullptr};
  uint64_t launches = 0;
  uint64_t reruns = 0;  // phase-major auctions run again step-major because a draw was rejected
  std::string err;
  // draw-stream configuration mirrored in the device constant pa_rng_config (pa_ctx_set_entropy, pa_debug_set)
  pa_rng_cfg rng_cfg = {};
  // TEST HOOK (pa_debug_set): flip one byte of one published record between proving and verifying
  struct Corrupt { int section = 0; size_t step = 0, bidder = 0, offset = 0; } corrupt;
  // peer exchange window of the bidder-sharded auction (pa_xchg_*)
  struct Xchg {
    unsigned char *local = nullptr;      // this rank's window (cudaMalloc, exported by IPC handle)
    unsigned char **d_peers = nullptr;   // device array [world] of every rank's window as mapped here
    std::vector<unsigned char *> peers;  // host copy; peers[rank] == local
    int world = 0, rank = -1;
    uint32_t epoch = 0;                  // one per sharded run, the same on every rank
    int *d_err = nullptr;                // set by a kernel whose wait for a peer timed out
  } xchg;
  // optional per-kernel CUDA-event timing (pa_profile_begin / pa_profile_end)
  int prio_main = 0;                   // priority of `stream` (the greatest the device offers)
  bool profiling = false;
  struct Span { int kid; cudaEvent_t e0, e1; cudaStream_t st; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
};

enum {
  PA_K_COMB = 0, PA_K_FIXED, PA_K_VAR, PA_K_DOUBLE, PA_K_LINCOMB2, PA_K_POINT_ADD, PA_K_NORMALIZE, PA_K_ENCODE,
  PA_K_PEAK, PA_K_VERDICT, PA_K_CHALLENGE, PA_K_RNG, PA_K_COMMIT, PA_K_YSCAN, PA_K_SUMINF,
  // per proof kind (+ PA_POK / PA_COM / PA_S1 / PA_S2)
  PA_K_VDERIVE, PA_K_VCHECKS = PA_K_VDERIVE + 4, PA_K_POPS = PA_K_VCHECKS + 4, PA_K_PRESPOND = PA_K_POPS + 4,
  PA_K_COUNT = PA_K_PRESPOND + 4
};
static const char *const PA_K_NAMES[PA_K_COUNT] = {"k_comb", "k_fixed_base", "k_var_base", "k_double_mul", "k_lincomb2",
                                                   "k_point_add", "k_normalize", "k_encode", "k_peak", "k_verdict",
                                                   "k_challenge", "k_rng_fill", "k_commit_points", "k_y_scan",
                                                   "k_point_sum_is_inf",
                                                   "k_verify_derive<pok>", "k_verify_derive<com>", "k_verify_derive<s1>", "k_verify_derive<s2>",
                                                   "k_verify_checks<pok>", "k_verify_checks<com>", "k_verify_checks<s1>", "k_verify_checks<s2>",
                                                   "k_prove_ops<pok>", "k_prove_ops<com>", "k_prove_ops<s1>", "k_prove_ops<s2>",
                                                   "k_prove_respond<pok>", "k_prove_respond<com>", "k_prove_respond<s1>", "k_prove_respond<s2>"};

static cudaEvent_t ev_get(pa_ctx *ctx) {
  if (!ctx->ev_pool.empty()) {
    cudaEvent_t e = ctx->ev_pool.back();
    ctx->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// launch a kernel on the context's stream; when profiling, bracket it with events
#define PA_LAUNCH(ctx, kid, ...)                                   \
  do {                                                             \
    pa_ctx::Span sp_{kid, nullptr, nullptr, (ctx)->stream};                       \
    if ((ctx)->profiling) {                                        \
      sp_.e0 = ev_get(ctx);                                        \
      sp_.e1 = ev_get(ctx);                                        \
      cudaEventRecord(sp_.e0, (ctx)->stream);                      \
    }                                                              \
    __VA_ARGS__;                                                   \
    if ((ctx)->profiling) {                                        \
      cudaEventRecord(sp_.e1, (ctx)->stream);                      \
      (ctx)->spans.push_back(sp_);                                 \
    }                                                              \
    (ctx)->launches++;                                             \
    PA_CUDA(ctx, cudaGetLastError());                              \
  } while (0)

static std::string g_create_err;

#define PA_CUDA(ctx, call)                                                                      \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      char buf_[512];                                                                           \
      snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      (ctx)->err = buf_;                                                                        \
      return PA_ECUDA;                                                                          \
    }                                                                                           \
  } while (0)

// Every exported function runs on the context's device whatever the caller's current device is
// (torch may have changed it; a process may hold contexts on several GPUs) and restores it on return.
struct pa_dev_guard {
  int prev = -1;
  bool switched = false;
  explicit pa_dev_guard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~pa_dev_guard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define PA_ENTER(ctx)              \
  if (!(ctx)) return PA_EINVAL;    \
  pa_dev_guard dev_guard_((ctx)->device)

static int pa_fail(pa_ctx *ctx, int code, const char *msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define PA_ARGCHECK(ctx, cond) \
  if (!(cond)) return pa_fail(ctx, PA_EINVAL, "invalid argument: " #cond)

static inline unsigned grid_for(size_t n) { return (unsigned)((n + PA_BLOCK - 1) / PA_BLOCK); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int ensure(pa_ctx *ctx, unsigned char **buf, size_t *cap, size_t need) {
  if (need <= *cap) return PA_OK;
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*buf) PA_CUDA(ctx, cudaFree(*buf));
  *buf = nullptr;
  *cap = 0;
  size_t want = align_up(need + need / 4, 1 << 20);
  PA_CUDA(ctx, cudaMalloc((void **)buf, want));
  *cap = want;
  return PA_OK;
}

extern "C" {

int pa_abi_version(void) { return 3; }  // 3: pa_seal_job.use_xchg / ok_all, peer window, entropy key, verifiers validate points
size_t pa_abi_sizeof(int which) {
  return which == 0 ? sizeof(pa_seal_job) : which == 1 ? sizeof(pa_ccs22_job) : which == 2 ? sizeof(pa_kernel_stat) : 0;
}

const char *pa_last_error(pa_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pa_ctx_create(pa_ctx **out, int device) {
  if (!out) return PA_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                   " (this engine has no CPU path)";
    return PA_ENODEV;
  }
  if (device < 0 || device >= count) {
    g_create_err = "device ordinal out of range";
    return PA_EINVAL;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    g_create_err = "device is not compute capability 10.x (the engine ships sm_100a code only)";
    return PA_ENODEV;
  }
  pa_ctx *ctx = new pa_ctx();
  ctx->device = device;
  pa_dev_guard guard(device);
  u32 *d_bases = nullptr;
  auto fail = [&](const char *what, cudaError_t ce) {
    g_create_err = std::string(what) + ": " + cudaGetErrorString(ce);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bases);
    cudaFree(ctx->d_comb);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PA_ECUDA;
  };
  int cur = -1;
  if ((e = cudaGetDevice(&cur)) != cudaSuccess || cur != device) return fail("cudaSetDevice", e != cudaSuccess ? e : cudaErrorInvalidDevice);
  // The context's own stream carries the sequential chain of a run (keys -> Y scan -> cryptograms -> walk); the side
  // lanes of the runners (proofs, verification: bulk work without a successor) are created at the lowest priority, so
  // that blocks of the chain are placed first whenever SM slots free up.
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  ctx->prio_main = prio_greatest;
  if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaMalloc((void **)&ctx->d_comb, PA_COMB_WORDS * sizeof(u32))) != cudaSuccess) return fail("cudaMalloc(comb)", e);
  if ((e = cudaMalloc((void **)&d_bases, PA_COMB_WINDOWS * 16 * sizeof(u32))) != cudaSuccess) return fail("cudaMalloc(bases)", e);
  k_comb_base<<<1, 32, 0, ctx->stream>>>(d_bases);
  k_comb_entries<<<PA_COMB_WINDOWS * PA_COMB_ENTRIES / PA_BLOCK, PA_BLOCK, 0, ctx->stream>>>(d_bases, ctx->d_comb);
  ctx->launches += 2;
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail("comb table build", e);
  cudaFree(d_bases);
  d_bases = nullptr;
  // the draw stream starts as the seeded test stream (pa_ctx_set_entropy installs a key)
  if ((e = cudaMemcpyToSymbol(pa_rng_config, &ctx->rng_cfg, sizeof ctx->rng_cfg)) != cudaSuccess) return fail("rng config", e);
  *out = ctx;
  return PA_OK;
}

int pa_ctx_destroy(pa_ctx *ctx) {
  PA_ENTER(ctx);
  cudaStreamSynchronize(ctx->stream);
  pa_xchg_close(ctx);
  cudaFree(ctx->d_comb);
  cudaFree(ctx->d_work);
  cudaFree(ctx->d_stage);
  cudaFree(ctx->d_pool);
  cudaFree(ctx->d_aux);
  cudaFree(ctx->d_work_alt);
  if (ctx->s_alt) cudaStreamDestroy(ctx->s_alt);
  for (auto &l : ctx->lanes) {
    if (l.stream) cudaStreamDestroy(l.stream);
    cudaFree(l.work);
  }
  for (auto e : ctx->lane_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->s_in) {
    cudaStreamDestroy(ctx->s_in);
    cudaStreamDestroy(ctx->s_out);
    for (int i = 0; i < 3; ++i) cudaEventDestroy(ctx->ev_in[i]), cudaEventDestroy(ctx->ev_comp[i]), cudaEventDestroy(ctx->ev_out[i]);
  }
  for (auto &sp : ctx->spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PA_OK;
}

int pa_sync(pa_ctx *ctx) {
  PA_ENTER(ctx);
  if (!ctx) return PA_EINVAL;
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PA_OK;
}

void *pa_ctx_stream(pa_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t pa_ctx_launches(pa_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pa_dev_alloc(pa_ctx *ctx, void **dptr, size_t bytes) {
  PA_ENTER(ctx);
  if (!ctx || !dptr) return PA_EINVAL;
  PA_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
  return PA_OK;
}
int pa_dev_free(pa_ctx *ctx, void *dptr) {
  PA_ENTER(ctx);
  if (!ctx) return PA_EINVAL;
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PA_CUDA(ctx, cudaFree(dptr));
  return PA_OK;
}
int pa_dev_upload(pa_ctx *ctx, void *dptr, const void *host, size_t bytes) {
  PA_ENTER(ctx);
  if (!ctx) return PA_EINVAL;
  PA_CUDA(ctx, cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return PA_OK;
}
int pa_dev_download(pa_ctx *ctx, void *host, const void *dptr, size_t bytes) {
  PA_ENTER(ctx);
  if (!ctx) return PA_EINVAL;
  PA_CUDA(ctx, cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PA_OK;
}

// ---- draw stream configuration, test hooks ----------------------------------------------------------
static int rng_cfg_push(pa_ctx *ctx) {
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PA_CUDA(ctx, cudaMemcpyToSymbol(pa_rng_config, &ctx->rng_cfg, sizeof ctx->rng_cfg));
  return PA_OK;
}
int pa_ctx_set_entropy(pa_ctx *ctx, const uint8_t *key32) {
  PA_ENTER(ctx);
  ctx->rng_cfg.keyed = key32 ? 1u : 0u;
  for (int w = 0; w < 8; ++w)
    ctx->rng_cfg.key[w] = key32 ? ((u32)key32[4 * w] << 24) | ((u32)key32[4 * w + 1] << 16) | ((u32)key32[4 * w + 2] << 8) | key32[4 * w + 3] : 0u;
  return rng_cfg_push(ctx);
}
int pa_debug_set(pa_ctx *ctx, int what, uint64_t a, uint64_t b, uint64_t c) {
  PA_ENTER(ctx);
  if (what == PA_DBG_REJECT_BITS) {
    if (a > 16) return pa_fail(ctx, PA_EINVAL, "pa_debug_set: reject bits must be 0..16");
    ctx->rng_cfg.reject_bits = (u32)a;
    return rng_cfg_push(ctx);
  }
  if (what == PA_DBG_CORRUPT) {
    if (a > 3) return pa_fail(ctx, PA_EINVAL, "pa_debug_set: section must be 0 (off), 1 (commitment), 2 (round one) or 3 (round two)");
    ctx->corrupt.section = (int)a;
    ctx->corrupt.step = (size_t)(b >> 32);
    ctx->corrupt.bidder = (size_t)(b & 0xFFFFFFFFu);
    ctx->corrupt.offset = (size_t)c;
    return PA_OK;
  }
  return pa_fail(ctx, PA_EINVAL, "pa_debug_set: unknown hook");
}

// ---- peer exchange window ---------------------------------------------------------------------------
// One auction sharded by bidder slice exchanges, per step, each rank's sum of cryptograms (and once per
// pass each rank's sum of public keys).  With a window the ranks' kernels write those 96-byte values
// straight into each other's HBM over NVLink and wait on tags there: no host round trip, no collective
// call per step.  Layout and protocol: pa_seal.cuh, "peer exchange".
int pa_xchg_create(pa_ctx *ctx, uint8_t *handle64) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, handle64 != nullptr);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ctx->xchg.local) {
    PA_CUDA(ctx, cudaMalloc((void **)&ctx->xchg.local, PA_XCHG_BYTES));
    PA_CUDA(ctx, cudaMemset(ctx->xchg.local, 0, PA_XCHG_BYTES));
    PA_CUDA(ctx, cudaMalloc((void **)&ctx->xchg.d_err, sizeof(int)));
    PA_CUDA(ctx, cudaMemset(ctx->xchg.d_err, 0, sizeof(int)));
    PA_CUDA(ctx, cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  PA_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->xchg.local));
  memcpy(handle64, &h, 64);
  return PA_OK;
}
int pa_xchg_connect(pa_ctx *ctx, const uint8_t *handles, int world, int rank) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, handles && world >= 1 && world <= PA_XCHG_MAX_WORLD && rank >= 0 && rank < world && ctx->xchg.local);
  PA_ARGCHECK(ctx, ctx->xchg.world == 0);
  std::vector<unsigned char *> peers(world, nullptr);
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      peers[r] = ctx->xchg.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    cudaError_t e = cudaIpcOpenMemHandle((void **)&peers[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != rank) cudaIpcCloseMemHandle(peers[q]);
      ctx->err = std::string("pa_xchg_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
      return PA_ECUDA;
    }
  }
  PA_CUDA(ctx, cudaMalloc((void **)&ctx->xchg.d_peers, world * sizeof(unsigned char *)));
  PA_CUDA(ctx, cudaMemcpy(ctx->xchg.d_peers, peers.data(), world * sizeof(unsigned char *), cudaMemcpyHostToDevice));
  ctx->xchg.peers = peers;
  ctx->xchg.world = world;
  ctx->xchg.rank = rank;
  ctx->xchg.epoch = 0;
  return PA_OK;
}
int pa_xchg_close(pa_ctx *ctx) {
  PA_ENTER(ctx);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < ctx->xchg.world; ++r)
    if (r != ctx->xchg.rank && ctx->xchg.peers[r]) cudaIpcCloseMemHandle(ctx->xchg.peers[r]);
  cudaFree(ctx->xchg.d_peers);
  cudaFree(ctx->xchg.local);
  cudaFree(ctx->xchg.d_err);
  ctx->xchg = pa_ctx::Xchg();
  return PA_OK;
}

}  // extern "C"

// ---- internal: Jacobian scratch -> affine output ------------------------------------
// work arena layout: [ n * 96 B Jacobian | n * 32 B prefix products ]
static int work_reserve(pa_ctx *ctx, size_t n) { return ensure(ctx, &ctx->d_work, &ctx->work_bytes, n * 128 + 256); }
static u32 *work_jac(pa_ctx *ctx) { return (u32 *)ctx->d_work; }
static u32 *work_prefix(pa_ctx *ctx, size_t n) { return (u32 *)(ctx->d_work + n * 96); }

static int normalize_to(pa_ctx *ctx, unsigned char *d_out, size_t n, int nper = 1, size_t stride = 64, int inner = 1,
                        size_t stride_in = 0) {
  // points per thread: share the ~270-multiplication inversion among up to 16 points as soon as
  // that still leaves >= 16 k threads (one lone warp per SM sub-partition is latency-bound anyway)
  size_t per = n / 16384;
  if (per < 1) per = 1;
  if (per > 16) per = 16;
  size_t T = (n + per - 1) / per;
  PA_LAUNCH(ctx, PA_K_NORMALIZE, k_normalize<<<grid_for(T), PA_BLOCK, 0, ctx->stream>>>(work_jac(ctx), work_prefix(ctx, n), pa_outlay{d_out, nper, stride, inner, stride_in}, (int)n, (int)T));
  return PA_OK;
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

extern "C" {

int pa_fixed_base_mul_dev(pa_ctx *ctx, const uint8_t *d_scalars, uint8_t *d_out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (d_scalars && d_out)) && n < (1u << 30));
  PA_ARGCHECK(ctx, aligned16(d_scalars) && aligned16(d_out));
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_FIXED, k_fixed_base<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(d_scalars, ctx->d_comb, work_jac(ctx), (int)n));
  return normalize_to(ctx, d_out, n);
}

int pa_var_base_mul_dev(pa_ctx *ctx, const uint8_t *d_points, const uint8_t *d_scalars, uint8_t *d_out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (d_points && d_scalars && d_out)) && n < (1u << 30));
  PA_ARGCHECK(ctx, aligned16(d_points) && aligned16(d_scalars) && aligned16(d_out));
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_VAR, k_var_base<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(d_points, d_scalars, work_jac(ctx), (int)n));
  return normalize_to(ctx, d_out, n);
}

int pa_double_mul_dev(pa_ctx *ctx, const uint8_t *d_a, const uint8_t *d_points, const uint8_t *d_b, uint8_t *d_out,
                      size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (d_a && d_points && d_b && d_out)) && n < (1u << 30));
  PA_ARGCHECK(ctx, aligned16(d_a) && aligned16(d_points) && aligned16(d_b) && aligned16(d_out));
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_DOUBLE, k_double_mul<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(d_a, d_points, d_b, ctx->d_comb, work_jac(ctx), (int)n));
  return normalize_to(ctx, d_out, n);
}

int pa_lincomb2_dev(pa_ctx *ctx, const uint8_t *d_p, const uint8_t *d_a, const uint8_t *d_q, const uint8_t *d_b,
                    uint8_t *d_out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (d_p && d_a && d_q && d_b && d_out)) && n < (1u << 30));
  PA_ARGCHECK(ctx, aligned16(d_p) && aligned16(d_a) && aligned16(d_q) && aligned16(d_b) && aligned16(d_out));
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_LINCOMB2, k_lincomb2<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(d_p, d_a, d_q, d_b, work_jac(ctx), (int)n));
  return normalize_to(ctx, d_out, n);
}

}  // extern "C"

// ---- host-buffer wrappers: stage in, run, stage out ---------------------------------
namespace {
struct Stage {
  pa_ctx *ctx;
  size_t off = 0;
  explicit Stage(pa_ctx *c) : ctx(c) {}
  unsigned char *take(size_t bytes) {
    unsigned char *p = ctx->d_stage + off;
    off += align_up(bytes, 256);
    return p;
  }
};
int stage_reserve(pa_ctx *ctx, size_t bytes) { return ensure(ctx, &ctx->d_stage, &ctx->stage_bytes, bytes); }
}  // namespace

namespace {
// Large host-buffer batches are processed in chunks so that the host->device copy of chunk
// k+1 and the device->host copy of chunk k-1 overlap the kernels of chunk k (three streams,
// three staging slots).  `per` = bytes per item of each argument.
struct PArg {
  const void *in;
  void *out;
  size_t per;
};
// A chunk is a whole number of waves of the scalar-multiplication kernels (148 SMs x 128 threads x 4 or 5
// resident blocks): 2^17 items left the second wave of k_var_base 38 % full.
// Measured end to end (2^20 + 2^20 mults per step): 2^17 items 82.7 M/s, 10 blocks per SM 84.7, 20: 85.2;
// a short first and last chunk (5 blocks per SM) shortens the copies that nothing overlaps: 86.5; chunks alternating
// between two compute streams (the next chunk's blocks fill the SMs while the previous chunk's last wave drains): 90.1.
// With k_var_base at 6 resident blocks (end of r01: 100.7 M/s), chunks of 24 / 6 blocks per SM measured 100.5: left at 20 / 5.
#ifndef PA_PIPE_CHUNK_BLOCKS
#define PA_PIPE_CHUNK_BLOCKS 20  // blocks per SM in a chunk
#endif
#ifndef PA_PIPE_EDGE_BLOCKS
#define PA_PIPE_EDGE_BLOCKS 5  // ... in the first and the last chunk
#endif
const size_t PA_PIPE_UNIT = (size_t)148 * PA_BLOCK;
const size_t PA_PIPE_CHUNK = PA_PIPE_UNIT * PA_PIPE_CHUNK_BLOCKS, PA_PIPE_EDGE = PA_PIPE_UNIT * PA_PIPE_EDGE_BLOCKS;

// One job of a pipelined call: n items, its per-item arguments, and the device-pointer implementation.
struct PJob {
  size_t n;
  PArg args[6];
  int nargs;
  std::function<int(unsigned char **, size_t)> run;
};
struct PChunk {
  int job;
  size_t off, cnt;
  double pos;  // middle of the chunk as a fraction of its job: chunks of several jobs are interleaved by this
};

// Several jobs in ONE copy/compute pipeline: the chunks of the jobs are interleaved (by their relative position in their
// job), so that the copies of a copy-heavy job (fixed base: 96 bytes per 2 us of GPU time) hide behind the kernels of a
// compute-heavy one (variable base: 160 bytes per 17 us).  Results are those of running the jobs one after the other.
int pipelined_jobs(pa_ctx *ctx, std::vector<PJob> &jobs) {
  if (!ctx->s_in) {
    PA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    PA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    PA_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->s_alt, cudaStreamNonBlocking, ctx->prio_main));  // alternates with the main stream
    for (int i = 0; i < 3; ++i) {
      PA_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
      PA_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
      PA_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    }
  }
  // whatever happens below, no copy into or out of the caller's buffers is left in flight
  struct Drain {
    pa_ctx *c;
    cudaStream_t main;
    ~Drain() {
      cudaStreamSynchronize(c->s_in);
      cudaStreamSynchronize(c->s_alt);
      cudaStreamSynchronize(main);
      cudaStreamSynchronize(c->s_out);
    }
  } drain{ctx, ctx->stream};
  // chunks: whole waves; a job longer than one chunk starts and ends with a short one (the first copy in and the last
  // copy out of a call are not hidden by anything)
  std::vector<PChunk> chunks;
  size_t nmax = 0, slot_bytes = 1024;
  for (size_t j = 0; j < jobs.size(); ++j) {
    const PJob &J = jobs[j];
    nmax = J.n > nmax ? J.n : nmax;
    const size_t CHj = J.n < PA_PIPE_CHUNK ? J.n : PA_PIPE_CHUNK;
    size_t bytes = 1024;
    for (int i = 0; i < J.nargs; ++i) bytes += align_up(J.args[i].per * CHj, 256);
    slot_bytes = bytes > slot_bytes ? bytes : slot_bytes;
    size_t k = 0, cnt = 0;
    for (size_t off = 0; off < J.n; off += cnt, ++k) {
      const size_t left = J.n - off;
      if (J.n <= PA_PIPE_CHUNK) cnt = J.n;
      else if (k == 0) cnt = PA_PIPE_EDGE;
      else if (left > PA_PIPE_CHUNK + PA_PIPE_EDGE) cnt = PA_PIPE_CHUNK;
      else if (left > PA_PIPE_EDGE) cnt = left - PA_PIPE_EDGE;
      else cnt = left;
      chunks.push_back(PChunk{(int)j, off, cnt, ((double)off + 0.5 * (double)cnt) / (double)J.n});
    }
  }
  if (chunks.empty()) return PA_OK;
  std::stable_sort(chunks.begin(), chunks.end(), [](const PChunk &a, const PChunk &b) { return a.pos < b.pos; });
  const size_t CH = nmax < PA_PIPE_CHUNK ? nmax : PA_PIPE_CHUNK;
  const int slots = chunks.size() > 1 ? 3 : 1;
  int rc = stage_reserve(ctx, slot_bytes * slots + 1024);
  if (rc) return rc;
  if ((rc = work_reserve(ctx, CH))) return rc;  // no arena growth (= sync) inside the pipeline
  // Chunks alternate between two compute streams (each with its own Jacobian scratch), so the blocks of the
  // next chunk fill the SMs while the last wave of the previous one drains.
  auto swap_alt = [&]() {
    std::swap(ctx->stream, ctx->s_alt);
    std::swap(ctx->d_work, ctx->d_work_alt);
    std::swap(ctx->work_bytes, ctx->work_alt_bytes);
  };
  if (chunks.size() > 1) {
    swap_alt();
    rc = work_reserve(ctx, CH);
    swap_alt();
    if (rc) return rc;
  }
  for (size_t k = 0; k < chunks.size(); ++k) {
    const PChunk &C = chunks[k];
    const PJob &J = jobs[C.job];
    const size_t CHj = J.n < PA_PIPE_CHUNK ? J.n : PA_PIPE_CHUNK;
    const int slot = (int)(k % slots);
    unsigned char *base = ctx->d_stage + slot_bytes * slot, *d[16];
    size_t o = 0;
    for (int i = 0; i < J.nargs; ++i) {
      d[i] = base + o;
      o += align_up(J.args[i].per * CHj, 256);
    }
    if (k >= (size_t)slots) PA_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_out[slot], 0));  // slot free again
    for (int i = 0; i < J.nargs; ++i)
      if (J.args[i].in)
        PA_CUDA(ctx, cudaMemcpyAsync(d[i], (const unsigned char *)J.args[i].in + J.args[i].per * C.off, J.args[i].per * C.cnt, cudaMemcpyHostToDevice, ctx->s_in));
    PA_CUDA(ctx, cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
    const bool alt = (k & 1) != 0;
    if (alt) swap_alt();
    cudaError_t e1 = cudaStreamWaitEvent(ctx->stream, ctx->ev_in[slot], 0);
    rc = e1 == cudaSuccess ? J.run(d, C.cnt) : PA_OK;
    cudaError_t e2 = cudaEventRecord(ctx->ev_comp[slot], ctx->stream);
    if (alt) swap_alt();
    PA_CUDA(ctx, e1);
    if (rc) return rc;
    PA_CUDA(ctx, e2);
    PA_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[slot], 0));
    for (int i = 0; i < J.nargs; ++i)
      if (J.args[i].out)
        PA_CUDA(ctx, cudaMemcpyAsync((unsigned char *)J.args[i].out + J.args[i].per * C.off, d[i], J.args[i].per * C.cnt, cudaMemcpyDeviceToHost, ctx->s_out));
    PA_CUDA(ctx, cudaEventRecord(ctx->ev_out[slot], ctx->s_out));
  }
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->s_alt));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PA_OK;
}

template <typename F>
int pipelined(pa_ctx *ctx, size_t n, const PArg *args, int nargs, F run) {
  std::vector<PJob> jobs(1);
  jobs[0].n = n;
  jobs[0].nargs = nargs;
  for (int i = 0; i < nargs; ++i) jobs[0].args[i] = args[i];
  jobs[0].run = run;
  return pipelined_jobs(ctx, jobs);
}
}  // namespace

extern "C" {

int pa_fixed_base_mul(pa_ctx *ctx, const uint8_t *scalars, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (scalars && out)));
  if (n == 0) return PA_OK;
  PArg a[] = {{scalars, 0, 32}, {0, out, 64}};
  return pipelined(ctx, n, a, 2, [&](unsigned char **d, size_t cnt) { return pa_fixed_base_mul_dev(ctx, d[0], d[1], cnt); });
}

int pa_var_base_mul(pa_ctx *ctx, const uint8_t *points, const uint8_t *scalars, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (points && scalars && out)));
  if (n == 0) return PA_OK;
  PArg a[] = {{points, 0, 64}, {scalars, 0, 32}, {0, out, 64}};
  return pipelined(ctx, n, a, 3, [&](unsigned char **d, size_t cnt) { return pa_var_base_mul_dev(ctx, d[0], d[1], d[2], cnt); });
}

int pa_double_mul(pa_ctx *ctx, const uint8_t *a, const uint8_t *points, const uint8_t *b, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (a && points && b && out)));
  if (n == 0) return PA_OK;
  PArg g[] = {{a, 0, 32}, {points, 0, 64}, {b, 0, 32}, {0, out, 64}};
  return pipelined(ctx, n, g, 4, [&](unsigned char **d, size_t cnt) { return pa_double_mul_dev(ctx, d[0], d[1], d[2], d[3], cnt); });
}

int pa_lincomb2(pa_ctx *ctx, const uint8_t *p, const uint8_t *a, const uint8_t *q, const uint8_t *b, uint8_t *out,
                size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (p && a && q && b && out)));
  if (n == 0) return PA_OK;
  PArg g[] = {{p, 0, 64}, {a, 0, 32}, {q, 0, 64}, {b, 0, 32}, {0, out, 64}};
  return pipelined(ctx, n, g, 5, [&](unsigned char **d, size_t cnt) { return pa_lincomb2_dev(ctx, d[0], d[1], d[2], d[3], d[4], cnt); });
}

int pa_scalar_mul_jobs(pa_ctx *ctx, const pa_mul_job *jobs, size_t njobs) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (njobs == 0 || jobs) && njobs <= 64);
  std::vector<PJob> pj;
  for (size_t j = 0; j < njobs; ++j) {
    const pa_mul_job &m = jobs[j];
    if (m.n == 0) continue;
    PJob J;
    J.n = m.n;
    switch (m.kind) {
      case PA_MUL_FIXED:
        PA_ARGCHECK(ctx, m.a && m.out);
        J.nargs = 2;
        J.args[0] = PArg{m.a, 0, 32}; J.args[1] = PArg{0, m.out, 64};
        J.run = [ctx](unsigned char **d, size_t cnt) { return pa_fixed_base_mul_dev(ctx, d[0], d[1], cnt); };
        break;
      case PA_MUL_VAR:
        PA_ARGCHECK(ctx, m.p && m.a && m.out);
        J.nargs = 3;
        J.args[0] = PArg{m.p, 0, 64}; J.args[1] = PArg{m.a, 0, 32}; J.args[2] = PArg{0, m.out, 64};
        J.run = [ctx](unsigned char **d, size_t cnt) { return pa_var_base_mul_dev(ctx, d[0], d[1], d[2], cnt); };
        break;
      case PA_MUL_DOUBLE:
        PA_ARGCHECK(ctx, m.a && m.p && m.b && m.out);
        J.nargs = 4;
        J.args[0] = PArg{m.a, 0, 32}; J.args[1] = PArg{m.p, 0, 64}; J.args[2] = PArg{m.b, 0, 32}; J.args[3] = PArg{0, m.out, 64};
        J.run = [ctx](unsigned char **d, size_t cnt) { return pa_double_mul_dev(ctx, d[0], d[1], d[2], d[3], cnt); };
        break;
      case PA_MUL_LINCOMB2:
        PA_ARGCHECK(ctx, m.p && m.a && m.q && m.b && m.out);
        J.nargs = 5;
        J.args[0] = PArg{m.p, 0, 64}; J.args[1] = PArg{m.a, 0, 32}; J.args[2] = PArg{m.q, 0, 64}; J.args[3] = PArg{m.b, 0, 32}; J.args[4] = PArg{0, m.out, 64};
        J.run = [ctx](unsigned char **d, size_t cnt) { return pa_lincomb2_dev(ctx, d[0], d[1], d[2], d[3], d[4], cnt); };
        break;
      default:
        return pa_fail(ctx, PA_EINVAL, "pa_scalar_mul_jobs: unknown job kind");
    }
    pj.push_back(J);
  }
  return pipelined_jobs(ctx, pj);
}

int pa_point_add(pa_ctx *ctx, const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, int sub) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (p && q && out)) && n < (1u << 30));
  if (n == 0) return PA_OK;
  int rc = stage_reserve(ctx, n * 192 + 1024);
  if (rc) return rc;
  if ((rc = work_reserve(ctx, n))) return rc;
  Stage s(ctx);
  unsigned char *d_p = s.take(n * 64), *d_q = s.take(n * 64), *d_o = s.take(n * 64);
  PA_CUDA(ctx, cudaMemcpyAsync(d_p, p, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  PA_CUDA(ctx, cudaMemcpyAsync(d_q, q, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  PA_LAUNCH(ctx, PA_K_POINT_ADD, k_point_add<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(d_p, d_q, work_jac(ctx), (int)n, sub));
  if ((rc = normalize_to(ctx, d_o, n))) return rc;
  PA_CUDA(ctx, cudaMemcpyAsync(out, d_o, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PA_OK;
}

int pa_point_on_curve(pa_ctx *ctx, const uint8_t *points, size_t n, uint8_t *ok) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (points && ok)) && n < (1u << 30));
  if (n == 0) return PA_OK;
  int rc = stage_reserve(ctx, n * 65 + 1024);
  if (rc) return rc;
  Stage s(ctx);
  unsigned char *d_p = s.take(n * 64), *d_o = s.take(n);
  PA_CUDA(ctx, cudaMemcpyAsync(d_p, points, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  PA_LAUNCH(ctx, PA_K_ENCODE, k_on_curve<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(d_p, (int)n, d_o));
  PA_CUDA(ctx, cudaMemcpyAsync(ok, d_o, n, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PA_OK;
}

int pa_point_encode(pa_ctx *ctx, const uint8_t *points, size_t n, int compressed, uint8_t *out, size_t stride,
                    uint32_t *lens) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (points && out && lens)) && n < (1u << 30));
  PA_ARGCHECK(ctx, stride >= (compressed ? 33u : 65u));
  if (n == 0) return PA_OK;
  int rc = stage_reserve(ctx, n * (64 + stride + 4) + 1024);
  if (rc) return rc;
  Stage s(ctx);
  unsigned char *d_p = s.take(n * 64), *d_o = s.take(n * stride), *d_l = s.take(n * 4);
  PA_CUDA(ctx, cudaMemcpyAsync(d_p, points, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  PA_LAUNCH(ctx, PA_K_ENCODE, k_encode<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(d_p, (int)n, compressed, d_o, stride, (u32 *)d_l));
  PA_CUDA(ctx, cudaMemcpyAsync(out, d_o, n * stride, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaMemcpyAsync(lens, d_l, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PA_OK;
}

int pa_profile_begin(pa_ctx *ctx) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx);
  ctx->profiling = true;
  return PA_OK;
}

int pa_profile_end(pa_ctx *ctx, pa_kernel_stat *out, size_t cap, size_t *count) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && count && (out || cap == 0));
  ctx->profiling = false;
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double total[PA_K_COUNT] = {0};
  uint64_t cnt[PA_K_COUNT] = {0};
  // development aid: PA_TIMELINE=<path prefix> appends "start_ms dur_ms lane kernel" per launch (relative to the first
  // launch of the profiled region; lane = the stream it ran on) to <prefix>.<device>
  if (const char *tl = getenv("PA_TIMELINE")) {
    if (!ctx->spans.empty()) {
      std::string path = std::string(tl) + "." + std::to_string(ctx->device);
      if (FILE *f = fopen(path.c_str(), "a")) {
        std::vector<cudaStream_t> lanes;
        fprintf(f, "# region with %zu launches\n", ctx->spans.size());
        for (auto &sp : ctx->spans) {
          float t0 = 0, ms = 0;
          cudaEventSynchronize(sp.e1);
          cudaEventElapsedTime(&t0, ctx->spans[0].e0, sp.e0);
          cudaEventElapsedTime(&ms, sp.e0, sp.e1);
          size_t lane = std::find(lanes.begin(), lanes.end(), sp.st) - lanes.begin();
          if (lane == lanes.size()) lanes.push_back(sp.st);
          fprintf(f, "%9.3f %8.3f %zu %s\n", t0, ms, lane, PA_K_NAMES[sp.kid]);
        }
        fclose(f);
      }
    }
  }
  for (auto &sp : ctx->spans) {
    float ms = 0;
    PA_CUDA(ctx, cudaEventElapsedTime(&ms, sp.e0, sp.e1));
    total[sp.kid] += ms;
    cnt[sp.kid]++;
    ctx->ev_pool.push_back(sp.e0);
    ctx->ev_pool.push_back(sp.e1);
  }
  ctx->spans.clear();
  size_t k = 0;
  for (int i = 0; i < PA_K_COUNT; ++i) {
    if (!cnt[i]) continue;
    if (k < cap) {
      memset(&out[k], 0, sizeof out[k]);
      strncpy(out[k].name, PA_K_NAMES[i], sizeof out[k].name - 1);
      out[k].launches = cnt[i];
      out[k].total_ms = total[i];
    }
    ++k;
  }
  *count = k;
  return PA_OK;
}

int pa_measure_int_peak(pa_ctx *ctx, double out[6]) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && out);
  int rc = stage_reserve(ctx, 4096);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  PA_CUDA(ctx, cudaEventCreate(&e0));
  PA_CUDA(ctx, cudaEventCreate(&e1));
  const int blocks = 148 * 8, threads = 256;
  out[5] = 0;
  for (int which = 0; which < 5; ++which) {
    int iters = which < 2 || which == 4 ? 4096 : 2048;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      PA_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
      if (which == 0) k_peak_imad<<<blocks, threads, 0, ctx->stream>>>((u32 *)ctx->d_stage, iters, 3u + rep, 7u);
      if (which == 1) k_peak_imad_wide<<<blocks, threads, 0, ctx->stream>>>((u64 *)ctx->d_stage, iters, 3u + rep, 1);
      if (which == 4) k_peak_imad_wide<<<blocks, threads, 0, ctx->stream>>>((u64 *)ctx->d_stage, iters, 3u + rep, 0);
      if (which == 2) k_peak_fe<<<blocks, threads, 0, ctx->stream>>>((u32 *)ctx->d_stage, iters, 0);
      if (which == 3) k_peak_fe<<<blocks, threads, 0, ctx->stream>>>((u32 *)ctx->d_stage, iters, 1);
      ctx->launches++;
      PA_CUDA(ctx, cudaGetLastError());
      PA_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
      PA_CUDA(ctx, cudaEventSynchronize(e1));
      float ms = 0;
      PA_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
      double ops = (double)blocks * threads * iters * (which < 2 || which == 4 ? 64.0 : 2.0);
      double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best) best = rate;  // first repetition is warm-up
    }
    out[which] = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return PA_OK;
}

}  // extern "C"


// =====================================================================================
// Proofs, challenges, round logic
// =====================================================================================
namespace {

// Staging for the host-buffer entry points: every argument gets a 256-byte aligned
// slot in the stage arena; inputs are uploaded before and outputs downloaded after
// the device-pointer implementation runs.
struct HArg {
  const void *in;
  void *out;
  size_t bytes;
};
template <typename F>
int staged(pa_ctx *ctx, const HArg *args, int nargs, F run) {
  size_t total = 1024;
  for (int i = 0; i < nargs; ++i) total += align_up(args[i].bytes, 256) + 256;
  int rc = stage_reserve(ctx, total);
  if (rc) return rc;
  Stage s(ctx);
  unsigned char *d[16];
  for (int i = 0; i < nargs; ++i) {
    d[i] = s.take(args[i].bytes ? args[i].bytes : 1);
    if (args[i].in && args[i].bytes)
      PA_CUDA(ctx, cudaMemcpyAsync(d[i], args[i].in, args[i].bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = run(d))) return rc;
  for (int i = 0; i < nargs; ++i)
    if (args[i].out && args[i].bytes)
      PA_CUDA(ctx, cudaMemcpyAsync(args[i].out, d[i], args[i].bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PA_OK;
}

// work arena for proofs: [ Jacobian scratch | prefix | derived scalars | check bytes ]
template <int KIND, int NCHK>
int verify_dev(pa_ctx *ctx, const unsigned char *proofs, const unsigned char *stmts, const u64 *ids,
               unsigned char *verdict, size_t n, pa_lay L = pa_lay_packed<KIND>()) {
  if (n == 0) return PA_OK;
  size_t need = n * 32 + n * NCHK + n + 1024;
  int rc = ensure(ctx, &ctx->d_work, &ctx->work_bytes, need);
  if (rc) return rc;
  u32 *derived = (u32 *)ctx->d_work;
  unsigned char *chk = ctx->d_work + align_up(n * 32, 256);
  unsigned char *valid = chk + align_up(n * NCHK, 256);
  PA_LAUNCH(ctx, PA_K_VDERIVE + KIND, (k_verify_derive<KIND><<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(proofs, stmts, ids, derived, valid, (int)n, L)));
  PA_LAUNCH(ctx, PA_K_VCHECKS + KIND, (k_verify_checks<KIND, NCHK><<<grid_for(n * NCHK), PA_BLOCK, 0, ctx->stream>>>(proofs, stmts, derived, ctx->d_comb, chk, (int)n, L)));
  PA_LAUNCH(ctx, PA_K_VERDICT, (k_verdict<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(chk, valid, NCHK, (int)n, verdict)));
  return PA_OK;
}

// wit: `secrets` holds the EXTENDED secrets (pa_proof.cuh, "prover with witnesses"; L.secret is their stride) and
// cb the committed bit per proof (stage 2 only): every published point costs one fixed-base multiplication and at
// most one variable-base multiplication.  Same bytes as the plain prover.
template <int KIND>
int prove_dev(pa_ctx *ctx, const unsigned char *stmts, const unsigned char *secrets, const unsigned char *b0,
              const unsigned char *b1, const u64 *ids, const unsigned char *rnd, unsigned char *proofs, size_t n,
              pa_lay L = pa_lay_packed<KIND>(), bool wit = false, const unsigned char *cb = nullptr) {
  typedef proof_kind<KIND> K;
  if (n == 0) return PA_OK;
  size_t m = n * K::NEPS;
  int rc = work_reserve(ctx, m);
  if (rc) return rc;
  if (wit) {
    if constexpr (KIND != PA_POK) {
      PA_LAUNCH(ctx, PA_K_POPS + KIND, (k_prove_ops_wit<KIND><<<grid_for(m), PA_BLOCK, 0, ctx->stream>>>(stmts, rnd, secrets, b0, b1, cb, ctx->d_comb, work_jac(ctx), (int)n, L)));
    }
  } else {
    PA_LAUNCH(ctx, PA_K_POPS + KIND, (k_prove_ops<KIND><<<grid_for(m), PA_BLOCK, 0, ctx->stream>>>(stmts, rnd, b0, b1, ctx->d_comb, work_jac(ctx), (int)n, L)));
  }
  if ((rc = normalize_to(ctx, proofs, m, K::NEPS, L.proof, L.inner, L.proof_in))) return rc;
  PA_LAUNCH(ctx, PA_K_PRESPOND + KIND, (k_prove_respond<KIND><<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(proofs, stmts, ids, secrets, rnd, b0, b1, (int)n, L)));
  return PA_OK;
}

}  // namespace

extern "C" {

// ---- device-pointer entry points -------------------------------------------------------
int pa_pokdlog_prove_dev(pa_ctx *ctx, const uint8_t *X, const uint8_t *x, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (X && x && ids && rnd && proofs)) && n < (1u << 26));
  return prove_dev<PA_POK>(ctx, X, x, nullptr, nullptr, (const u64 *)ids, rnd, proofs, n);
}
int pa_pokdlog_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *X, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && X && ids && verdict)) && n < (1u << 26));
  return verify_dev<PA_POK, 1>(ctx, proofs, X, (const u64 *)ids, verdict, n);
}
int pa_powfcom_prove_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *alpha, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && alpha && bits && ids && rnd && proofs)) && n < (1u << 26));
  return prove_dev<PA_COM>(ctx, stmt, alpha, bits, nullptr, (const u64 *)ids, rnd, proofs, n);
}
int pa_powfcom_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && stmt && ids && verdict)) && n < (1u << 26));
  return verify_dev<PA_COM, 4>(ctx, proofs, stmt, (const u64 *)ids, verdict, n);
}
int pa_stage1_prove_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bits && ids && rnd && proofs)) && n < (1u << 26));
  return prove_dev<PA_S1>(ctx, stmt, secrets, bits, nullptr, (const u64 *)ids, rnd, proofs, n);
}
int pa_stage1_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && stmt && ids && verdict)) && n < (1u << 26));
  return verify_dev<PA_S1, 8>(ctx, proofs, stmt, (const u64 *)ids, verdict, n);
}
int pa_stage2_prove_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bi && bj && ids && rnd && proofs)) && n < (1u << 26));
  return prove_dev<PA_S2>(ctx, stmt, secrets, bi, bj, (const u64 *)ids, rnd, proofs, n);
}
int pa_stage2_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && stmt && ids && verdict)) && n < (1u << 26));
  return verify_dev<PA_S2, 16>(ctx, proofs, stmt, (const u64 *)ids, verdict, n);
}

// provers with witnesses (pa_proof.cuh): extended secrets, same proofs
static pa_lay lay_w(pa_lay L, size_t secret_stride) {
  L.secret = secret_stride;
  return L;
}
int pa_powfcom_prove_w_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bits && ids && rnd && proofs)) && n < (1u << 26));
  return prove_dev<PA_COM>(ctx, stmt, secrets, bits, nullptr, (const u64 *)ids, rnd, proofs, n, lay_w(pa_lay_packed<PA_COM>(), 64), true);
}
int pa_stage1_prove_w_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bits && ids && rnd && proofs)) && n < (1u << 26));
  return prove_dev<PA_S1>(ctx, stmt, secrets, bits, nullptr, (const u64 *)ids, rnd, proofs, n, lay_w(pa_lay_packed<PA_S1>(), 128), true);
}
int pa_stage2_prove_w_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint8_t *cbit, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bi && bj && cbit && ids && rnd && proofs)) && n < (1u << 26));
  return prove_dev<PA_S2>(ctx, stmt, secrets, bi, bj, (const u64 *)ids, rnd, proofs, n, lay_w(pa_lay_packed<PA_S2>(), 192), true, cbit);
}

int pa_commit_points_dev(pa_ctx *ctx, const uint8_t *alpha, const uint8_t *beta, const uint8_t *bits, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (alpha && beta && bits && out)) && n < (1u << 26));
  if (n == 0) return PA_OK;
  int rc = work_reserve(ctx, 3 * n);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_COMMIT, (k_commit_points<<<grid_for(3 * n), PA_BLOCK, 0, ctx->stream>>>(alpha, beta, bits, ctx->d_comb, work_jac(ctx), (int)n)));
  return normalize_to(ctx, out, 3 * n);
}

int pa_y_scan_dev(pa_ctx *ctx, const uint8_t *X, uint8_t *Y, const uint32_t *offsets, size_t nseg, size_t npoints) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (npoints == 0 || (X && Y)) && npoints < (1u << 28) && nseg >= 1);
  if (npoints == 0) return PA_OK;
  int rc = work_reserve(ctx, npoints);
  if (rc) return rc;
  PA_LAUNCH(ctx, PA_K_YSCAN, (k_y_scan<<<(unsigned)nseg, PA_SCAN_T, 0, ctx->stream>>>(X, 64, offsets, (int)npoints, work_jac(ctx))));
  return normalize_to(ctx, Y, npoints);
}

int pa_point_sum_is_inf_dev(pa_ctx *ctx, const uint8_t *B, const uint32_t *offsets, size_t nseg, size_t npoints, int32_t *flags) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && flags && (npoints == 0 || B) && nseg >= 1);
  PA_LAUNCH(ctx, PA_K_SUMINF, (k_point_sum_is_inf<<<(unsigned)nseg, PA_SCAN_T, 0, ctx->stream>>>(B, 64, offsets, (int)npoints, flags)));
  return PA_OK;
}

int pa_challenge_dev(pa_ctx *ctx, const uint8_t *points, size_t k, const uint64_t *ids, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && k <= 32 && (n == 0 || (ids && out && (points || k == 0))));
  PA_ARGCHECK(ctx, aligned16(points) && aligned16(out));  // wire points and scalars are moved with 16-byte loads and stores
  if (n == 0) return PA_OK;
  PA_LAUNCH(ctx, PA_K_CHALLENGE, (k_challenge<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(points, (int)k, (const u64 *)ids, out, (int)n)));
  return PA_OK;
}

int pa_rng_fill256_dev(pa_ctx *ctx, uint64_t seed, const uint64_t *streams, uint64_t *counters, size_t per_item, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (streams && counters && out)));
  if (n == 0 || per_item == 0) return PA_OK;
  PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(seed, (const u64 *)streams, (u64 *)counters, nullptr, (int)per_item, out, (int)n, 1)));
  return PA_OK;
}

int pa_ccs22_setup_hash_dev(pa_ctx *ctx, const uint8_t *scalars, size_t k, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (scalars && out)) && k < (1u << 24));
  if (n == 0) return PA_OK;
  PA_LAUNCH(ctx, PA_K_CHALLENGE, (k_ccs22_setup_hash<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(scalars, (int)k, out, (int)n)));
  return PA_OK;
}

int pa_rng_fill_dev(pa_ctx *ctx, uint64_t seed, const uint64_t *streams, uint64_t *counters, size_t per_item, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (streams && counters && out)));
  if (n == 0 || per_item == 0) return PA_OK;
  PA_LAUNCH(ctx, PA_K_RNG, (k_rng_fill<<<grid_for(n), PA_BLOCK, 0, ctx->stream>>>(seed, (const u64 *)streams, (u64 *)counters, nullptr, (int)per_item, out, (int)n)));
  return PA_OK;
}

// ---- host-buffer entry points ---------------------------------------------------------------
int pa_pokdlog_prove(pa_ctx *ctx, const uint8_t *X, const uint8_t *x, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (X && x && ids && rnd && proofs)));
  HArg a[] = {{X, 0, n * 64}, {x, 0, n * 32}, {ids, 0, n * 8}, {rnd, 0, n * 32}, {0, proofs, n * 96}};
  return staged(ctx, a, 5, [&](unsigned char **d) { return pa_pokdlog_prove_dev(ctx, d[0], d[1], (const uint64_t *)d[2], d[3], d[4], n); });
}
int pa_pokdlog_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *X, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && X && ids && verdict)));
  HArg a[] = {{proofs, 0, n * 96}, {X, 0, n * 64}, {ids, 0, n * 8}, {0, verdict, n}};
  return staged(ctx, a, 4, [&](unsigned char **d) { return pa_pokdlog_verify_dev(ctx, d[0], d[1], (const uint64_t *)d[2], d[3], n); });
}
int pa_powfcom_prove(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *alpha, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && alpha && bits && ids && rnd && proofs)));
  HArg a[] = {{stmt, 0, n * 192}, {alpha, 0, n * 32}, {bits, 0, n}, {ids, 0, n * 8}, {rnd, 0, n * 96}, {0, proofs, n * 352}};
  return staged(ctx, a, 6, [&](unsigned char **d) { return pa_powfcom_prove_dev(ctx, d[0], d[1], d[2], (const uint64_t *)d[3], d[4], d[5], n); });
}
int pa_powfcom_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && stmt && ids && verdict)));
  HArg a[] = {{proofs, 0, n * 352}, {stmt, 0, n * 192}, {ids, 0, n * 8}, {0, verdict, n}};
  return staged(ctx, a, 4, [&](unsigned char **d) { return pa_powfcom_verify_dev(ctx, d[0], d[1], (const uint64_t *)d[2], d[3], n); });
}
int pa_stage1_prove(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bits && ids && rnd && proofs)));
  HArg a[] = {{stmt, 0, n * 448}, {secrets, 0, n * 64}, {bits, 0, n}, {ids, 0, n * 8}, {rnd, 0, n * 160}, {0, proofs, n * 672}};
  return staged(ctx, a, 6, [&](unsigned char **d) { return pa_stage1_prove_dev(ctx, d[0], d[1], d[2], (const uint64_t *)d[3], d[4], d[5], n); });
}
int pa_stage1_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && stmt && ids && verdict)));
  HArg a[] = {{proofs, 0, n * 672}, {stmt, 0, n * 448}, {ids, 0, n * 8}, {0, verdict, n}};
  return staged(ctx, a, 4, [&](unsigned char **d) { return pa_stage1_verify_dev(ctx, d[0], d[1], (const uint64_t *)d[2], d[3], n); });
}
int pa_stage2_prove(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bi && bj && ids && rnd && proofs)));
  for (size_t i = 0; i < n; ++i)
    if (bi[i] && !bj[i]) return pa_fail(ctx, PA_EINVAL, "stage 2: bi == 1 requires bj == 1 (assert at SEAL/bidder.cpp:604)");
  HArg a[] = {{stmt, 0, n * 704}, {secrets, 0, n * 96}, {bi, 0, n}, {bj, 0, n}, {ids, 0, n * 8}, {rnd, 0, n * 352}, {0, proofs, n * 1344}};
  return staged(ctx, a, 7, [&](unsigned char **d) { return pa_stage2_prove_dev(ctx, d[0], d[1], d[2], d[3], (const uint64_t *)d[4], d[5], d[6], n); });
}
int pa_powfcom_prove_w(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bits && ids && rnd && proofs)));
  HArg a[] = {{stmt, 0, n * 192}, {secrets, 0, n * 64}, {bits, 0, n}, {ids, 0, n * 8}, {rnd, 0, n * 96}, {0, proofs, n * 352}};
  return staged(ctx, a, 6, [&](unsigned char **d) { return pa_powfcom_prove_w_dev(ctx, d[0], d[1], d[2], (const uint64_t *)d[3], d[4], d[5], n); });
}
int pa_stage1_prove_w(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bits && ids && rnd && proofs)));
  HArg a[] = {{stmt, 0, n * 448}, {secrets, 0, n * 128}, {bits, 0, n}, {ids, 0, n * 8}, {rnd, 0, n * 160}, {0, proofs, n * 672}};
  return staged(ctx, a, 6, [&](unsigned char **d) { return pa_stage1_prove_w_dev(ctx, d[0], d[1], d[2], (const uint64_t *)d[3], d[4], d[5], n); });
}
int pa_stage2_prove_w(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint8_t *cbit, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (stmt && secrets && bi && bj && cbit && ids && rnd && proofs)));
  for (size_t i = 0; i < n; ++i)
    if (bi[i] && !bj[i]) return pa_fail(ctx, PA_EINVAL, "stage 2: bi == 1 requires bj == 1 (assert at SEAL/bidder.cpp:604)");
  HArg a[] = {{stmt, 0, n * 704}, {secrets, 0, n * 192}, {bi, 0, n}, {bj, 0, n}, {cbit, 0, n}, {ids, 0, n * 8}, {rnd, 0, n * 352}, {0, proofs, n * 1344}};
  return staged(ctx, a, 8, [&](unsigned char **d) { return pa_stage2_prove_w_dev(ctx, d[0], d[1], d[2], d[3], d[4], (const uint64_t *)d[5], d[6], d[7], n); });
}
int pa_stage2_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (proofs && stmt && ids && verdict)));
  HArg a[] = {{proofs, 0, n * 1344}, {stmt, 0, n * 704}, {ids, 0, n * 8}, {0, verdict, n}};
  return staged(ctx, a, 4, [&](unsigned char **d) { return pa_stage2_verify_dev(ctx, d[0], d[1], (const uint64_t *)d[2], d[3], n); });
}
int pa_commit_points(pa_ctx *ctx, const uint8_t *alpha, const uint8_t *beta, const uint8_t *bits, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (alpha && beta && bits && out)));
  HArg a[] = {{alpha, 0, n * 32}, {beta, 0, n * 32}, {bits, 0, n}, {0, out, n * 192}};
  return staged(ctx, a, 4, [&](unsigned char **d) { return pa_commit_points_dev(ctx, d[0], d[1], d[2], d[3], n); });
}
int pa_y_scan(pa_ctx *ctx, const uint8_t *X, uint8_t *Y, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (X && Y)));
  if (n == 0) return PA_OK;
  HArg a[] = {{X, 0, n * 64}, {0, Y, n * 64}};
  return staged(ctx, a, 2, [&](unsigned char **d) { return pa_y_scan_dev(ctx, d[0], d[1], nullptr, 1, n); });
}
int pa_y_scan_batch(pa_ctx *ctx, const uint8_t *X, uint8_t *Y, const uint32_t *offsets, size_t nseg) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && offsets && nseg >= 1);
  size_t n = offsets[nseg];
  PA_ARGCHECK(ctx, n == 0 || (X && Y));
  if (n == 0) return PA_OK;
  HArg a[] = {{X, 0, n * 64}, {0, Y, n * 64}, {offsets, 0, (nseg + 1) * 4}};
  return staged(ctx, a, 3, [&](unsigned char **d) { return pa_y_scan_dev(ctx, d[0], d[1], (const uint32_t *)d[2], nseg, n); });
}
int pa_point_sum_is_inf(pa_ctx *ctx, const uint8_t *B, size_t n, int *is_inf) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && is_inf && (n == 0 || B));
  int32_t flag = 0;
  HArg a[] = {{B, 0, n * 64}, {0, &flag, 4}};
  int rc = staged(ctx, a, 2, [&](unsigned char **d) { return pa_point_sum_is_inf_dev(ctx, d[0], nullptr, 1, n, (int32_t *)d[1]); });
  *is_inf = flag;
  return rc;
}
int pa_point_sum_is_inf_batch(pa_ctx *ctx, const uint8_t *B, const uint32_t *offsets, size_t nseg, int32_t *flags) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && offsets && flags && nseg >= 1);
  size_t n = offsets[nseg];
  HArg a[] = {{B, 0, n * 64}, {offsets, 0, (nseg + 1) * 4}, {0, flags, nseg * 4}};
  return staged(ctx, a, 3, [&](unsigned char **d) { return pa_point_sum_is_inf_dev(ctx, d[0], (const uint32_t *)d[1], nseg, n, (int32_t *)d[2]); });
}
int pa_challenge(pa_ctx *ctx, const uint8_t *points, size_t k, const uint64_t *ids, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && k <= 32 && (n == 0 || (ids && out)));
  HArg a[] = {{points, 0, n * k * 64}, {ids, 0, n * 8}, {0, out, n * 32}};
  return staged(ctx, a, 3, [&](unsigned char **d) { return pa_challenge_dev(ctx, d[0], k, (const uint64_t *)d[1], d[2], n); });
}
int pa_rng_fill256(pa_ctx *ctx, uint64_t seed, const uint64_t *streams, uint64_t *counters, size_t per_item, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (streams && counters && out)));
  HArg a[] = {{streams, 0, n * 8}, {counters, counters, n * 8}, {0, out, n * per_item * 32}};
  return staged(ctx, a, 3, [&](unsigned char **d) { return pa_rng_fill256_dev(ctx, seed, (const uint64_t *)d[0], (uint64_t *)d[1], per_item, d[2], n); });
}

int pa_ccs22_setup_hash(pa_ctx *ctx, const uint8_t *scalars, size_t k, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (scalars && out)));
  HArg a[] = {{scalars, 0, n * k * 32}, {0, out, n * 32}};
  return staged(ctx, a, 2, [&](unsigned char **d) { return pa_ccs22_setup_hash_dev(ctx, d[0], k, d[1], n); });
}

int pa_rng_fill(pa_ctx *ctx, uint64_t seed, const uint64_t *streams, uint64_t *counters, size_t per_item, uint8_t *out, size_t n) {
  PA_ENTER(ctx);
  PA_ARGCHECK(ctx, ctx && (n == 0 || (streams && counters && out)));
  HArg a[] = {{streams, 0, n * 8}, {counters, counters, n * 8}, {0, out, n * per_item * 32}};
  return staged(ctx, a, 3, [&](unsigned char **d) { return pa_rng_fill_dev(ctx, seed, (const uint64_t *)d[0], (uint64_t *)d[1], per_item, d[2], n); });
}

}  // extern "C"

#include "pa_seal.cuh"
#include "pa_ccs22.cuh"
