/*
 * pa_engine.h — C ABI of the B200-native engine for the AV-net / NIZK hot path
 * of Privacy-Auction (SEAL and CCS22 protocols on secp256k1).
 *
 * The reference has no FFI boundary of its own: its hot path is reached through
 * C++ member calls and bottoms out in OpenSSL's C API.  Every
 * entry point below therefore names the reference call (file:line under
 * /root/reference) whose work it replaces; INTEGRATION.md shows the binding a
 * maintainer of the reference would write on top of it.
 *
 * Conventions
 *   - plain pointers and sizes only; all functions return 0 on success or a
 *     negative PA_E* code, never throw; pa_last_error(ctx) gives a message;
 *   - point  = 64 bytes, affine X || Y, each 32-byte big-endian; the point at
 *              infinity is 64 zero bytes ((0,0) is not on the curve);
 *   - scalar = 32 bytes big-endian; any 256-bit value is accepted and reduced
 *              modulo the group order first (the reference passes unreduced
 *              scalars in CCS22);
 *   - id     = uint64_t, hashed as 8 little-endian bytes (SEAL/hash.cpp:40);
 *   - functions without a suffix take HOST buffers, copy them to the device,
 *     run on the context's stream, copy the result back and synchronise — the
 *     call a user of the reference would make;
 *   - functions ending in _dev take DEVICE pointers, are asynchronous on the
 *     context's stream and need pa_sync() before results are read;
 *   - there is no CPU implementation behind this ABI: without a CUDA device
 *     pa_ctx_create fails with PA_ENODEV.
 */
#ifndef PA_ENGINE_H
#define PA_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_OK 0
#define PA_EINVAL (-1)  /* bad argument */
#define PA_ENODEV (-2)  /* no CUDA device / wrong architecture */
#define PA_ECUDA (-3)   /* CUDA runtime error, see pa_last_error */
#define PA_ENOMEM (-4)

#define PA_POINT_BYTES 64
#define PA_SCALAR_BYTES 32

typedef struct pa_ctx pa_ctx;

/* ---- context ------------------------------------------------------------ */

/* Create an engine context on CUDA device `device` (one context per GPU per
 * process; the reference's analogue is EC_GROUP_new_by_curve_name(CURVE) +
 * BN_CTX_new(), SEAL/bidder.cpp:36, 1117).  Builds the fixed-base comb table
 * for the generator on the GPU. */
int pa_ctx_create(pa_ctx **out, int device);
int pa_ctx_destroy(pa_ctx *ctx);
/* wait for everything queued on the context's stream */
int pa_sync(pa_ctx *ctx);
/* message of the last error on this context (or of pa_ctx_create if ctx == NULL) */
const char *pa_last_error(pa_ctx *ctx);
/* ABI version, bumped on incompatible change */
int pa_abi_version(void);
/* the CUDA stream (cudaStream_t) the context launches on, for callers that
 * time with CUDA events or enqueue their own copies */
void *pa_ctx_stream(pa_ctx *ctx);
/* number of kernel launches issued through this context so far */
uint64_t pa_ctx_launches(pa_ctx *ctx);

/* device memory helpers for the _dev entry points (cudaMalloc / cudaMemcpyAsync
 * on the context's stream) so that a host program needs no CUDA headers */
int pa_dev_alloc(pa_ctx *ctx, void **dptr, size_t bytes);
int pa_dev_free(pa_ctx *ctx, void *dptr);
int pa_dev_upload(pa_ctx *ctx, void *dptr, const void *host, size_t bytes);
int pa_dev_download(pa_ctx *ctx, void *host, const void *dptr, size_t bytes);

/* ---- scalar multiplication ------------------------------------------------
 * The reference spends > 99 % of its time in these three call shapes. */

/* out[i] = scalars[i] * G.      EC_POINT_mul(group, r, k, NULL, NULL, ctx)
 * SEAL/bidder.cpp:98, 1137-1138, 1215-1216; CCS22/bidder.cpp:67 */
int pa_fixed_base_mul(pa_ctx *ctx, const uint8_t *scalars, uint8_t *out, size_t n);
int pa_fixed_base_mul_dev(pa_ctx *ctx, const uint8_t *d_scalars, uint8_t *d_out, size_t n);

/* out[i] = scalars[i] * points[i].   EC_POINT_mul(group, r, NULL, P, k, ctx)
 * SEAL/bidder.cpp:129, 173, 1303, 1307; CCS22/bidder.cpp:142 */
int pa_var_base_mul(pa_ctx *ctx, const uint8_t *points, const uint8_t *scalars, uint8_t *out, size_t n);
int pa_var_base_mul_dev(pa_ctx *ctx, const uint8_t *d_points, const uint8_t *d_scalars, uint8_t *d_out, size_t n);

/* out[i] = a[i] * G + b[i] * points[i].   EC_POINT_mul(group, r, a, P, b, ctx)
 * SEAL/bidder.cpp:175, 664-666, 1135; CCS22/bidder.cpp:86, 173 */
int pa_double_mul(pa_ctx *ctx, const uint8_t *a, const uint8_t *points, const uint8_t *b, uint8_t *out, size_t n);
int pa_double_mul_dev(pa_ctx *ctx, const uint8_t *d_a, const uint8_t *d_points, const uint8_t *d_b, uint8_t *d_out, size_t n);

/* out[i] = a[i] * p[i] + b[i] * q[i]: the "P^a * Q^b" shape of every
 * verification check, two EC_POINT_mul + EC_POINT_add, SEAL/bidder.cpp:266-268 */
int pa_lincomb2(pa_ctx *ctx, const uint8_t *p, const uint8_t *a, const uint8_t *q, const uint8_t *b, uint8_t *out, size_t n);
int pa_lincomb2_dev(pa_ctx *ctx, const uint8_t *d_p, const uint8_t *d_a, const uint8_t *d_q, const uint8_t *d_b, uint8_t *d_out, size_t n);

/* out[i] = p[i] + q[i] (sub != 0: p[i] - q[i]).   EC_POINT_add / EC_POINT_invert,
 * SEAL/bidder.cpp:130, 178-180 */
int pa_point_add(pa_ctx *ctx, const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, int sub);

/* EC_POINT_point2oct, SEAL/hash.cpp:27-29, SEAL/bulletinBoard.cpp:277.
 * compressed == 0: 04 || X || Y (65 bytes); compressed != 0: 02/03 || X (33 bytes);
 * infinity: the single byte 00.  Each output slot is `stride` bytes (>= 65 or 33),
 * zero padded; lens[i] receives the encoded length. */
int pa_point_encode(pa_ctx *ctx, const uint8_t *points, size_t n, int compressed, uint8_t *out, size_t stride, uint32_t *lens);

/* ---- measurement ----------------------------------------------------------
 * Per-kernel device timing: between pa_profile_begin and pa_profile_end every
 * kernel the context launches is bracketed by CUDA events on the context's
 * stream; pa_profile_end synchronises and returns, per kernel name, the number
 * of launches and the summed duration.  bench.py uses it for the roofline of
 * the dominant kernel inside its timed region. */
typedef struct {
  char name[32];
  uint64_t launches;
  double total_ms;
} pa_kernel_stat;
int pa_profile_begin(pa_ctx *ctx);
int pa_profile_end(pa_ctx *ctx, pa_kernel_stat *out, size_t cap, size_t *count);

/*
 * Register-only integer-pipe microbenchmark (SURVEY.md §8d asks for the IMAD
 * peak to be measured on the box).  out[0] = 32-bit IMAD / s, out[1] = 32x32+64
 * IMAD.WIDE / s, out[2] = field multiplications / s, out[3] = field squarings / s,
 * all whole-GPU, timed with CUDA events on the context's stream. */
int pa_measure_int_peak(pa_ctx *ctx, double out[4]);

#ifdef __cplusplus
}
#endif
#endif
