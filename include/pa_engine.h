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
/* sizeof of the structs that cross the ABI, so that a binding can check its own layout:
 * which = 0: pa_seal_job, 1: pa_ccs22_job, 2: pa_kernel_stat; anything else: 0 */
size_t pa_abi_sizeof(int which);
/* the CUDA stream (cudaStream_t) the context launches on, for callers that
 * time with CUDA events or enqueue their own copies */
void *pa_ctx_stream(pa_ctx *ctx);
/* number of kernel launches issued through this context so far */
uint64_t pa_ctx_launches(pa_ctx *ctx);

/* Entropy.  The draw stream that replaces BN_rand_range (see "seeded randomness" below) is a function of a
 * 64-bit seed only, which is what tests and benchmarks need and what no deployment may use: whoever knows
 * the seed knows every secret.  pa_ctx_set_entropy installs a 32-byte key (from getrandom(2), per process,
 * never published) that is mixed into every draw on this device:
 *   draw = SHA-256("PAv2" || key || LE64 seed || LE64 stream || LE64 ctr).
 * key32 == NULL goes back to the seeded test stream.  The reference draws from OpenSSL's DRBG
 * (BN_rand_range, SEAL/bidder.cpp:97) and std::random_device (SEAL/bidder.cpp:27); the host classes and the
 * CLIs of this repository install a key unless --seed is given. */
int pa_ctx_set_entropy(pa_ctx *ctx, const uint8_t *key32);

/* TEST HOOKS (never needed by a caller of the reference's path):
 *   PA_DBG_REJECT_BITS  a = k: a draw whose top k bits are all ones is redrawn as if it were >= the group
 *                       order (probability 2^-k instead of 2^-128), to exercise the redraw logic; 0 = off
 *   PA_DBG_CORRUPT      a = section (0 off, 1 commitment record, 2 round-one record, 3 round-two proof),
 *                       b = step << 32 | local bidder (section 1: step = bit index), c = byte offset in the
 *                       record: pa_seal_run flips the lowest bit of that byte between proving and verifying */
enum { PA_DBG_REJECT_BITS = 1, PA_DBG_CORRUPT = 2 };
int pa_debug_set(pa_ctx *ctx, int what, uint64_t a, uint64_t b, uint64_t c);
/* how many times pa_seal_run ran an auction again step-major because a draw was rejected */
uint64_t pa_ctx_reruns(pa_ctx *ctx);

/* Peer exchange window (one process per GPU, all GPUs of one node): lets the kernels of a bidder-sharded
 * auction exchange their per-step sums by writing into each other's HBM over NVLink instead of returning
 * to the host for a collective call per step (pa_seal_job.use_xchg).
 *   pa_xchg_create   allocates this rank's window (PA_XCHG_BYTES) and returns its 64-byte CUDA IPC handle;
 *   pa_xchg_connect  handles = world x 64 bytes, every rank's handle in rank order (exchange them with
 *                    whatever the host program has: MPI_Allgather, torch.distributed, a file);
 *   pa_xchg_skip     a rank that owns no bidder of a sharded auction calls this instead of pa_seal_run
 *                    (keeps the run counter the ranks share in step);
 *   pa_xchg_close    unmaps and frees (also done by pa_ctx_destroy). */
#define PA_XCHG_BYTES ((size_t)8 << 20)
#define PA_XCHG_MAX_WORLD 16
int pa_xchg_create(pa_ctx *ctx, uint8_t *handle64);
int pa_xchg_connect(pa_ctx *ctx, const uint8_t *handles, int world, int rank);
int pa_xchg_skip(pa_ctx *ctx);
int pa_xchg_close(pa_ctx *ctx);

/* device memory helpers for the _dev entry points (cudaMalloc / cudaMemcpyAsync
 * on the context's stream) so that a host program needs no CUDA headers.  Every device pointer
 * handed to a _dev entry point must be 16-byte aligned (what pa_dev_alloc returns is; record strides are
 * multiples of 32): points and scalars are moved with 16-byte loads and stores.  PA_EINVAL otherwise where
 * the entry point can tell. */
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

/* Several scalar-multiplication batches in ONE call (host buffers): what a caller gets from issuing pa_fixed_base_mul /
 * pa_var_base_mul / pa_double_mul / pa_lincomb2 one after the other (e.g. the g^x of Bidder::roundOne, SEAL/bidder.cpp:
 * 1217-1218, and the Y^x of roundTwo, :1303), with the chunks of the jobs interleaved in one copy/compute pipeline: the
 * copies of a copy-heavy job (fixed base: 96 bytes per 2 us of GPU time) hide behind the kernels of a compute-heavy one
 * (variable base).  On a host whose memory path is shared by 8 GPUs this is what keeps the end-to-end rate at the kernels'
 * (profiles/r02f_e2e_probe8.txt).  kind FIXED: out = a*G; VAR: out = a*p; DOUBLE: out = a*G + b*p; LINCOMB2: out = a*p + b*q. */
typedef struct {
  int kind;
  const uint8_t *a, *p, *b, *q;   /* scalars 32 B each, points 64 B each; unused ones NULL */
  uint8_t *out;                   /* n x 64 B */
  size_t n;
} pa_mul_job;
enum { PA_MUL_FIXED = 0, PA_MUL_VAR = 1, PA_MUL_DOUBLE = 2, PA_MUL_LINCOMB2 = 3 };
int pa_scalar_mul_jobs(pa_ctx *ctx, const pa_mul_job *jobs, size_t njobs);

/* out[i] = p[i] + q[i] (sub != 0: p[i] - q[i]).   EC_POINT_add / EC_POINT_invert,
 * SEAL/bidder.cpp:130, 178-180 */
int pa_point_add(pa_ctx *ctx, const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, int sub);

/* ok[i] = 1 if points[i] is the point at infinity or a valid curve point (both coordinates < p and
 * y^2 = x^3 + 7): what EC_POINT_set_affine_coordinates checks when a point enters libcrypto.  The four
 * proof VERIFIERS apply this test to every point of the proof and of the statement themselves (a proof
 * with an off-curve or non-canonical point gets verdict 0); the scalar-multiplication entry points do
 * not validate their inputs (the reference only ever feeds them points it computed itself). */
int pa_point_on_curve(pa_ctx *ctx, const uint8_t *points, size_t n, uint8_t *ok);

/* EC_POINT_point2oct, SEAL/hash.cpp:27-29, SEAL/bulletinBoard.cpp:277.
 * compressed == 0: 04 || X || Y (65 bytes); compressed != 0: 02/03 || X (33 bytes);
 * infinity: the single byte 00.  Each output slot is `stride` bytes (>= 65 or 33),
 * zero padded; lens[i] receives the encoded length. */
int pa_point_encode(pa_ctx *ctx, const uint8_t *points, size_t n, int compressed, uint8_t *out, size_t stride, uint32_t *lens);

/* ---- Fiat-Shamir challenge ---------------------------------------------------
 * out[i] = SHA-256( enc(g) || enc(points[i][0]) .. enc(points[i][k-1]) || LE64(ids[i]) )
 * read big-endian, mod the group order; enc = 04||X||Y, or the single byte 00
 * for infinity.  The generator is prepended by the engine, as the reference's
 * points[] arrays do.  k <= 32.   SHA256inNIZKPoKDLog / PoWFCom / PoWFStage1 /
 * PoWFStage2, SEAL/hash.cpp:8-53, 55-104, 106-162, 164-228 */
int pa_challenge(pa_ctx *ctx, const uint8_t *points, size_t k, const uint64_t *ids, uint8_t *out, size_t n);
int pa_challenge_dev(pa_ctx *ctx, const uint8_t *d_points, size_t k, const uint64_t *d_ids, uint8_t *d_out, size_t n);

/* ---- the four NIZK proofs, batched ----------------------------------------------
 * Record layouts (points 64 B, scalars 32 B, field order of SEAL/types.h:13-93):
 *   PoKDLog   eps | rho                                               96 B
 *   PoWFCom   eps11 eps12 eps21 eps22 | rho1 rho2 ch2                352 B
 *   Stage1    eps11..eps14 eps21..eps24 | rho11 rho12 rho21 rho22 ch2   672 B
 *   Stage2    eps11 eps12 eps13 eps11' eps12' eps13' eps21 eps22 eps23 eps21' eps22'
 *             eps23' eps31 eps32 eps31' eps32' | rho11 rho12 rho13 rho21 rho22 rho23
 *             rho31 rho32 ch2 ch3                                   1344 B
 * Statement points are n x k contiguous points in the order of the reference's
 * parameter lists.  A prover takes the values the reference would draw with
 * BN_rand_range as an explicit array `rnd` (n x draws x 32 B, in the reference's
 * draw order), which makes it a pure function of its
 * inputs; pa_rng_fill produces such arrays from a seed.  A verifier writes one
 * byte per proof (1 = every check holds); like the reference it evaluates all
 * checks, no early exit (SEAL/bidder.cpp:244-298).  Every point a verifier receives (eps and
 * statement) must be the 64 zero bytes of infinity or canonical coordinates on the curve, as
 * EC_POINT_set_affine_coordinates would demand of a received point: otherwise the verdict is 0. */

/* genNIZKPoKDLog SEAL/bidder.cpp:90-107; X = g^x; rnd: v */
int pa_pokdlog_prove(pa_ctx *ctx, const uint8_t *X, const uint8_t *x, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_pokdlog_prove_dev(pa_ctx *ctx, const uint8_t *X, const uint8_t *x, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
/* verNIZKPoKDLog SEAL/bidder.cpp:119-136 */
int pa_pokdlog_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *X, const uint64_t *ids, uint8_t *verdict, size_t n);
int pa_pokdlog_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *X, const uint64_t *ids, uint8_t *verdict, size_t n);

/* genNIZKPoWFCom SEAL/bidder.cpp:149-226; stmt = (phi, A, B); rnd: r1, then
 * bit 0: ch2, rho2 / bit 1: ch1, rho1 */
int pa_powfcom_prove(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *alpha, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_powfcom_prove_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *alpha, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
/* verNIZKPoWFCom SEAL/bidder.cpp:241-299 */
int pa_powfcom_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);
int pa_powfcom_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);

/* genNIZKPoWFStage1 SEAL/bidder.cpp:318-451; stmt = (b, X, Y, R, c, A, B);
 * secrets = (x, alpha); rnd: r11, r12, then bit 0: rho21, rho22, ch2 / bit 1: rho11, rho12, ch1 */
int pa_stage1_prove(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_stage1_prove_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
/* verNIZKPoWFStage1 SEAL/bidder.cpp:470-571 */
int pa_stage1_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);
int pa_stage1_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);

/* genNIZKPoWFStage2 SEAL/bidder.cpp:598-890; stmt = (Bi, Xi, Ri, Bj, Xj, Rj, Ci, A, B, Yi, Yj);
 * secrets = (xi, xj, alpha); bi/bj as the reference's int arguments (bi == 1 requires
 * bj == 1, the assert at :604); rnd: the 11 draws of :643-655 / :692-699 / :749-756,
 * including the ones the reference discards */
int pa_stage2_prove(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_stage2_prove_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
/* verNIZKPoWFStage2 SEAL/bidder.cpp:913-1101 */
int pa_stage2_verify(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);
int pa_stage2_verify_dev(pa_ctx *ctx, const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);

/* Provers WITH WITNESSES.  A prover knows the discrete logarithm of almost every point of its own statement
 * (X = g^x, R = g^r, A = g^alpha, B = g^beta, c = g^(alpha beta + bit); its cryptogram is g^(r x) when it vetoes
 * and Y^x when it does not), so each published point a*P + b*Q is ONE fixed-base multiplication plus at most one
 * variable-base multiplication of the foreign point Y, instead of what EC_POINT_mul on the bases as given costs
 * (SEAL/bidder.cpp:171-202, 355-417, 657-814).  Same group elements, hence byte-identical proofs; 2-3 times
 * less work per proof.  Extended secrets per proof (32 B each):
 *   pa_powfcom_prove_w  (alpha, beta)
 *   pa_stage1_prove_w   (x, alpha, r, beta)               bits[i] = the committed bit = "the bidder vetoes"
 *   pa_stage2_prove_w   (xi, xj, alpha, ri, rj, beta)     bi / bj as in pa_stage2_prove (Bi = Ri^xi iff bi, Bj = Rj^xj iff bj),
 *                                                         cbit[i] = the bit committed to in Ci
 * The statement must be consistent with these secrets (it is when it was built from them); the plain provers above
 * make no such assumption. */
int pa_powfcom_prove_w(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_powfcom_prove_w_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_stage1_prove_w(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_stage1_prove_w_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_stage2_prove_w(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint8_t *cbit, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int pa_stage2_prove_w_dev(pa_ctx *ctx, const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj, const uint8_t *cbit, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);

/* ---- round logic -----------------------------------------------------------------------
 * out[i] = (phi, A, B) = (g^(alpha*beta) * g^bit, g^alpha, g^beta).  Bidder::commitBid,
 * SEAL/bidder.cpp:1131-1138 */
int pa_commit_points(pa_ctx *ctx, const uint8_t *alpha, const uint8_t *beta, const uint8_t *bits, uint8_t *out, size_t n);
int pa_commit_points_dev(pa_ctx *ctx, const uint8_t *alpha, const uint8_t *beta, const uint8_t *bits, uint8_t *out, size_t n);

/* Y[id] = sum_{i<id} X[i] - sum_{i>id} X[i] for every id of one auction, in one
 * prefix/suffix scan.  Bidder::roundTwo, SEAL/bidder.cpp:1286-1299 (there: O(n^2)
 * additions, repeated by every bidder).  _batch / _dev: many auctions at once,
 * auction s owns points offsets[s] .. offsets[s+1] (offsets == NULL: one auction). */
int pa_y_scan(pa_ctx *ctx, const uint8_t *X, uint8_t *Y, size_t n);
int pa_y_scan_batch(pa_ctx *ctx, const uint8_t *X, uint8_t *Y, const uint32_t *offsets, size_t nseg);
int pa_y_scan_dev(pa_ctx *ctx, const uint8_t *d_X, uint8_t *d_Y, const uint32_t *d_offsets, size_t nseg, size_t npoints);

/* *is_inf = (sum_i B[i] is the point at infinity).  Bidder::roundThree,
 * SEAL/bidder.cpp:1393-1397 */
int pa_point_sum_is_inf(pa_ctx *ctx, const uint8_t *B, size_t n, int *is_inf);
int pa_point_sum_is_inf_batch(pa_ctx *ctx, const uint8_t *B, const uint32_t *offsets, size_t nseg, int32_t *flags);
int pa_point_sum_is_inf_dev(pa_ctx *ctx, const uint8_t *d_B, const uint32_t *d_offsets, size_t nseg, size_t npoints, int32_t *d_flags);

/* ---- CCS22 -------------------------------------------------------------------------------
 * The CCS22 protocol's curve work (CCS22/bidder.cpp:48-198, CCS22/evaluator.cpp:22-156) is made
 * of the call shapes above: X_i = g^x (pa_fixed_base_mul); Com = g^bid g1^H + h^R (pa_double_mul,
 * pa_var_base_mul, pa_point_add); own Y (pa_y_scan); B = Y^x | g^r; T2 = g^k, G = g^beta g1^alpha
 * (pa_double_mul), H = T2^alpha + h^beta (pa_lincomb2); z = g^s h^t, C0 = G^s + H^t + B,
 * C1 = (G - g1)^s + (H - T2)^t + M1 (pa_lincomb2 + pa_point_add); sum_j (C0_j - beta_j z_j) + B
 * (pa_var_base_mul, pa_point_add, pa_point_sum_is_inf).  The one operation of its own is the
 * setup hash:
 *   out[i] = SHA-256(minimal big-endian bytes of scalars[i][0..k)) mod order; a zero scalar takes
 *   the reference's error path and yields 0.   SHA256inSetup, CCS22/hash.cpp:9-57 */
int pa_ccs22_setup_hash(pa_ctx *ctx, const uint8_t *scalars, size_t k, uint8_t *out, size_t n);
int pa_ccs22_setup_hash_dev(pa_ctx *ctx, const uint8_t *d_scalars, size_t k, uint8_t *d_out, size_t n);

/* Fused oblivious-transfer messages: the 3-6 scalar multiplications of one message run
 * concurrently (one warp each) instead of one after the other.  params[i] = (g1, h) of item i.
 *   pa_ccs22_ot_recv1: (k, beta, alpha) -> (T2, G, H) = (g^k, g^beta g1^alpha, T2^alpha h^beta)
 *                      Evaluator::OTReceive1, CCS22/evaluator.cpp:91-111
 *   pa_ccs22_ot_send:  r1 = (T2, G, H), B, st = (s, t), m -> (z, C0, C1) = (g^s h^t, G^s H^t B,
 *                      (G/g1)^s (H/T2)^t g^m)        Bidder::OTSend, CCS22/bidder.cpp:155-198 */
int pa_ccs22_ot_recv1(pa_ctx *ctx, const uint8_t *k, const uint8_t *beta, const uint8_t *alpha, const uint8_t *params, uint8_t *out, size_t n);
int pa_ccs22_ot_recv1_dev(pa_ctx *ctx, const uint8_t *k, const uint8_t *beta, const uint8_t *alpha, const uint8_t *params, uint8_t *out, size_t n);
int pa_ccs22_ot_send(pa_ctx *ctx, const uint8_t *r1, const uint8_t *params, const uint8_t *B, const uint8_t *st, const uint8_t *m, uint8_t *out, size_t n);
int pa_ccs22_ot_send_dev(pa_ctx *ctx, const uint8_t *r1, const uint8_t *params, const uint8_t *B, const uint8_t *st, const uint8_t *m, uint8_t *out, size_t n);

/* The other three per-party steps of CCS22 as one call each (SURVEY.md 8b lower seam).
 *   pa_ccs22_commit:     H = SHA256inSetup(k scalars) and Com = g^bid * g1^H + h^R, params[i] = (g1, h)
 *                        Bidder::setupInner, CCS22/bidder.cpp:80-88; Evaluator::setupInner, evaluator.cpp:54-62
 *   pa_ccs22_bes_encode: X = the n public keys of this step; for each of m parties ids[i] of that auction
 *                        B = Y_ids[i]^x (d = 0) or g^r (d = 1)        Bidder::BESEncodeInner, CCS22/bidder.cpp:118-147
 *   pa_ccs22_ot_recv2:   is_inf = (sum_j (C0_j - beta_j z_j) + B == infinity) over the n OT_S records (z, C0, C1)
 *                        Evaluator::OTReceive2, CCS22/evaluator.cpp:117-156 */
int pa_ccs22_commit(pa_ctx *ctx, const uint8_t *scalars, size_t k, const uint8_t *bid, const uint8_t *R, const uint8_t *params, uint8_t *out_H, uint8_t *out_com, size_t n);
int pa_ccs22_commit_dev(pa_ctx *ctx, const uint8_t *scalars, size_t k, const uint8_t *bid, const uint8_t *R, const uint8_t *params, uint8_t *out_H, uint8_t *out_com, size_t n);
int pa_ccs22_bes_encode(pa_ctx *ctx, const uint8_t *X, size_t n, const uint64_t *ids, const uint8_t *d, const uint8_t *x, const uint8_t *r, uint8_t *out, size_t m);
int pa_ccs22_bes_encode_dev(pa_ctx *ctx, const uint8_t *X, size_t n, const uint64_t *ids, const uint8_t *d, const uint8_t *x, const uint8_t *r, uint8_t *out, size_t m);
int pa_ccs22_ot_recv2(pa_ctx *ctx, const uint8_t *ots, const uint8_t *beta, const uint8_t *B, size_t n, int *is_inf);
int pa_ccs22_ot_recv2_dev(pa_ctx *ctx, const uint8_t *ots, const uint8_t *beta, const uint8_t *B, size_t n, int32_t *d_is_inf);

/* ---- seeded randomness ------------------------------------------------------------------
 * The reference draws from OpenSSL's DRBG and is not reproducible.
 * The PA stream replaces BN_rand_range(., order) (SEAL/bidder.cpp:97 and 44 more sites):
 *   draw(seed, stream, ctr) = SHA-256("PAv1" || LE64 seed || LE64 stream || LE64 ctr)
 * as a big-endian integer, redrawn with ctr+1 while >= order.  Item i receives
 * per_item consecutive draws of stream streams[i] starting at counters[i];
 * counters[i] is advanced. */
int pa_rng_fill(pa_ctx *ctx, uint64_t seed, const uint64_t *streams, uint64_t *counters, size_t per_item, uint8_t *out, size_t n);
int pa_rng_fill_dev(pa_ctx *ctx, uint64_t seed, const uint64_t *d_streams, uint64_t *d_counters, size_t per_item, uint8_t *d_out, size_t n);
/* same stream with BN_rand(., 256, -1, 0) semantics: the 256-bit value unreduced, never redrawn
 * (CCS22/bidder.cpp:170, CCS22/evaluator.cpp:97, CCS22/bulletinBoard.cpp:33, 45) */
int pa_rng_fill256(pa_ctx *ctx, uint64_t seed, const uint64_t *streams, uint64_t *counters, size_t per_item, uint8_t *out, size_t n);
int pa_rng_fill256_dev(pa_ctx *ctx, uint64_t seed, const uint64_t *d_streams, uint64_t *d_counters, size_t per_item, uint8_t *d_out, size_t n);

/* ---- whole auctions ------------------------------------------------------------------------
 * pa_seal_run advances every bidder of a batch of SEAL auctions in lock step with all state
 * resident in HBM: what the reference's main does one bidder and one EC_POINT_mul at a time
 * (SEAL/main.cpp:32-120).  Every published proof is verified once (the reference lets each
 * of the n bidders repeat the same deterministic checks); verdicts are the same.
 *
 * Bidder j of auction a draws from PA stream (seed, (auction_id[a] << 32) | j) in the
 * reference's draw order, so the output is a function of
 * (seed, n, c, bids) only.
 *
 * Partitioning over GPUs (one process per GPU):
 *   - independent auctions: give each rank its own auctions; no exchange;
 *   - ONE auction sharded by bidder slice: every rank passes the same n / c, its own id
 *     range [lo, hi) and the bids of that range, plus an all-gather callback.  In the step-major
 *     schedule, once per step the X_i and once the b_i of a slice (slice * 64 bytes, zero padded)
 *     are placed in d_send; allgather(user, which) must leave the concatenation of the first
 *     so-many bytes of all ranks' d_send, in rank order, in d_recv (ranks own ascending id ranges
 *     of `slice` bidders each) and return 0 when d_recv is ready.  (See `schedule` below for the
 *     exchanges of the phase-major schedule.)  Each rank proves and verifies its own slice.
 *
 * Optional host outputs ("sections", NULL to skip), m = local bidders, slot = position of a
 * local bidder (auction-major, id order), Mb = sum over local bidders of c:
 *   out_commit    [Mb x 736]  CommitmentPerBit records, bidder-major      out_commit_ok [Mb]
 *   out_r1        [cmax x m x 320]  RoundOnePub records                  out_r1_ok  [cmax x m]
 *   out_r2_tag    [cmax x m]  1 = stage 1, 2 = stage 2, 0 = auction over  out_r2_ok  [cmax x m]
 *   out_r2_b      [cmax x m x 64]   out_r2_proof [cmax x m x 1344] (stage 1 uses 672)
 *   out_r3        [cmax x n_auctions]  1 = deciding step
 * tests/seal_flow.py turns sections into the PASEALT1 transcript. */
typedef int (*pa_allgather_fn)(void *user, int which /* 0: X of round one, 1: b of round two (slice * 64 bytes each); 2, 3: see `schedule` */);
typedef struct {
  uint64_t seed;
  size_t n_auctions;
  const uint32_t *n;            /* [n_auctions] bidders per auction */
  const uint32_t *c;            /* [n_auctions] bits per bid (<= 64) */
  const uint64_t *auction_ids;  /* [n_auctions] or NULL for 0 .. n_auctions-1 */
  const uint64_t *bids;         /* bids of the local bidders, auction-major, id order */
  int verify;                   /* 0: skip verification, 1: verify every proof once, k > 1: k times each (k = n - 1
                                   is the work of the reference's all-pairs verification; same verdicts) */
  /* sharding of one auction (allgather != NULL requires n_auctions == 1) */
  uint32_t lo, hi, slice;
  pa_allgather_fn allgather;
  void *user;
  uint8_t *d_send, *d_recv;     /* device buffers: slice * 64 and world * slice * 64 bytes */
  /* outputs */
  uint64_t *max_bid;            /* [n_auctions] */
  uint8_t *ok;                  /* [n_auctions] 1 = every local verification held */
  uint8_t *out_commit, *out_commit_ok, *out_r1, *out_r1_ok, *out_r2_tag, *out_r2_b, *out_r2_proof, *out_r2_ok, *out_r3;
  /* How the steps are scheduled; the results are the same bytes either way.
   *   step-major:  one batched kernel sequence per protocol step (any batch, any sharding);
   *   phase-major: ONE unsharded auction - keys, Y and both cryptogram candidates of every step in
   *                three large launches, the step-by-step decisions by one thread block on the
   *                device, then all proofs of all steps in one batch per kind (a single auction is
   *                otherwise a chain of lone-warp latencies, ~2 ms per step whatever n is).
   *                A sharded auction can use it too when the exchange buffers are large enough
   *                (xchg_bytes >= max(c * slice * 64, 128)): allgather(user, 2) then exchanges the
   *                first c * slice * 64 bytes of d_send (the X of all steps) and allgather(user, 3)
   *                the first 128 bytes (a rank's partial sum of one step's cryptograms), each into
   *                d_recv at that stride; one `2` per pass (at most two) and one `3` per step.
   * 0 picks phase-major where it applies. */
  int schedule;
  size_t xchg_bytes;            /* capacity of d_send; d_recv holds world times as much (0: slice * 64) */
  /* use_xchg != 0: the sharded auction exchanges through the peer window of this context (pa_xchg_connect)
   * instead of the callback: allgather, user, d_send, d_recv are ignored, nothing returns to the host per
   * step.  Rank r must own bidders [r * slice, min(n, (r + 1) * slice)); ranks beyond ceil(n / slice) call
   * pa_xchg_skip.  Phase-major: the Y reconstruction is sharded too (only the ranks' sums of public keys and
   * of cryptograms cross the link, 96 bytes per rank and step). */
  int use_xchg;
  uint8_t *ok_all;              /* use_xchg: 1 = every verification on every rank held (NULL to skip) */
} pa_seal_job;
enum { PA_SEAL_AUTO = 0, PA_SEAL_STEP_MAJOR = 1, PA_SEAL_PHASE_MAJOR = 2 };
int pa_seal_run(pa_ctx *ctx, const pa_seal_job *job);

/* pa_ccs22_run: the same for the CCS22 protocol (CCS22/main.cpp:16-130): every party of a batch of
 * auctions in lock step, no zero-knowledge proofs (the reference's verification phase is a TODO,
 * CCS22/main.cpp:132-134).  Party i of auction a draws from PA stream (seed, (auction_id << 32) | i),
 * the bulletin board's g1, h from (seed, (auction_id << 32) | 0xFFFFFFFF).  Independent auctions only
 * (partition them over GPUs; no exchange).
 * m = sum n, Mb = sum n*c, ms = sum (n-1) ("slots": the non-evaluator bidders, auction-major in id order).
 * Optional host outputs: out_params [A x 128] (g1, h), out_com [m x 64], out_pub [Mb x 64] (party-major),
 *   out_r1 [cmax x ms x 192] (T2, G, H), out_ots [cmax x ms x 192] (z, C0, C1), out_d [cmax x A]. */
typedef struct {
  uint64_t seed;
  size_t n_auctions;
  const uint32_t *n, *c, *evaluator; /* [n_auctions] parties, bits, id of the evaluator */
  const uint64_t *auction_ids;       /* [n_auctions] or NULL */
  const uint64_t *bids;              /* all parties, auction-major, id order */
  uint64_t *max_bid;                 /* [m] the maximum each party computed */
  uint8_t *out_params, *out_com, *out_pub, *out_r1, *out_ots, *out_d;
  /* PA_SEAL_AUTO / PA_SEAL_STEP_MAJOR / PA_SEAL_PHASE_MAJOR as in pa_seal_job: one auction can be run
   * phase-major (every candidate of every step in a few large launches, one block walks the steps,
   * one kernel assembles the published records); same bytes. */
  int schedule;
} pa_ccs22_job;
int pa_ccs22_run(pa_ctx *ctx, const pa_ccs22_job *job);

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
 * Register-only integer-pipe microbenchmarks (SURVEY.md section 8d asks for the IMAD peak to be measured on
 * the box), whole-GPU rates timed with CUDA events on the context's stream:
 *   out[0]  32-bit IMAD / s
 *   out[1]  32 x 32 + 64 multiply-adds / s in the shape of the field multiplier: chains of four linked by
 *           the carry flag (mad.lo.cc / madc.hi.cc -> IMAD.WIDE.U32.X), nothing else in the loop
 *   out[2]  field multiplications / s      out[3]  field squarings / s
 *   out[4]  32 x 32 + 64 multiply-adds / s without carry links (mad.wide.u32; ptxas turns each into an
 *           IMAD.WIDE.U32 with a zero addend plus a 64-bit addition on the other pipe)
 *   out[5]  reserved (0) */
int pa_measure_int_peak(pa_ctx *ctx, double out[6]);

#ifdef __cplusplus
}
#endif
#endif
