/*
 * TEST INFRASTRUCTURE ("Tier-B oracle") — CPU restatement, on OpenSSL libcrypto,
 * of the reference's AV-net / NIZK hot path with the SAME batched signatures
 * and byte formats as the engine's C ABI (include/pa_engine.h), so that a parity
 * test is "call both on the same buffers, compare bytes".
 *
 * The arithmetic itself (EC_POINT_mul, EC_POINT_add, BN_mod_mul, SHA-256) is NOT
 * restated here: it is the very library the reference calls (OpenSSL libcrypto,
 * "3.2.0" per reference README.md:13; 3.0.13 in this image — results are
 * mathematically determined).  What is restated is everything the
 * reference builds on top of it, each function citing the lines it follows.
 * An independent restatement of the group law itself lives in
 * oracle/secp256k1_py.py and is checked against this file in tests/.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against transcripts
 * produced by the UNMODIFIED reference (oracle/_ref/seal_ref, Tier A) that are
 * committed under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product never does.
 */
#ifndef PA_ORACLE_H
#define PA_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* curve constants straight from libcrypto: p, n, Gx, Gy as 32-byte big-endian */
int po_curve_constants(uint8_t p[32], uint8_t n[32], uint8_t gx[32], uint8_t gy[32]);

/* EC_POINT_mul call shapes, SEAL/bidder.cpp:98, 129, 175 */
int po_fixed_base_mul(const uint8_t *scalars, uint8_t *out, size_t n);
int po_var_base_mul(const uint8_t *points, const uint8_t *scalars, uint8_t *out, size_t n);
int po_double_mul(const uint8_t *a, const uint8_t *points, const uint8_t *b, uint8_t *out, size_t n);
/* two EC_POINT_mul + EC_POINT_add, SEAL/bidder.cpp:266-268 */
int po_lincomb2(const uint8_t *p, const uint8_t *a, const uint8_t *q, const uint8_t *b, uint8_t *out, size_t n);
/* EC_POINT_add / EC_POINT_invert, SEAL/bidder.cpp:130, 178-180 */
int po_point_add(const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, int sub);
/* EC_POINT_point2oct, SEAL/hash.cpp:27-29 */
int po_point_encode(const uint8_t *points, size_t n, int compressed, uint8_t *out, size_t stride, uint32_t *lens);


/* ---- Fiat-Shamir challenge, SEAL/hash.cpp:8-228: points WITHOUT the generator
 * (it is prepended, as the reference's points[] arrays do); k points per item */
int po_challenge(const uint8_t *points, size_t k, const uint64_t *ids, uint8_t *out, size_t n);

/* ---- the four NIZK proofs.  Layouts and argument order as in include/pa_engine.h.
 * rnd holds the values the reference would draw with BN_rand_range, in draw
 * order, so a prover is a pure function. */
int po_pokdlog_prove(const uint8_t *X, const uint8_t *x, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int po_pokdlog_verify(const uint8_t *proofs, const uint8_t *X, const uint64_t *ids, uint8_t *verdict, size_t n);
int po_powfcom_prove(const uint8_t *stmt, const uint8_t *alpha, const uint8_t *bits, const uint64_t *ids,
                     const uint8_t *rnd, uint8_t *proofs, size_t n);
int po_powfcom_verify(const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);
int po_stage1_prove(const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids,
                    const uint8_t *rnd, uint8_t *proofs, size_t n);
int po_stage1_verify(const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);
int po_stage2_prove(const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj,
                    const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n);
int po_stage2_verify(const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n);

/* ---- round logic */
/* phi = g^(alpha*beta) * g^bit, A = g^alpha, B = g^beta; out = n x (phi, A, B).  SEAL/bidder.cpp:1131-1138 */
int po_commit_points(const uint8_t *alpha, const uint8_t *beta, const uint8_t *bits, uint8_t *out, size_t n);
/* Y_id = sum_{i<id} X_i - sum_{i>id} X_i for every id, the reference's O(n^2) loops.  SEAL/bidder.cpp:1286-1299 */
int po_y_scan(const uint8_t *X, uint8_t *Y, size_t n);
/* *is_inf = (sum_i b_i == infinity).  SEAL/bidder.cpp:1393-1397 */
int po_point_sum_is_inf(const uint8_t *b, size_t n, int *is_inf);


/* ---- CCS22 */
/* H = SHA-256(minimal big-endian bytes of k scalars) mod order; a zero scalar takes the reference's
 * error path and leaves H = 0.  SHA256inSetup, CCS22/hash.cpp:9-57 */
int po_ccs22_setup_hash(const uint8_t *scalars, size_t k, uint8_t *out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
