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

#ifdef __cplusplus
}
#endif
#endif
