/*
 * TEST INFRASTRUCTURE ("Tier-B oracle") — see pa_oracle.h.  CPU only; never
 * linked into or called by the product.
 */
#include "pa_oracle.h"

#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/evp.h>
#include <openssl/obj_mac.h>
#include <openssl/sha.h>
#include <stdlib.h>
#include <string.h>

/* reference SEAL/params.h:4  #define CURVE 714 */
#define PO_CURVE NID_secp256k1

typedef struct {
  EC_GROUP *group;
  const EC_POINT *g;
  const BIGNUM *order;
  BN_CTX *ctx;
} po_env;

static int env_init(po_env *e) {
  e->group = EC_GROUP_new_by_curve_name(PO_CURVE); /* SEAL/bidder.cpp:36 */
  e->g = EC_GROUP_get0_generator(e->group);        /* SEAL/bidder.cpp:39 */
  e->order = EC_GROUP_get0_order(e->group);        /* SEAL/bidder.cpp:42 */
  e->ctx = BN_CTX_new();
  return e->group && e->g && e->order && e->ctx ? 0 : -1;
}
static void env_free(po_env *e) {
  BN_CTX_free(e->ctx);
  EC_GROUP_free(e->group);
}

/* wire form <-> EC_POINT: 64 bytes X||Y, zeros = infinity */
static EC_POINT *pt_in(po_env *e, const uint8_t *b) {
  EC_POINT *P = EC_POINT_new(e->group);
  int nz = 0;
  for (int i = 0; i < 64; ++i) nz |= b[i];
  if (!nz) {
    EC_POINT_set_to_infinity(e->group, P);
  } else {
    BIGNUM *x = BN_bin2bn(b, 32, NULL), *y = BN_bin2bn(b + 32, 32, NULL);
    if (EC_POINT_set_affine_coordinates(e->group, P, x, y, e->ctx) != 1) {
      EC_POINT_free(P);
      P = NULL;
    }
    BN_free(x);
    BN_free(y);
  }
  return P;
}
static void pt_out(po_env *e, uint8_t *b, const EC_POINT *P) {
  memset(b, 0, 64);
  if (!EC_POINT_is_at_infinity(e->group, P)) {
    BIGNUM *x = BN_new(), *y = BN_new();
    EC_POINT_get_affine_coordinates(e->group, P, x, y, e->ctx);
    BN_bn2binpad(x, b, 32);
    BN_bn2binpad(y, b + 32, 32);
    BN_free(x);
    BN_free(y);
  }
}
static BIGNUM *sc_in(const uint8_t *b) { return BN_bin2bn(b, 32, NULL); }

int po_curve_constants(uint8_t p[32], uint8_t n[32], uint8_t gx[32], uint8_t gy[32]) {
  po_env e;
  if (env_init(&e)) return -1;
  BIGNUM *bp = BN_new(), *a = BN_new(), *b = BN_new(), *x = BN_new(), *y = BN_new();
  EC_GROUP_get_curve(e.group, bp, a, b, e.ctx);
  EC_POINT_get_affine_coordinates(e.group, e.g, x, y, e.ctx);
  BN_bn2binpad(bp, p, 32);
  BN_bn2binpad(e.order, n, 32);
  BN_bn2binpad(x, gx, 32);
  BN_bn2binpad(y, gy, 32);
  BN_free(bp); BN_free(a); BN_free(b); BN_free(x); BN_free(y);
  env_free(&e);
  return 0;
}

int po_fixed_base_mul(const uint8_t *scalars, uint8_t *out, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  EC_POINT *r = EC_POINT_new(e.group);
  for (size_t i = 0; i < n; ++i) {
    BIGNUM *k = sc_in(scalars + 32 * i);
    EC_POINT_mul(e.group, r, k, NULL, NULL, e.ctx); /* SEAL/bidder.cpp:98 */
    pt_out(&e, out + 64 * i, r);
    BN_free(k);
  }
  EC_POINT_free(r);
  env_free(&e);
  return 0;
}

int po_var_base_mul(const uint8_t *points, const uint8_t *scalars, uint8_t *out, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  EC_POINT *r = EC_POINT_new(e.group);
  int rc = 0;
  for (size_t i = 0; i < n; ++i) {
    EC_POINT *P = pt_in(&e, points + 64 * i);
    if (!P) { rc = -2; break; }
    BIGNUM *k = sc_in(scalars + 32 * i);
    EC_POINT_mul(e.group, r, NULL, P, k, e.ctx); /* SEAL/bidder.cpp:129 */
    pt_out(&e, out + 64 * i, r);
    BN_free(k);
    EC_POINT_free(P);
  }
  EC_POINT_free(r);
  env_free(&e);
  return rc;
}

int po_double_mul(const uint8_t *a, const uint8_t *points, const uint8_t *b, uint8_t *out, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  EC_POINT *r = EC_POINT_new(e.group);
  int rc = 0;
  for (size_t i = 0; i < n; ++i) {
    EC_POINT *P = pt_in(&e, points + 64 * i);
    if (!P) { rc = -2; break; }
    BIGNUM *ka = sc_in(a + 32 * i), *kb = sc_in(b + 32 * i);
    EC_POINT_mul(e.group, r, ka, P, kb, e.ctx); /* SEAL/bidder.cpp:175 */
    pt_out(&e, out + 64 * i, r);
    BN_free(ka);
    BN_free(kb);
    EC_POINT_free(P);
  }
  EC_POINT_free(r);
  env_free(&e);
  return rc;
}

int po_lincomb2(const uint8_t *p, const uint8_t *a, const uint8_t *q, const uint8_t *b, uint8_t *out, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  EC_POINT *t1 = EC_POINT_new(e.group), *t2 = EC_POINT_new(e.group);
  int rc = 0;
  for (size_t i = 0; i < n; ++i) {
    EC_POINT *P = pt_in(&e, p + 64 * i), *Q = pt_in(&e, q + 64 * i);
    if (!P || !Q) { rc = -2; break; }
    BIGNUM *ka = sc_in(a + 32 * i), *kb = sc_in(b + 32 * i);
    EC_POINT_mul(e.group, t1, NULL, P, ka, e.ctx); /* SEAL/bidder.cpp:266 */
    EC_POINT_mul(e.group, t2, NULL, Q, kb, e.ctx); /* SEAL/bidder.cpp:267 */
    EC_POINT_add(e.group, t2, t1, t2, e.ctx);      /* SEAL/bidder.cpp:268 */
    pt_out(&e, out + 64 * i, t2);
    BN_free(ka);
    BN_free(kb);
    EC_POINT_free(P);
    EC_POINT_free(Q);
  }
  EC_POINT_free(t1);
  EC_POINT_free(t2);
  env_free(&e);
  return rc;
}

int po_point_add(const uint8_t *p, const uint8_t *q, uint8_t *out, size_t n, int sub) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n; ++i) {
    EC_POINT *P = pt_in(&e, p + 64 * i), *Q = pt_in(&e, q + 64 * i);
    if (!P || !Q) { rc = -2; break; }
    if (sub) EC_POINT_invert(e.group, Q, e.ctx); /* SEAL/bidder.cpp:179 */
    EC_POINT_add(e.group, P, P, Q, e.ctx);       /* SEAL/bidder.cpp:180 */
    pt_out(&e, out + 64 * i, P);
    EC_POINT_free(P);
    EC_POINT_free(Q);
  }
  env_free(&e);
  return rc;
}

int po_point_encode(const uint8_t *points, size_t n, int compressed, uint8_t *out, size_t stride, uint32_t *lens) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n; ++i) {
    EC_POINT *P = pt_in(&e, points + 64 * i);
    if (!P) { rc = -2; break; }
    memset(out + stride * i, 0, stride);
    /* SEAL/hash.cpp:27-29 (uncompressed); compressed is what BASELINE.json's north_star asks to pin too */
    size_t len = EC_POINT_point2oct(e.group, P, compressed ? POINT_CONVERSION_COMPRESSED : POINT_CONVERSION_UNCOMPRESSED,
                                    out + stride * i, stride, e.ctx);
    lens[i] = (uint32_t)len;
    EC_POINT_free(P);
  }
  env_free(&e);
  return rc;
}
