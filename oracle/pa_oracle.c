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

/* ======================= Fiat-Shamir challenge ============================= */

/* h = SHA-256(enc(g) || enc(p_0) .. enc(p_{k-1}) || id) mod order; SEAL/hash.cpp:25-51.
 * enc = EC_POINT_point2oct uncompressed (1 byte for infinity); id = raw size_t. */
static void po_hash(po_env *e, BIGNUM *h, EC_POINT *const *pts, size_t k, uint64_t id) {
  EVP_MD_CTX *md = EVP_MD_CTX_new();
  unsigned char buf[65], dg[32];
  EVP_DigestInit_ex(md, EVP_sha256(), NULL);
  for (size_t i = 0; i <= k; ++i) {
    const EC_POINT *P = i == 0 ? e->g : pts[i - 1];
    size_t len = EC_POINT_point2oct(e->group, P, POINT_CONVERSION_UNCOMPRESSED, buf, sizeof buf, e->ctx);
    EVP_DigestUpdate(md, buf, len);
  }
  size_t id_ = (size_t)id;
  EVP_DigestUpdate(md, &id_, sizeof id_); /* SEAL/hash.cpp:40 */
  EVP_DigestFinal_ex(md, dg, NULL);
  BN_bin2bn(dg, 32, h);
  BN_mod(h, h, e->order, e->ctx); /* SEAL/hash.cpp:50-51 */
  EVP_MD_CTX_free(md);
}

int po_challenge(const uint8_t *points, size_t k, const uint64_t *ids, uint8_t *out, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  BIGNUM *h = BN_new();
  EC_POINT **pts = calloc(k ? k : 1, sizeof *pts);
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    for (size_t j = 0; j < k; ++j)
      if (!(pts[j] = pt_in(&e, points + 64 * (i * k + j)))) rc = -2;
    if (!rc) {
      po_hash(&e, h, pts, k, ids[i]);
      BN_bn2binpad(h, out + 32 * i, 32);
    }
    for (size_t j = 0; j < k; ++j) EC_POINT_free(pts[j]);
  }
  free(pts);
  BN_free(h);
  env_free(&e);
  return rc;
}

/* ============================ proof helpers ================================ */

/* r = a*P (+ b*Q): the reference always does this as separate EC_POINT_mul
 * calls followed by EC_POINT_add (e.g. SEAL/bidder.cpp:364-366); P == NULL is the generator. */
static EC_POINT *po_lin(po_env *e, const EC_POINT *P, const BIGNUM *a, const EC_POINT *Q, const BIGNUM *b) {
  EC_POINT *r = EC_POINT_new(e->group);
  if (P)
    EC_POINT_mul(e->group, r, NULL, P, a, e->ctx);
  else
    EC_POINT_mul(e->group, r, a, NULL, NULL, e->ctx);
  if (Q) {
    EC_POINT *t = EC_POINT_new(e->group);
    EC_POINT_mul(e->group, t, NULL, Q, b, e->ctx);
    EC_POINT_add(e->group, r, r, t, e->ctx);
    EC_POINT_free(t);
  }
  return r;
}
/* c/g: tmp = g; invert; tmp = c + tmp.  SEAL/bidder.cpp:178-180 */
static EC_POINT *po_over_g(po_env *e, const EC_POINT *c) {
  EC_POINT *t = EC_POINT_dup(e->g, e->group);
  EC_POINT_invert(e->group, t, e->ctx);
  EC_POINT_add(e->group, t, c, t, e->ctx);
  return t;
}
/* a*P + b*Q == E ?  EC_POINT_cmp, SEAL/bidder.cpp:131 */
static int po_check(po_env *e, const EC_POINT *P, const BIGNUM *a, const EC_POINT *Q, const BIGNUM *b, const EC_POINT *E) {
  EC_POINT *r = po_lin(e, P, a, Q, b);
  int ok = EC_POINT_cmp(e->group, r, E, e->ctx) == 0;
  EC_POINT_free(r);
  return ok;
}
/* r = x - y*z mod order  (BN_mod_mul then BN_mod_sub, SEAL/bidder.cpp:102-103) */
static BIGNUM *po_resp(po_env *e, const BIGNUM *x, const BIGNUM *y, const BIGNUM *z) {
  BIGNUM *t = BN_new(), *r = BN_new();
  BN_mod_mul(t, y, z, e->order, e->ctx);
  BN_mod_sub(r, x, t, e->order, e->ctx);
  BN_free(t);
  return r;
}
static BIGNUM *po_sub(po_env *e, const BIGNUM *x, const BIGNUM *y) {
  BIGNUM *r = BN_new();
  BN_mod_sub(r, x, y, e->order, e->ctx);
  return r;
}
static void sc_out(uint8_t *b, const BIGNUM *k) { BN_bn2binpad(k, b, 32); }

#define PTS_IN(arr, cnt, src)                                        \
  for (size_t j_ = 0; j_ < (cnt); ++j_)                              \
    if (!((arr)[j_] = pt_in(&e, (src) + 64 * j_))) { rc = -2; }
#define PTS_FREE(arr, cnt) \
  for (size_t j_ = 0; j_ < (cnt); ++j_) EC_POINT_free((arr)[j_]);

/* ============================ NIZKPoKDLog ================================== */
/* SEAL/bidder.cpp:90-107.  rnd: v */
int po_pokdlog_prove(const uint8_t *X, const uint8_t *x, const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    EC_POINT *gx = pt_in(&e, X + 64 * i);
    if (!gx) { rc = -2; break; }
    BIGNUM *v = sc_in(rnd + 32 * i), *sx = sc_in(x + 32 * i), *h = BN_new();
    EC_POINT *gv = po_lin(&e, NULL, v, NULL, NULL); /* :98 */
    EC_POINT *hp[2] = {gv, gx};
    po_hash(&e, h, hp, 2, ids[i]);                  /* :100 */
    BIGNUM *rho = po_resp(&e, v, h, sx);            /* :102-103 */
    pt_out(&e, proofs + 96 * i, gv);
    sc_out(proofs + 96 * i + 64, rho);
    BN_free(v); BN_free(sx); BN_free(h); BN_free(rho);
    EC_POINT_free(gv); EC_POINT_free(gx);
  }
  env_free(&e);
  return rc;
}
/* SEAL/bidder.cpp:119-136 */
int po_pokdlog_verify(const uint8_t *proofs, const uint8_t *X, const uint64_t *ids, uint8_t *verdict, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    EC_POINT *eps = pt_in(&e, proofs + 96 * i), *gx = pt_in(&e, X + 64 * i);
    if (!eps || !gx) { rc = -2; break; }
    BIGNUM *rho = sc_in(proofs + 96 * i + 64), *h = BN_new();
    EC_POINT *hp[2] = {eps, gx};
    po_hash(&e, h, hp, 2, ids[i]);                      /* :125 */
    verdict[i] = (uint8_t)po_check(&e, NULL, rho, gx, h, eps); /* :128-131 */
    BN_free(rho); BN_free(h);
    EC_POINT_free(eps); EC_POINT_free(gx);
  }
  env_free(&e);
  return rc;
}

/* ============================ NIZKPoWFCom ================================== */
/* SEAL/bidder.cpp:149-226.  stmt (phi, A, B); rnd: r1, then bit0: ch2, rho2 / bit1: ch1, rho1 */
int po_powfcom_prove(const uint8_t *stmt, const uint8_t *alpha, const uint8_t *bits, const uint64_t *ids,
                     const uint8_t *rnd, uint8_t *proofs, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    EC_POINT *S[3];
    PTS_IN(S, 3, stmt + 192 * i);
    if (rc) break;
    EC_POINT *phi = S[0], *A = S[1], *B = S[2], *pg = po_over_g(&e, phi);
    BIGNUM *al = sc_in(alpha + 32 * i), *r1 = sc_in(rnd + 96 * i), *chs = sc_in(rnd + 96 * i + 32),
           *rhos = sc_in(rnd + 96 * i + 64), *ch = BN_new();
    EC_POINT *E[4];
    if (bits[i] == 0) {
      E[0] = po_lin(&e, NULL, r1, NULL, NULL);  /* eps11 = g^r1            :171 */
      E[1] = po_lin(&e, B, r1, NULL, NULL);     /* eps12 = B^r1            :173 */
      E[2] = po_lin(&e, NULL, rhos, A, chs);    /* eps21 = g^rho2 A^ch2    :175 */
      E[3] = po_lin(&e, B, rhos, pg, chs);      /* eps22 = B^rho2 (phi/g)^ch2 :178-185 */
    } else {
      E[0] = po_lin(&e, NULL, rhos, A, chs);    /* eps11 = g^rho1 A^ch1    :190-193 */
      E[1] = po_lin(&e, phi, chs, B, rhos);     /* eps12 = phi^ch1 B^rho1  :195-198 */
      E[2] = po_lin(&e, NULL, r1, NULL, NULL);  /* eps21 = g^r1            :200 */
      E[3] = po_lin(&e, B, r1, NULL, NULL);     /* eps22 = B^r1            :202 */
    }
    EC_POINT *hp[7] = {E[0], E[1], E[2], E[3], phi, A, B};
    po_hash(&e, ch, hp, 7, ids[i]);             /* :205 */
    BIGNUM *chr = po_sub(&e, ch, chs);          /* real challenge          :209 / :213 */
    BIGNUM *rhor = po_resp(&e, r1, al, chr);    /* r1 - alpha*ch_real      :210-211 / :214-215 */
    uint8_t *o = proofs + 352 * i;
    for (int k = 0; k < 4; ++k) pt_out(&e, o + 64 * k, E[k]);
    if (bits[i] == 0) {
      sc_out(o + 256, rhor); sc_out(o + 288, rhos); sc_out(o + 320, chs);
    } else {
      sc_out(o + 256, rhos); sc_out(o + 288, rhor); sc_out(o + 320, chr); /* ch2 = ch - ch1 :216 */
    }
    BN_free(al); BN_free(r1); BN_free(chs); BN_free(rhos); BN_free(ch); BN_free(chr); BN_free(rhor);
    PTS_FREE(E, 4); PTS_FREE(S, 3); EC_POINT_free(pg);
  }
  env_free(&e);
  return rc;
}
/* SEAL/bidder.cpp:241-299 — all four checks run, no early exit */
int po_powfcom_verify(const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    const uint8_t *p = proofs + 352 * i;
    EC_POINT *S[3], *E[4];
    PTS_IN(S, 3, stmt + 192 * i);
    PTS_IN(E, 4, p);
    if (rc) break;
    EC_POINT *phi = S[0], *A = S[1], *B = S[2], *pg = po_over_g(&e, phi);
    BIGNUM *rho1 = sc_in(p + 256), *rho2 = sc_in(p + 288), *ch2 = sc_in(p + 320), *ch = BN_new();
    EC_POINT *hp[7] = {E[0], E[1], E[2], E[3], phi, A, B};
    po_hash(&e, ch, hp, 7, ids[i]);   /* :250 */
    BIGNUM *ch1 = po_sub(&e, ch, ch2); /* :253 */
    int ok = 1;
    ok &= po_check(&e, NULL, rho1, A, ch1, E[0]); /* check 1 :256-259 */
    ok &= po_check(&e, B, rho1, phi, ch1, E[1]);  /* check 2 :266-269 */
    ok &= po_check(&e, NULL, rho2, A, ch2, E[2]); /* check 3 :276-279 */
    ok &= po_check(&e, B, rho2, pg, ch2, E[3]);   /* check 4 :286-292 */
    verdict[i] = (uint8_t)ok;
    BN_free(rho1); BN_free(rho2); BN_free(ch2); BN_free(ch); BN_free(ch1);
    PTS_FREE(E, 4); PTS_FREE(S, 3); EC_POINT_free(pg);
  }
  env_free(&e);
  return rc;
}

/* ============================ NIZKPoWFStage1 =============================== */
/* SEAL/bidder.cpp:318-451.  stmt (b, X, Y, R, c, A, B); secrets (x, alpha);
 * rnd: r11, r12, then bit0: rho21, rho22, ch2 / bit1: rho11, rho12, ch1 */
int po_stage1_prove(const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bits, const uint64_t *ids,
                    const uint8_t *rnd, uint8_t *proofs, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    EC_POINT *S[7];
    PTS_IN(S, 7, stmt + 448 * i);
    if (rc) break;
    EC_POINT *b = S[0], *X = S[1], *Y = S[2], *R = S[3], *c = S[4], *A = S[5], *B = S[6], *cg = po_over_g(&e, c);
    const uint8_t *rd = rnd + 160 * i;
    BIGNUM *x = sc_in(secrets + 64 * i), *al = sc_in(secrets + 64 * i + 32);
    BIGNUM *r11 = sc_in(rd), *r12 = sc_in(rd + 32), *d1 = sc_in(rd + 64), *d2 = sc_in(rd + 96), *dch = sc_in(rd + 128), *ch = BN_new();
    EC_POINT *E[8];
    int real = bits[i] == 0 ? 0 : 4, sim = 4 - real; /* which half of the eps array is the real branch */
    /* real branch: g^r11, g^r12, (Y | R)^r11, B^r12                 :355-361 / :410-417 */
    E[real + 0] = po_lin(&e, NULL, r11, NULL, NULL);
    E[real + 1] = po_lin(&e, NULL, r12, NULL, NULL);
    E[real + 2] = po_lin(&e, bits[i] == 0 ? Y : R, r11, NULL, NULL);
    E[real + 3] = po_lin(&e, B, r12, NULL, NULL);
    /* simulated branch with drawn (rho_a, rho_b, ch)               :364-384 / :391-408 */
    E[sim + 0] = po_lin(&e, NULL, d1, X, dch);
    E[sim + 1] = po_lin(&e, NULL, d2, A, dch);
    E[sim + 2] = po_lin(&e, bits[i] == 0 ? R : Y, d1, b, dch);
    E[sim + 3] = po_lin(&e, B, d2, bits[i] == 0 ? cg : c, dch);
    EC_POINT *hp[15] = {E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], b, X, Y, R, c, A, B};
    po_hash(&e, ch, hp, 15, ids[i]);             /* :420-422 */
    BIGNUM *chr = po_sub(&e, ch, dch);           /* :425 / :431 */
    BIGNUM *rx = po_resp(&e, r11, chr, x);       /* :426-427 / :432-433 */
    BIGNUM *ra = po_resp(&e, r12, chr, al);      /* :428-429 / :434-435 */
    uint8_t *o = proofs + 672 * i;
    for (int k = 0; k < 8; ++k) pt_out(&e, o + 64 * k, E[k]);
    o += 512; /* rho11 rho12 rho21 rho22 ch2 */
    if (bits[i] == 0) {
      sc_out(o, rx); sc_out(o + 32, ra); sc_out(o + 64, d1); sc_out(o + 96, d2); sc_out(o + 128, dch);
    } else {
      sc_out(o, d1); sc_out(o + 32, d2); sc_out(o + 64, rx); sc_out(o + 96, ra); sc_out(o + 128, chr);
    }
    BN_free(x); BN_free(al); BN_free(r11); BN_free(r12); BN_free(d1); BN_free(d2); BN_free(dch); BN_free(ch);
    BN_free(chr); BN_free(rx); BN_free(ra);
    PTS_FREE(E, 8); PTS_FREE(S, 7); EC_POINT_free(cg);
  }
  env_free(&e);
  return rc;
}
/* SEAL/bidder.cpp:470-571 */
int po_stage1_verify(const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    const uint8_t *p = proofs + 672 * i;
    EC_POINT *S[7], *E[8];
    PTS_IN(S, 7, stmt + 448 * i);
    PTS_IN(E, 8, p);
    if (rc) break;
    EC_POINT *b = S[0], *X = S[1], *Y = S[2], *R = S[3], *c = S[4], *A = S[5], *B = S[6], *cg = po_over_g(&e, c);
    BIGNUM *rho11 = sc_in(p + 512), *rho12 = sc_in(p + 544), *rho21 = sc_in(p + 576), *rho22 = sc_in(p + 608),
           *ch2 = sc_in(p + 640), *ch = BN_new();
    EC_POINT *hp[15] = {E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], b, X, Y, R, c, A, B};
    po_hash(&e, ch, hp, 15, ids[i]);    /* :481-484 */
    BIGNUM *ch1 = po_sub(&e, ch, ch2);  /* :485 */
    int ok = 1;
    ok &= po_check(&e, NULL, rho11, X, ch1, E[0]); /* check 1 :488-491 */
    ok &= po_check(&e, NULL, rho12, A, ch1, E[1]); /* check 2 :498-501 */
    ok &= po_check(&e, Y, rho11, b, ch1, E[2]);    /* check 3 :508-511 */
    ok &= po_check(&e, B, rho12, c, ch1, E[3]);    /* check 4 :518-521 */
    ok &= po_check(&e, NULL, rho21, X, ch2, E[4]); /* check 5 :528-531 */
    ok &= po_check(&e, NULL, rho22, A, ch2, E[5]); /* check 6 :538-541 */
    ok &= po_check(&e, R, rho21, b, ch2, E[6]);    /* check 7 :548-551 */
    ok &= po_check(&e, B, rho22, cg, ch2, E[7]);   /* check 8 :558-564 */
    verdict[i] = (uint8_t)ok;
    BN_free(rho11); BN_free(rho12); BN_free(rho21); BN_free(rho22); BN_free(ch2); BN_free(ch); BN_free(ch1);
    PTS_FREE(E, 8); PTS_FREE(S, 7); EC_POINT_free(cg);
  }
  env_free(&e);
  return rc;
}

/* ============================ NIZKPoWFStage2 =============================== */
/* eps order in the record and in the hash (SEAL/types.h:64-80, SEAL/hash.cpp:191-194):
 *  0 eps11 1 eps12 2 eps13 3 eps11' 4 eps12' 5 eps13' 6 eps21 7 eps22 8 eps23 9 eps21'
 * 10 eps22' 11 eps23' 12 eps31 13 eps32 14 eps31' 15 eps32' */
/* SEAL/bidder.cpp:598-890.  stmt (Bi, Xi, Ri, Bj, Xj, Rj, Ci, A, B, Yi, Yj); secrets (xi, xj, alpha);
 * rnd: 11 draws in the order of :643-655 / :692-699 / :749-756 */
int po_stage2_prove(const uint8_t *stmt, const uint8_t *secrets, const uint8_t *bi, const uint8_t *bj,
                    const uint64_t *ids, const uint8_t *rnd, uint8_t *proofs, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    if (bi[i] == 1 && bj[i] == 0) { rc = -3; break; } /* assert at :604-605 */
    EC_POINT *S[11];
    PTS_IN(S, 11, stmt + 704 * i);
    if (rc) break;
    EC_POINT *Bi = S[0], *Xi = S[1], *Ri = S[2], *Bj = S[3], *Xj = S[4], *Rj = S[5], *Ci = S[6], *A = S[7], *B = S[8],
             *Yi = S[9], *Yj = S[10], *cg = po_over_g(&e, Ci);
    BIGNUM *xi = sc_in(secrets + 96 * i), *xj = sc_in(secrets + 96 * i + 32), *al = sc_in(secrets + 96 * i + 64);
    BIGNUM *d[11], *ch = BN_new(), *zero = BN_new();
    BN_zero(zero);
    for (int k = 0; k < 11; ++k) d[k] = sc_in(rnd + 352 * i + 32 * k);
    BIGNUM *r11 = d[0], *r12 = d[1], *r13 = d[2];
    EC_POINT *E[16];
    const BIGNUM *out[10]; /* rho11 rho12 rho13 rho21 rho22 rho23 rho31 rho32 ch2 ch3 */
    BIGNUM *chr = NULL, *q1 = NULL, *q2 = NULL, *q3 = NULL;
    if (bi[i] == 1) {
      /* draws: rho21 rho22 rho23 rho31 rho32 rho33 ch2 ch3 = d[3..10]; rho33 is never used */
      BIGNUM *rho21 = d[3], *rho22 = d[4], *rho23 = d[5], *rho31 = d[6], *rho32 = d[7], *ch2 = d[9], *ch3 = d[10];
      E[0] = po_lin(&e, NULL, r11, NULL, NULL); E[1] = po_lin(&e, NULL, r12, NULL, NULL); E[2] = po_lin(&e, NULL, r13, NULL, NULL); /* :657-659 */
      E[3] = po_lin(&e, Ri, r11, NULL, NULL); E[4] = po_lin(&e, Rj, r12, NULL, NULL); E[5] = po_lin(&e, B, r13, NULL, NULL);      /* :660-662 */
      E[6] = po_lin(&e, NULL, rho21, Xi, ch2); E[7] = po_lin(&e, NULL, rho22, Xj, ch2); E[8] = po_lin(&e, NULL, rho23, A, ch2);    /* :664-666 */
      E[9] = po_lin(&e, Yi, rho21, Bi, ch2); E[10] = po_lin(&e, Rj, rho22, Bj, ch2); E[11] = po_lin(&e, B, rho23, Ci, ch2);        /* :668-678 */
      E[12] = po_lin(&e, NULL, rho31, Xi, ch3); E[13] = po_lin(&e, NULL, rho32, Xj, ch3);                                          /* :680-681 */
      E[14] = po_lin(&e, Yi, rho31, Bi, ch3); E[15] = po_lin(&e, Yj, rho32, Bj, ch3);                                              /* :683-689 */
    } else if (bj[i] == 1) {
      /* draws: rho11 rho12 rho13 rho31 rho32 rho33 ch1 ch3 */
      BIGNUM *rho11 = d[3], *rho12 = d[4], *rho13 = d[5], *rho31 = d[6], *rho32 = d[7], *ch1 = d[9], *ch3 = d[10];
      E[6] = po_lin(&e, NULL, r11, NULL, NULL); E[7] = po_lin(&e, NULL, r12, NULL, NULL); E[8] = po_lin(&e, NULL, r13, NULL, NULL); /* :701-703 */
      E[9] = po_lin(&e, Yi, r11, NULL, NULL); E[10] = po_lin(&e, Rj, r12, NULL, NULL); E[11] = po_lin(&e, B, r13, NULL, NULL);     /* :704-706 */
      E[0] = po_lin(&e, NULL, rho11, Xi, ch1); E[1] = po_lin(&e, NULL, rho12, Xj, ch1); E[2] = po_lin(&e, NULL, rho13, A, ch1);    /* :709-719 */
      E[3] = po_lin(&e, Ri, rho11, Bi, ch1); E[4] = po_lin(&e, Rj, rho12, Bj, ch1); E[5] = po_lin(&e, B, rho13, cg, ch1);          /* :721-734 */
      E[12] = po_lin(&e, NULL, rho31, Xi, ch3); E[13] = po_lin(&e, NULL, rho32, Xj, ch3);                                          /* :736-739 */
      E[14] = po_lin(&e, Yi, rho31, Bi, ch3); E[15] = po_lin(&e, Yj, rho32, Bj, ch3);                                              /* :741-747 */
    } else {
      /* draws: rho21 rho22 rho23 (discarded) rho21 rho22 rho23 ch1 ch2; rho11 = rho12 = rho13 = 0 */
      BIGNUM *rho21 = d[6], *rho22 = d[7], *rho23 = d[8], *ch1 = d[9], *ch2 = d[10];
      E[0] = po_lin(&e, NULL, zero, Xi, ch1); E[1] = po_lin(&e, NULL, zero, Xj, ch1); E[2] = po_lin(&e, NULL, zero, A, ch1);       /* :759-769 */
      E[3] = po_lin(&e, Ri, zero, Bi, ch1); E[4] = po_lin(&e, Rj, zero, Bj, ch1); E[5] = po_lin(&e, B, zero, cg, ch1);             /* :771-784 */
      E[6] = po_lin(&e, NULL, rho21, Xi, ch2); E[7] = po_lin(&e, NULL, rho22, Xj, ch2); E[8] = po_lin(&e, NULL, rho23, A, ch2);    /* :787-797 */
      E[9] = po_lin(&e, Yi, rho21, Bi, ch2); E[10] = po_lin(&e, Rj, rho22, Bj, ch2); E[11] = po_lin(&e, B, rho23, Ci, ch2);        /* :799-809 */
      E[12] = po_lin(&e, NULL, r11, NULL, NULL); E[13] = po_lin(&e, NULL, r12, NULL, NULL);                                        /* :811-812 */
      E[14] = po_lin(&e, Yi, r11, NULL, NULL); E[15] = po_lin(&e, Yj, r12, NULL, NULL);                                            /* :813-814 */
    }
    EC_POINT *hp[27] = {E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], E[8], E[9], E[10], E[11], E[12], E[13], E[14], E[15],
                        Xi, Xj, A, Bi, Bj, B, Ri, Rj, Ci, Yi, Yj};
    po_hash(&e, ch, hp, 27, ids[i]); /* :818-822 */
    BIGNUM *t = po_sub(&e, ch, d[9]);
    chr = po_sub(&e, t, d[10]);      /* real challenge = ch - (the two drawn)  :826-827 / :840-841 / :853-854 */
    BN_free(t);
    q1 = po_resp(&e, r11, xi, chr);  /* r11 - xi*ch   :829-830 / :843-844 / :856-857 */
    q2 = po_resp(&e, r12, xj, chr);  /* r12 - xj*ch   :832-833 / :846-847 / :859-860 */
    q3 = po_resp(&e, r13, al, chr);  /* r13 - alpha*ch :835-836 / :849-850 (unused in branch 3) */
    if (bi[i] == 1) {
      const BIGNUM *o_[10] = {q1, q2, q3, d[3], d[4], d[5], d[6], d[7], d[9], d[10]};
      memcpy(out, o_, sizeof o_);
    } else if (bj[i] == 1) {
      const BIGNUM *o_[10] = {d[3], d[4], d[5], q1, q2, q3, d[6], d[7], chr, d[10]};
      memcpy(out, o_, sizeof o_);
    } else {
      const BIGNUM *o_[10] = {zero, zero, zero, d[6], d[7], d[8], q1, q2, d[10], chr};
      memcpy(out, o_, sizeof o_);
    }
    uint8_t *o = proofs + 1344 * i;
    for (int k = 0; k < 16; ++k) pt_out(&e, o + 64 * k, E[k]);
    for (int k = 0; k < 10; ++k) sc_out(o + 1024 + 32 * k, out[k]);
    for (int k = 0; k < 11; ++k) BN_free(d[k]);
    BN_free(xi); BN_free(xj); BN_free(al); BN_free(ch); BN_free(zero); BN_free(chr); BN_free(q1); BN_free(q2); BN_free(q3);
    PTS_FREE(E, 16); PTS_FREE(S, 11); EC_POINT_free(cg);
  }
  env_free(&e);
  return rc;
}
/* SEAL/bidder.cpp:913-1101 */
int po_stage2_verify(const uint8_t *proofs, const uint8_t *stmt, const uint64_t *ids, uint8_t *verdict, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  for (size_t i = 0; i < n && !rc; ++i) {
    const uint8_t *p = proofs + 1344 * i;
    EC_POINT *S[11], *E[16];
    PTS_IN(S, 11, stmt + 704 * i);
    PTS_IN(E, 16, p);
    if (rc) break;
    EC_POINT *Bi = S[0], *Xi = S[1], *Ri = S[2], *Bj = S[3], *Xj = S[4], *Rj = S[5], *Ci = S[6], *A = S[7], *B = S[8],
             *Yi = S[9], *Yj = S[10], *cg = po_over_g(&e, Ci);
    BIGNUM *s[10], *ch = BN_new();
    for (int k = 0; k < 10; ++k) s[k] = sc_in(p + 1024 + 32 * k);
    BIGNUM *rho11 = s[0], *rho12 = s[1], *rho13 = s[2], *rho21 = s[3], *rho22 = s[4], *rho23 = s[5], *rho31 = s[6],
           *rho32 = s[7], *ch2 = s[8], *ch3 = s[9];
    EC_POINT *hp[27] = {E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], E[8], E[9], E[10], E[11], E[12], E[13], E[14], E[15],
                        Xi, Xj, A, Bi, Bj, B, Ri, Rj, Ci, Yi, Yj};
    po_hash(&e, ch, hp, 27, ids[i]);   /* :928-933 */
    BIGNUM *t = po_sub(&e, ch, ch2), *ch1 = po_sub(&e, t, ch3); /* :934-935 */
    int ok = 1;
    ok &= po_check(&e, NULL, rho11, Xi, ch1, E[0]);  /* check 1  :938-941 */
    ok &= po_check(&e, NULL, rho12, Xj, ch1, E[1]);  /* check 2  :948-951 */
    ok &= po_check(&e, NULL, rho13, A, ch1, E[2]);   /* check 3  :958-961 */
    ok &= po_check(&e, Ri, rho11, Bi, ch1, E[3]);    /* check 4  :968-971 */
    ok &= po_check(&e, Rj, rho12, Bj, ch1, E[4]);    /* check 5  :978-981 */
    ok &= po_check(&e, B, rho13, cg, ch1, E[5]);     /* check 6  :988-994 */
    ok &= po_check(&e, NULL, rho21, Xi, ch2, E[6]);  /* check 7  :1001-1004 */
    ok &= po_check(&e, NULL, rho22, Xj, ch2, E[7]);  /* check 8  :1011-1014 */
    ok &= po_check(&e, NULL, rho23, A, ch2, E[8]);   /* check 9  :1021-1024 */
    ok &= po_check(&e, Yi, rho21, Bi, ch2, E[9]);    /* check 10 :1031-1034 */
    ok &= po_check(&e, Rj, rho22, Bj, ch2, E[10]);   /* check 11 :1041-1044 */
    ok &= po_check(&e, B, rho23, Ci, ch2, E[11]);    /* check 12 :1051-1054 */
    ok &= po_check(&e, NULL, rho31, Xi, ch3, E[12]); /* check 13 :1061-1064 */
    ok &= po_check(&e, NULL, rho32, Xj, ch3, E[13]); /* check 14 :1071-1074 */
    ok &= po_check(&e, Yi, rho31, Bi, ch3, E[14]);   /* check 15 :1081-1084 */
    ok &= po_check(&e, Yj, rho32, Bj, ch3, E[15]);   /* check 16 :1091-1094 */
    verdict[i] = (uint8_t)ok;
    for (int k = 0; k < 10; ++k) BN_free(s[k]);
    BN_free(ch); BN_free(t); BN_free(ch1);
    PTS_FREE(E, 16); PTS_FREE(S, 11); EC_POINT_free(cg);
  }
  env_free(&e);
  return rc;
}

/* ============================== round logic =============================== */

int po_commit_points(const uint8_t *alpha, const uint8_t *beta, const uint8_t *bits, uint8_t *out, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  EC_POINT *r = EC_POINT_new(e.group);
  for (size_t i = 0; i < n; ++i) {
    BIGNUM *a = sc_in(alpha + 32 * i), *b = sc_in(beta + 32 * i), *ab = BN_new(), *bit = BN_new();
    BN_set_word(bit, bits[i]);                             /* :1129 */
    BN_mul(ab, a, b, e.ctx);                               /* unreduced product, :1133 */
    EC_POINT_mul(e.group, r, ab, e.g, bit, e.ctx);         /* phi = g^(alpha beta) g^bit  :1135 */
    pt_out(&e, out + 192 * i, r);
    EC_POINT_mul(e.group, r, a, NULL, NULL, e.ctx);        /* A = g^alpha  :1137 */
    pt_out(&e, out + 192 * i + 64, r);
    EC_POINT_mul(e.group, r, b, NULL, NULL, e.ctx);        /* B = g^beta   :1138 */
    pt_out(&e, out + 192 * i + 128, r);
    BN_free(a); BN_free(b); BN_free(ab); BN_free(bit);
  }
  EC_POINT_free(r);
  env_free(&e);
  return 0;
}

int po_y_scan(const uint8_t *X, uint8_t *Y, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  EC_POINT **xs = calloc(n ? n : 1, sizeof *xs);
  PTS_IN(xs, n, X);
  if (!rc) {
    EC_POINT *first = EC_POINT_new(e.group), *second = EC_POINT_new(e.group), *y = EC_POINT_new(e.group);
    for (size_t id = 0; id < n; ++id) {
      EC_POINT_set_to_infinity(e.group, first);                                              /* :1287 */
      for (size_t i = 0; i < id; ++i) EC_POINT_add(e.group, first, first, xs[i], e.ctx);     /* :1288-1290 */
      EC_POINT_set_to_infinity(e.group, second);                                             /* :1292 */
      for (size_t i = id + 1; i < n; ++i) EC_POINT_add(e.group, second, second, xs[i], e.ctx); /* :1293-1295 */
      EC_POINT_invert(e.group, second, e.ctx);                                               /* :1297 */
      EC_POINT_add(e.group, y, first, second, e.ctx);                                        /* :1298 */
      pt_out(&e, Y + 64 * id, y);
    }
    EC_POINT_free(first); EC_POINT_free(second); EC_POINT_free(y);
  }
  for (size_t j = 0; j < n; ++j) EC_POINT_free(xs[j]);
  free(xs);
  env_free(&e);
  return rc;
}

int po_point_sum_is_inf(const uint8_t *b, size_t n, int *is_inf) {
  po_env e;
  if (env_init(&e)) return -1;
  int rc = 0;
  EC_POINT *sum = EC_POINT_new(e.group); /* a fresh EC_POINT is the point at infinity, :1390 */
  for (size_t i = 0; i < n && !rc; ++i) {
    EC_POINT *P = pt_in(&e, b + 64 * i);
    if (!P) { rc = -2; break; }
    EC_POINT_add(e.group, sum, sum, P, e.ctx); /* :1394 */
    EC_POINT_free(P);
  }
  *is_inf = EC_POINT_is_at_infinity(e.group, sum); /* :1397 */
  EC_POINT_free(sum);
  env_free(&e);
  return rc;
}

/* ================================ CCS22 ==================================== */

int po_ccs22_setup_hash(const uint8_t *scalars, size_t k, uint8_t *out, size_t n) {
  po_env e;
  if (env_init(&e)) return -1;
  for (size_t i = 0; i < n; ++i) {
    EVP_MD_CTX *md = EVP_MD_CTX_new();
    unsigned char buf[32], dg[32];
    BIGNUM *h = BN_new(); /* stays 0 on the error path, CCS22/hash.cpp:31-35 */
    int failed = 0;
    EVP_DigestInit_ex(md, EVP_sha256(), NULL);
    for (size_t j = 0; j < k && !failed; ++j) {
      BIGNUM *b = sc_in(scalars + 32 * (i * k + j));
      int len = BN_bn2bin(b, buf); /* minimal length, CCS22/hash.cpp:26-31 */
      if (len == 0) failed = 1;    /* zero bignum: handelSHA256Error + return */
      else EVP_DigestUpdate(md, buf, (size_t)len);
      BN_free(b);
    }
    if (!failed) {
      EVP_DigestFinal_ex(md, dg, NULL);
      BN_bin2bn(dg, 32, h);
      BN_mod(h, h, e.order, e.ctx); /* CCS22/hash.cpp:53-54 */
    }
    BN_bn2binpad(h, out + 32 * i, 32);
    BN_free(h);
    EVP_MD_CTX_free(md);
  }
  env_free(&e);
  return 0;
}
