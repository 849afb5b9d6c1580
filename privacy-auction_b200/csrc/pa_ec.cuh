// secp256k1 group law (y^2 = x^3 + 7, a = 0) in Jacobian coordinates.
// Replaces libcrypto's EC_POINT_add / EC_POINT_dbl / EC_POINT_invert /
// EC_POINT_cmp / EC_POINT_is_at_infinity as the reference calls them
// (SEAL/bidder.cpp:130-131, 178-185, 1289-1298, 1394-1397).
//
// Affine results are mathematically unique, so they are bit-identical to what
// libcrypto returns no matter which formulas or evaluation order produce them;
// the only places where care is needed are the exceptional cases (infinity,
// P + P, P + (-P)), all of which are handled explicitly here because the
// protocol reaches them (n = 1 gives Y = infinity; stage-2 branch
// 3 multiplies by the scalar 0).
#pragma once
#include "pa_fe.cuh"

struct jac {  // infinity <=> Z == 0
  fe X, Y, Z;
};
struct aff {  // infinity <=> (0, 0), which is not on the curve
  fe x, y;
};

PA_HD void jac_set_inf(jac &r) {
  fe_set_one(r.X);
  fe_set_one(r.Y);
  fe_set_zero(r.Z);
}
PA_HD bool jac_is_inf(const jac &p) { return fe_is_zero(p.Z); }
PA_HD bool aff_is_inf(const aff &p) {
  u32 z = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) z |= p.x.v[i] | p.y.v[i];
  return z == 0;
}
PA_HD void aff_set_inf(aff &r) {
  fe_set_zero(r.x);
  fe_set_zero(r.y);
}
PA_HD void jac_from_aff(jac &r, const aff &p) {
  if (aff_is_inf(p)) {
    jac_set_inf(r);
  } else {
    r.X = p.x;
    r.Y = p.y;
    fe_set_one(r.Z);
  }
}
PA_HD void aff_neg(aff &r, const aff &p) {
  r.x = p.x;
  if (aff_is_inf(p)) {
    r.y = p.y;
  } else {
    fe_neg(r.y, p.y);
    fe_canon(r.y, r.y);
  }
}
PA_HD void jac_neg(jac &r, const jac &p) {
  r.X = p.X;
  fe_neg(r.Y, p.Y);
  r.Z = p.Z;
}

#if defined(__CUDA_ARCH__) && !defined(PA_EC_INLINE)
static __device__ __noinline__ jac jac_dbl_call(jac p);
PA_D void jac_dbl(jac &r, const jac &p) { r = jac_dbl_call(p); }
#else
PA_HD void jac_dbl_inl(jac &r, const jac &p);
PA_HD void jac_dbl(jac &r, const jac &p) { jac_dbl_inl(r, p); }
#endif

// r = 2p   (dbl-2009-l, 2M + 5S).  No point of order 2 exists (prime order).
// Arranged as three stages of two independent products plus one final product.
// Infinity needs no test: Z = 0 gives Z3 = 2 Y Z = 0, and infinity is recognised by Z alone.
PA_HD void jac_dbl_inl(jac &r, const jac &p) {
  fe A, B, C, D, E, F, t, yz;
  fe_sqr2(A, p.X, B, p.Y);  // A = X^2, B = Y^2
  fe_add(t, p.X, B);
  fe_sqr2(C, B, t, t);  // C = B^2, t = (X + B)^2
  fe_sub2(t, t, A, C);
  fe_dbl(D, t);  // D = 2((X+B)^2 - A - C)
  fe_mul3(E, A);  // E = 3A
  fe_sqrmul(F, E, yz, p.Y, p.Z);  // F = E^2, yz = Y*Z
  fe_dbl(r.Z, yz);  // Z3 = 2YZ
  fe_sub2(r.X, F, D, D);  // X3 = F - 2D
  fe_sub(t, D, r.X);
  fe_mul(t, E, t);
  fe_shl<3>(C, C);
  fe_sub(r.Y, t, C);  // Y3 = E(D - X3) - 8C
}

// r = p + q, q affine   (8M + 3S), in six stages of (mostly) two independent products
// Q_NOT_INF: the caller knows q is a real point (window-table entries), so the test is left out.
template <bool Q_NOT_INF>
PA_HD void jac_madd_t(jac &r, const jac &p, const aff &q) {
  if (!Q_NOT_INF && aff_is_inf(q)) {
    r = p;
    return;
  }
  if (jac_is_inf(p)) {
    r.X = q.x;
    r.Y = q.y;
    fe_set_one(r.Z);
    return;
  }
  fe zz, u2, zzz, s2, z3, h, rr, hh, r2, hhh, v, t, yh;
  fe_sqr(zz, p.Z);
  fe_mul2(u2, q.x, zz, zzz, p.Z, zz);  // U2 = x2 Z1^2, Z1^3
  fe_sub(h, u2, p.X);
  fe_mul2(s2, q.y, zzz, z3, p.Z, h);  // S2 = y2 Z1^3, Z3 = Z1 H
  fe_sub(rr, s2, p.Y);
  if (fe_is_zero(h)) {
    if (fe_is_zero(rr)) {
      jac_dbl(r, p);
    } else {
      jac_set_inf(r);
    }
    return;
  }
  fe_sqr2(hh, h, r2, rr);              // H^2, R^2
  fe_mul2(hhh, h, hh, v, p.X, hh);     // H^3, V = X1 H^2
  fe_sub3(t, r2, hhh, v, v);  // X3 = R^2 - H^3 - 2V
  fe_sub(v, v, t);
  fe_mul2(yh, p.Y, hhh, v, rr, v);     // Y1 H^3, R (V - X3)
  r.X = t;
  r.Z = z3;
  fe_sub(r.Y, v, yh);  // Y3 = R(V - X3) - Y1 H^3
}

PA_HD void jac_madd_inl(jac &r, const jac &p, const aff &q) { jac_madd_t<false>(r, p, q); }

// r = p + q   (12M + 4S)
PA_HD void jac_add_inl(jac &r, const jac &p, const jac &q) {
  if (jac_is_inf(q)) {
    r = p;
    return;
  }
  if (jac_is_inf(p)) {
    r = q;
    return;
  }
  fe z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t;
  fe_sqr(z1z1, p.Z);
  fe_sqr(z2z2, q.Z);
  fe_mul(u1, p.X, z2z2);
  fe_mul(u2, q.X, z1z1);
  fe_mul(s1, q.Z, z2z2);
  fe_mul(s1, p.Y, s1);
  fe_mul(s2, p.Z, z1z1);
  fe_mul(s2, q.Y, s2);
  fe_sub(h, u2, u1);
  fe_sub(rr, s2, s1);
  if (fe_is_zero(h)) {
    if (fe_is_zero(rr)) {
      jac_dbl(r, p);
    } else {
      jac_set_inf(r);
    }
    return;
  }
  fe_sqr(hh, h);
  fe_mul(hhh, h, hh);
  fe_mul(v, u1, hh);
  fe_mul(t, p.Z, q.Z);
  fe_mul(r.Z, t, h);
  fe_sqr(t, rr);
  fe_sub3(t, t, hhh, v, v);
  fe_mul(hhh, s1, hhh);
  r.X = t;
  fe_sub(v, v, t);
  fe_mul(v, rr, v);
  fe_sub(r.Y, v, hhh);
}

// On the device the three group operations are real functions (arguments and
// result in registers) so that a scalar multiplication is a few KB of code.
#if defined(__CUDA_ARCH__) && !defined(PA_EC_INLINE)
static __device__ __noinline__ jac jac_dbl_call(jac p) {
  jac r;
  jac_dbl_inl(r, p);
  return r;
}
static __device__ __noinline__ jac jac_madd_call(jac p, aff q) {
  jac r;
  jac_madd_inl(r, p, q);
  return r;
}
static __device__ __noinline__ jac jac_add_call(jac p, jac q) {
  jac r;
  jac_add_inl(r, p, q);
  return r;
}
PA_D void jac_madd(jac &r, const jac &p, const aff &q) { r = jac_madd_call(p, q); }
PA_D void jac_add(jac &r, const jac &p, const jac &q) { r = jac_add_call(p, q); }
#else
PA_HD void jac_madd(jac &r, const jac &p, const aff &q) { jac_madd_inl(r, p, q); }
PA_HD void jac_add(jac &r, const jac &p, const jac &q) { jac_add_inl(r, p, q); }
#endif

// Jacobian == affine without an inversion (EC_POINT_cmp, SEAL/bidder.cpp:131)
PA_HD bool jac_eq_aff(const jac &p, const aff &q) {
  bool pi = jac_is_inf(p), qi = aff_is_inf(q);
  if (pi || qi) return pi && qi;
  fe zz, t;
  fe_sqr(zz, p.Z);
  fe_mul(t, q.x, zz);
  if (!fe_eq(t, p.X)) return false;
  fe_mul(zz, zz, p.Z);
  fe_mul(t, q.y, zz);
  return fe_eq(t, p.Y);
}

// is the affine point on the curve (or infinity)?
PA_HD bool aff_on_curve(const aff &p) {
  if (aff_is_inf(p)) return true;
  fe l, r;
  fe_sqr(l, p.y);
  fe_sqr(r, p.x);
  fe_mul(r, r, p.x);
  fe seven;
  fe_set_zero(seven);
  seven.v[0] = 7;
  fe_add(r, r, seven);
  return fe_eq(l, r);
}

// given zi = 1/Z: affine coordinates, canonical
PA_HD void jac_to_aff_with_zinv(aff &r, const jac &p, const fe &zi) {
  fe zi2, zi3;
  fe_sqr(zi2, zi);
  fe_mul(zi3, zi2, zi);
  fe_mul(r.x, p.X, zi2);
  fe_mul(r.y, p.Y, zi3);
  fe_canon(r.x, r.x);
  fe_canon(r.y, r.y);
}
PA_HD void jac_to_aff(aff &r, const jac &p) {
  if (jac_is_inf(p)) {
    aff_set_inf(r);
    return;
  }
  fe zi;
  fe_inv(zi, p.Z);
  jac_to_aff_with_zinv(r, p, zi);
}

// 64-byte wire form: X || Y big-endian, infinity = 64 zero bytes
PA_HD void aff_from_be64(aff &r, const unsigned char *b) {
  fe_from_be(r.x, b);
  fe_from_be(r.y, b + 32);
}
PA_HD void aff_to_be64(unsigned char *b, const aff &p) {
  fe_to_be(b, p.x);
  fe_to_be(b + 32, p.y);
}
