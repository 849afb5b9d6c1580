// Scalar multiplication on secp256k1 — the >99 % of the reference's run time
// that sits in EC_POINT_mul:
//   fixed base   g^k        EC_POINT_mul(group, r, k, NULL, NULL, ctx)   SEAL/bidder.cpp:98
//   variable     P^k        EC_POINT_mul(group, r, NULL, P, k, ctx)      SEAL/bidder.cpp:129
//   double       g^a P^b    EC_POINT_mul(group, r, a, P, b, ctx)         SEAL/bidder.cpp:175
//   a check      P^a Q^b    two EC_POINT_mul + EC_POINT_add              SEAL/bidder.cpp:266-268
//
// Fixed base: signed 16-bit windows.  TAB[w][d-1] = d * 2^(16w) * G in affine form for
// w = 0..16, d = 1..32768 (34 MiB, built once per context on the GPU, read through L2).  The
// scalar is recoded to signed digits by adding 2^15 to every window up front, so g^k is the sum of
// at most 17 table entries (negated when the digit is negative): no doublings.
//
// Variable base: GLV split (128 doublings instead of 256), signed 4-bit fixed
// windows over a per-thread co-Z table of 1P..8P (mixed additions), one or two
// bases sharing the doublings (Strauss).  The signed digits come from
// k' = |k| + 0x88..8: nibble_i(k') - 8 is digit i, so no per-digit carry has to
// be tracked while walking from the top.
#pragma once
#include "pa_ec.cuh"
#include "pa_sc.cuh"

#ifndef PA_COMB_BITS
#define PA_COMB_BITS 16  // measured on B200, ms per 2^20 fixed-base mults: 13 bits 2.50, 14: 2.37, 15: 2.24, 16: 2.11
#endif
// Signed windows: digit_w in [-2^(B-1), 2^(B-1)), so a window stores d * 2^(B w) G for d = 1 .. 2^(B-1)
// only (the negative is a negated y).  W * B >= 258 leaves room for the recoding carry.
#define PA_COMB_WINDOWS ((258 + PA_COMB_BITS - 1) / PA_COMB_BITS)  // 17
#define PA_COMB_ENTRIES (1 << (PA_COMB_BITS - 1))                  // 32768
#define PA_COMB_WORDS ((size_t)PA_COMB_WINDOWS * PA_COMB_ENTRIES * 16)  // u32 words (34 MiB, L2-resident)

PA_HD void comb_load(aff &q, const u32 *tab, int w, u32 d) {
  const u32 *e = tab + ((size_t)w * PA_COMB_ENTRIES + (d - 1)) * 16;  // d in [1, PA_COMB_ENTRIES]
#if defined(__CUDA_ARCH__)
  const uint4 *e4 = reinterpret_cast<const uint4 *>(e);
  uint4 a = __ldg(e4), b = __ldg(e4 + 1), c = __ldg(e4 + 2), dd = __ldg(e4 + 3);
  q.x.v[0] = a.x; q.x.v[1] = a.y; q.x.v[2] = a.z; q.x.v[3] = a.w;
  q.x.v[4] = b.x; q.x.v[5] = b.y; q.x.v[6] = b.z; q.x.v[7] = b.w;
  q.y.v[0] = c.x; q.y.v[1] = c.y; q.y.v[2] = c.z; q.y.v[3] = c.w;
  q.y.v[4] = dd.x; q.y.v[5] = dd.y; q.y.v[6] = dd.z; q.y.v[7] = dd.w;
#else
  for (int i = 0; i < 8; ++i) {
    q.x.v[i] = e[i];
    q.y.v[i] = e[8 + i];
  }
#endif
}

// r = k * G,  k < n.  k' = k + sum_w 2^(B-1) 2^(B w): window w of k' minus 2^(B-1) is the signed digit.
PA_HD void fixed_base_mul(jac &r, const sc &k, const u32 *tab) {
  u32 kp[10];
  {
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      u32 add = 0;  // bits B w + B - 1 of the recoding constant that fall into limb i
#pragma unroll
      for (int w = 0; w < PA_COMB_WINDOWS; ++w) {
        int bit = w * PA_COMB_BITS + PA_COMB_BITS - 1;
        if ((bit >> 5) == i) add |= 1u << (bit & 31);
      }
      u64 t = (u64)(i < 8 ? k.v[i] : 0u) + add + c;
      kp[i] = (u32)t;
      c = t >> 32;
    }
    kp[9] = 0;
  }
  jac_set_inf(r);
#pragma unroll 1
  for (int w = 0; w < PA_COMB_WINDOWS; ++w) {
    int bit = w * PA_COMB_BITS, limb = bit >> 5, sh = bit & 31;
    u32 d = kp[limb] >> sh;
    if (sh + PA_COMB_BITS > 32) d |= kp[limb + 1] << (32 - sh);
    int sd = (int)(d & ((1u << PA_COMB_BITS) - 1u)) - (1 << (PA_COMB_BITS - 1));
    if (sd) {
      aff q;
      comb_load(q, tab, w, (u32)(sd < 0 ? -sd : sd));
      if (sd < 0) fe_neg(q.y, q.y);
      jac_madd_t<true>(r, r, q);  // inlined once in this loop; comb entries are never infinity
    }
  }
}

// ---- variable base: GLV + co-Z window tables -------------------------------------------------
// secp256k1 has the endomorphism lambda * (x, y) = (beta * x, y).  A scalar k is split as
// k = k1 + k2 * lambda (mod n) with |k1|, |k2| < 2^128 (lattice reduction with the basis
// (a1, b1), (a2, b2) below), so only 128 doublings are needed instead of 256 and the two halves
// share them.  Per base a table 1P'..8P' is built by a chain of mixed additions and rescaled to a
// COMMON Z: on the curve isomorphic by that Z the entries are affine, so every addition in the main
// loop is a mixed addition (8M+3S instead of 12M+4S); the result's Z is multiplied by the common Z
// at the end (the formulas for a = 0 do not involve the curve constant b).
// Work per variable-base mult: ~1,800 field mults (nominal double-and-add with 4-bit windows: 2,900).
struct glv_split {
  u32 k1[5], k2[5];  // |k1|, |k2| (< 2^129), little-endian limbs
  bool neg1, neg2;
};

// acc (9 limbs, two's complement) += / -= x[0..nx) * y[0..ny)
template <int NX, int NY>
PA_HD void mp9_mac(u32 acc[9], const u32 *x, const u32 *y, bool subtract) {
  u32 prod[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) prod[i] = 0;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
      if (i + j < 9) {
        u64 t = (u64)x[i] * y[j] + prod[i + j] + c;
        prod[i + j] = (u32)t;
        c = t >> 32;
      }
    }
    if (i + NY < 9) prod[i + NY] += (u32)c;
  }
  u64 c = subtract ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    u64 t = (u64)acc[i] + (subtract ? ~prod[i] : prod[i]) + c;
    acc[i] = (u32)t;
    c = t >> 32;
  }
}

// c = round(k * g / 2^384): the top 4 limbs of the 16-limb product, rounded
PA_HD void glv_mulshift384(u32 c[5], const u32 k[8], const u32 g[8]) {
  u32 t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = 0;
  mp_addmul<16, 8, 8>(t, k, g);
  u64 carry = (t[11] >> 31) & 1u;  // bit 383
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    u64 v = (u64)t[12 + i] + carry;
    c[i] = (u32)v;
    carry = v >> 32;
  }
  c[4] = (u32)carry;
}

PA_HD void glv_decompose(glv_split &o, const sc &k) {
  // lattice basis: a1 + b1*lambda = a2 + b2*lambda = 0 (mod n); b1 < 0, a1 = b2
  const u32 A1[4] = {0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
  const u32 NB1[4] = {0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};  // -b1
  const u32 A2[5] = {0x9D44CFD8u, 0x57C1108Du, 0xA8E2F3F6u, 0x14CA50F7u, 0x1u};
  // g1 = round(2^384 * b2 / n), g2 = round(2^384 * (-b1) / n)
  const u32 G1[8] = {0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
  const u32 G2[8] = {0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};
  u32 c1[5], c2[5];
  glv_mulshift384(c1, k.v, G1);
  glv_mulshift384(c2, k.v, G2);
  // k1 = k - c1*a1 - c2*a2 ; k2 = c1*(-b1) - c2*b2   (b2 = a1), both small signed integers
  u32 r1[9], r2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    r1[i] = i < 8 ? k.v[i] : 0;
    r2[i] = 0;
  }
  mp9_mac<5, 4>(r1, c1, A1, true);
  mp9_mac<5, 5>(r1, c2, A2, true);
  mp9_mac<5, 4>(r2, c1, NB1, false);
  mp9_mac<5, 4>(r2, c2, A1, true);
  o.neg1 = (r1[8] >> 31) != 0;
  o.neg2 = (r2[8] >> 31) != 0;
  u64 c = o.neg1 ? 1 : 0, d = o.neg2 ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    u64 t = (u64)(o.neg1 ? ~r1[i] : r1[i]) + c;
    o.k1[i] = (u32)t;
    c = t >> 32;
    u64 u = (u64)(o.neg2 ? ~r2[i] : r2[i]) + d;
    o.k2[i] = (u32)u;
    d = u >> 32;
  }
}

// |k1|, |k2| < 2^128 for every k < n (the bound libsecp256k1 proves for these lattice constants; the largest
// halves seen over 2 10^5 random and rounding-boundary scalars are 0.64 2^128, tests/test_hostcheck.py), so
// kp = |k| + 0x88...8 (32 nibbles) < 2^129: 32 signed digits and a top digit that is 0 or 1 - 128 doublings.
#define PA_GLV_WINDOWS 32  // signed 4-bit windows over 128 bits; digit 32 is the carry out of the recoding

// kp = |k| + 0x888...8 (32 nibbles): digit i = nibble_i(kp) - 8 for i < 32, digit 32 = what is left above (0 or 1)
PA_HD void glv_recode(u32 kp[5], const u32 k[5]) {
  const u32 add[5] = {0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u, 0x0u};
  u64 c = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    u64 t = (u64)k[i] + add[i] + c;
    kp[i] = (u32)t;
    c = t >> 32;
  }
}
PA_HD int glv_digit(const u32 kp[5], int i) {  // i in [0, 32]
  if (i == PA_GLV_WINDOWS) return (int)(kp[4] & 15u);  // <= 1 by the bound above (the table would serve up to 8)
  return (int)((kp[i >> 3] >> ((i & 7) * 4)) & 15u) - 8;
}

struct glv_table {  // 1P'..8P' affine on the curve isomorphic by a common Z; bx = beta * x
  fe x[8], y[8], bx[8];
};

PA_HD void fe_set_beta(fe &b) {
  const u32 B[8] = {0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu};
#pragma unroll
  for (int i = 0; i < 8; ++i) b.v[i] = B[i];
}

// r = p + q, q affine, p not infinity and p != +-q (table construction); zr = Z3 / Z1
PA_HD void jac_madd_zr(jac &r, fe &zr, const jac &p, const fe &qx, const fe &qy) {
  fe zz, u2, s2, h, rr, hh, hhh, v, t;
  fe_sqr(zz, p.Z);
  fe_mul(u2, qx, zz);
  fe_mul(s2, p.Z, zz);
  fe_mul(s2, qy, s2);
  fe_sub(h, u2, p.X);
  fe_sub(rr, s2, p.Y);
  fe_sqr(hh, h);
  fe_mul(hhh, h, hh);
  fe_mul(v, p.X, hh);
  fe_mul(r.Z, p.Z, h);
  fe_sqr(t, rr);
  fe_sub3(t, t, hhh, v, v);
  fe_mul(hhh, p.Y, hhh);
  r.X = t;
  fe_sub(v, v, t);
  fe_mul(v, rr, v);
  fe_sub(r.Y, v, hhh);
  zr = h;
}

// Table of a base given as Jacobian (X, Y, Z0), not infinity.  On return zeta is the common Z
// INCLUDING Z0: the true points are (x[i], y[i], zeta).
PA_HD void glv_build_table(glv_table &T, fe &zeta, const jac &P) {
  // work on the curve isomorphic by Z0, where P is the affine point (X, Y)
  jac acc;
  fe zr[8];
  acc.X = P.X;
  acc.Y = P.Y;
  fe_set_one(acc.Z);
  T.x[0] = acc.X;
  T.y[0] = acc.Y;
  jac_dbl(acc, acc);  // 2P: Z = 2Y
  T.x[1] = acc.X;
  T.y[1] = acc.Y;
  zr[1] = acc.Z;  // Z2 / Z1
#pragma unroll 1
  for (int i = 2; i < 8; ++i) {
    jac_madd_zr(acc, zr[i], acc, P.X, P.Y);
    T.x[i] = acc.X;
    T.y[i] = acc.Y;
  }
  // bring everything to Z8: ratio_i = Z8 / Z_i = zr[i+1] * ... * zr[7]
  fe ratio, r2, r3, beta;
  fe_set_beta(beta);
  ratio = zr[7];
#pragma unroll 1
  for (int i = 6; i >= 0; --i) {
    fe_sqr(r2, ratio);
    fe_mul(r3, r2, ratio);
    fe_mul(T.x[i], T.x[i], r2);
    fe_mul(T.y[i], T.y[i], r3);
    if (i > 0) fe_mul(ratio, ratio, zr[i]);
  }
#pragma unroll 1
  for (int i = 0; i < 8; ++i) fe_mul(T.bx[i], T.x[i], beta);
  fe_mul(zeta, acc.Z, P.Z);
}

// rescale a table from its common Z to (its common Z) * f
PA_HD void glv_scale_table(glv_table &T, const fe &f) {
  fe f2, f3;
  fe_sqr(f2, f);
  fe_mul(f3, f2, f);
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    fe_mul(T.x[i], T.x[i], f2);
    fe_mul(T.bx[i], T.bx[i], f2);
    fe_mul(T.y[i], T.y[i], f3);
  }
}

// r = a*P (+ b*Q when NB == 2), scalars < n, bases Jacobian (may be infinity)
template <int NB>
PA_HD void strauss(jac &r, const jac &P, const sc &a, const jac &Q, const sc &b) {
  glv_table TP, TQ;
  glv_split sa, sb;
  u32 a1[5], a2[5], b1[5], b2[5];
  fe zeta, zq;
  bool useP = !jac_is_inf(P) && !sc_is_zero(a);
  bool useQ = NB == 2 && !jac_is_inf(Q) && !sc_is_zero(b);
  fe_set_one(zeta);
  if (useP) {
    glv_build_table(TP, zeta, P);
    glv_decompose(sa, a);
    glv_recode(a1, sa.k1);
    glv_recode(a2, sa.k2);
  }
  if (useQ) {
    glv_build_table(TQ, zq, Q);
    glv_decompose(sb, b);
    glv_recode(b1, sb.k1);
    glv_recode(b2, sb.k2);
    if (useP) {  // one common Z for both tables: zeta_P * zeta_Q
      glv_scale_table(TP, zq);
      glv_scale_table(TQ, zeta);
      fe_mul(zeta, zeta, zq);
    } else {
      zeta = zq;
    }
  }
  jac_set_inf(r);
  if (!useP && !useQ) return;
  // One doubling body and one mixed-addition body in the loop (point formulas inlined here, field
  // products still shared calls): the 2*NB digit sources go through the same addition code, which
  // keeps the hot loop small and saves the register shuffling of a call per point operation.
#pragma unroll 1
  for (int i = PA_GLV_WINDOWS; i >= 0; --i) {
    if (i != PA_GLV_WINDOWS) {
#pragma unroll 1
      for (int j = 0; j < 4; ++j) jac_dbl_inl(r, r);
    }
#pragma unroll 1
    for (int t = 0; t < 2 * NB; ++t) {
      bool second = t >= 2, lam = t & 1;
      if (second ? !useQ : !useP) continue;
      const glv_table &T = second ? TQ : TP;
      const u32 *kk = second ? (lam ? b2 : b1) : (lam ? a2 : a1);
      bool neg = second ? (lam ? sb.neg2 : sb.neg1) : (lam ? sa.neg2 : sa.neg1);
      int d = glv_digit(kk, i);
      if (d == 0) continue;
      int idx = (d > 0 ? d : -d) - 1;
      aff q;
      q.x = lam ? T.bx[idx] : T.x[idx];
      if ((d < 0) != neg) fe_neg(q.y, T.y[idx]); else q.y = T.y[idx];
      jac_madd_t<true>(r, r, q);  // k * P' is never infinity for k = 1..8 (prime order)
    }
  }
  if (!jac_is_inf(r)) fe_mul(r.Z, r.Z, zeta);  // back from the isomorphic curve
}

PA_HD void var_base_mul(jac &r, const jac &P, const sc &k) { strauss<1>(r, P, k, P, k); }

// ---- comb table construction (GPU, once per context; also host-checked) ----
// phase 1: B_w = 2^(PA_COMB_BITS * w) G, affine, for every window w
PA_HD void comb_base(aff &out, int w, const aff &G) {
  jac p;
  jac_from_aff(p, G);
  for (int i = 0; i < PA_COMB_BITS * w; ++i) jac_dbl(p, p);
  jac_to_aff(out, p);
}
// phase 2: entry d of window w from B_w
PA_HD void comb_entry(aff &out, u32 d, const aff &Bw) {  // d in [1, PA_COMB_ENTRIES]
  jac r;
  jac_set_inf(r);
  for (int bit = PA_COMB_BITS - 1; bit >= 0; --bit) {
    jac_dbl(r, r);
    if ((d >> bit) & 1u) jac_madd(r, r, Bw);
  }
  jac_to_aff(out, r);
}

// secp256k1 generator (SEC 2, section 2.4.1; EC_GROUP_get0_generator, SEAL/bidder.cpp:39)
PA_HD void aff_set_generator(aff &g) {
  const u32 gx[8] = {0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu, 0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu};
  const u32 gy[8] = {0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u, 0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    g.x.v[i] = gx[i];
    g.y.v[i] = gy[i];
  }
}
