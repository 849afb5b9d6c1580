// Scalar multiplication on secp256k1 — the >99 % of the reference's run time
// that sits in EC_POINT_mul:
//   fixed base   g^k        EC_POINT_mul(group, r, k, NULL, NULL, ctx)   SEAL/bidder.cpp:98
//   variable     P^k        EC_POINT_mul(group, r, NULL, P, k, ctx)      SEAL/bidder.cpp:129
//   double       g^a P^b    EC_POINT_mul(group, r, a, P, b, ctx)         SEAL/bidder.cpp:175
//   a check      P^a Q^b    two EC_POINT_mul + EC_POINT_add              SEAL/bidder.cpp:266-268
//
// Fixed base: 8-bit comb.  TAB[w][d] = d * 2^(8w) * G in affine form for
// w < 32, d < 256 (512 KiB, L2-resident, built once per context on the GPU);
// g^k is 32 mixed additions and no doublings.
//
// Variable base: signed 4-bit fixed windows over a per-thread table of
// 1P..8P, one or two bases sharing the 256 doublings (Strauss).  The signed
// digits come from k' = k + 0x88..8: nibble_i(k') - 8 is digit i, so no
// per-digit carry has to be tracked while walking from the top.
#pragma once
#include "pa_ec.cuh"
#include "pa_sc.cuh"

#define PA_COMB_WINDOWS 32
#define PA_COMB_ENTRIES 256
#define PA_COMB_WORDS (PA_COMB_WINDOWS * PA_COMB_ENTRIES * 16)  // u32 words

PA_HD void comb_load(aff &q, const u32 *tab, int w, u32 d) {
  const u32 *e = tab + ((size_t)(w * PA_COMB_ENTRIES) + d) * 16;
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

// r = k * G,  k < n
PA_HD void fixed_base_mul(jac &r, const sc &k, const u32 *tab) {
  jac_set_inf(r);
#pragma unroll 1
  for (int w = 0; w < PA_COMB_WINDOWS; ++w) {
    u32 d = (k.v[w >> 2] >> ((w & 3) * 8)) & 0xFFu;
    if (d) {
      aff q;
      comb_load(q, tab, w, d);
      jac_madd(r, r, q);
    }
  }
}

// T[i] = (i + 1) * P
PA_HD void smul_table8(jac *T, const jac &P) {
  T[0] = P;
  jac_dbl(T[1], P);
#pragma unroll 1
  for (int i = 2; i < 8; ++i) jac_add(T[i], T[i - 1], P);
}

// kp = k + 0x8888...8  (9 limbs)
PA_HD void smul_recode(u32 kp[9], const sc &k) {
  u64 c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u64 s = (u64)k.v[i] + 0x88888888u + c;
    kp[i] = (u32)s;
    c = s >> 32;
  }
  kp[8] = (u32)c;
}
PA_HD int smul_digit(const u32 kp[9], int i) {  // i in [0, 64]
  if (i == 64) return (int)kp[8];
  return (int)((kp[i >> 3] >> ((i & 7) * 4)) & 15u) - 8;
}

PA_HD void smul_add_digit(jac &r, const jac *T, int d) {
  if (d > 0) {
    jac_add(r, r, T[d - 1]);
  } else if (d < 0) {
    jac t;
    jac_neg(t, T[-d - 1]);
    jac_add(r, r, t);
  }
}

// r = a*P (+ b*Q when NB == 2), scalars < n, bases Jacobian (may be infinity)
template <int NB>
PA_HD void strauss(jac &r, const jac &P, const sc &a, const jac &Q, const sc &b) {
  jac TP[8], TQ[NB == 2 ? 8 : 1];
  u32 ka[9], kb[9];
  smul_table8(TP, P);
  smul_recode(ka, a);
  if (NB == 2) {
    smul_table8(TQ, Q);
    smul_recode(kb, b);
  }
  jac_set_inf(r);
#pragma unroll 1
  for (int i = 64; i >= 0; --i) {
    if (i != 64) {
      jac_dbl(r, r);
      jac_dbl(r, r);
      jac_dbl(r, r);
      jac_dbl(r, r);
    }
    smul_add_digit(r, TP, smul_digit(ka, i));
    if (NB == 2) smul_add_digit(r, TQ, smul_digit(kb, i));
  }
}

PA_HD void var_base_mul(jac &r, const jac &P, const sc &k) { strauss<1>(r, P, k, P, k); }

// ---- comb table construction (GPU, once per context; also host-checked) ----
// phase 1: B_w = 2^(8w) G, affine, for w in [0, 32)
PA_HD void comb_base(aff &out, int w, const aff &G) {
  jac p;
  jac_from_aff(p, G);
  for (int i = 0; i < 8 * w; ++i) jac_dbl(p, p);
  jac_to_aff(out, p);
}
// phase 2: entry d of window w from B_w
PA_HD void comb_entry(aff &out, u32 d, const aff &Bw) {
  jac r;
  jac_set_inf(r);
  for (int bit = 7; bit >= 0; --bit) {
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
