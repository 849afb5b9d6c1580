// Scalar arithmetic modulo the secp256k1 group order
//   n = FFFFFFFF FFFFFFFF FFFFFFFF FFFFFFFE BAAEDCE6 AF48A03B BFD25E8C D0364141.
// Replaces the BN_mod_mul / BN_mod_sub / BN_mod calls of the reference's
// provers (SEAL/bidder.cpp:102-103, 209-216, 425-435, 826-860) and the
// "hash mod order" step of SEAL/hash.cpp:50-51.  A handful of these run per
// proof next to thousands of field multiplications, so this code is written
// for clarity (2^256 = d (mod n) folding with 64-bit accumulators), not speed.
#pragma once
#include "pa_ptx.cuh"

struct sc {
  u32 v[8];
};

// n and d = 2^256 - n, little-endian limbs.  Declared as local constant arrays
// inside each function so that, after unrolling, they become immediates.
#define PA_N_LIMBS 0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu
#define PA_ND_LIMBS 0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u, 0x1u

PA_HD void sc_set_zero(sc &r) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0;
}
PA_HD bool sc_is_zero(const sc &a) {
  u32 z = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) z |= a.v[i];
  return z == 0;
}
PA_HD bool sc_eq(const sc &a, const sc &b) {
  u32 z = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) z |= a.v[i] ^ b.v[i];
  return z == 0;
}

// t (8 limbs) >= n ?
PA_HD bool sc_ge_n(const u32 *t) {
  const u32 PA_N[8] = {PA_N_LIMBS};
  u64 br = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u64 s = (u64)t[i] - PA_N[i] - br;
    br = (s >> 63) & 1;
  }
  return br == 0;
}
PA_HD void sc_sub_n(u32 *t) {
  const u32 PA_N[8] = {PA_N_LIMBS};
  u64 br = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u64 s = (u64)t[i] - PA_N[i] - br;
    t[i] = (u32)s;
    br = (s >> 63) & 1;
  }
}

// any 256-bit value -> [0, n)   (CCS22 draws unreduced 256-bit scalars)
PA_HD void sc_reduce(sc &r) {
  if (sc_ge_n(r.v)) sc_sub_n(r.v);
}

PA_HD void sc_add(sc &r, const sc &a, const sc &b) {  // a, b < n
  u32 t[8];
  u64 c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u64 s = (u64)a.v[i] + b.v[i] + c;
    t[i] = (u32)s;
    c = s >> 32;
  }
  if (c || sc_ge_n(t)) sc_sub_n(t);  // on carry the wrapped subtraction is still exact
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}

PA_HD void sc_sub(sc &r, const sc &a, const sc &b) {  // a, b < n
  u32 t[8];
  u64 br = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u64 s = (u64)a.v[i] - b.v[i] - br;
    t[i] = (u32)s;
    br = (s >> 63) & 1;
  }
  if (br) {
    const u32 PA_N[8] = {PA_N_LIMBS};
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      u64 s = (u64)t[i] + PA_N[i] + c;
      t[i] = (u32)s;
      c = s >> 32;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}

PA_HD void sc_neg(sc &r, const sc &a) {
  sc z;
  sc_set_zero(z);
  sc_sub(r, z, a);
}

// acc[0..nacc) += x[0..nx) * y[0..ny)
template <int NACC, int NX, int NY>
PA_HD void mp_addmul(u32 *acc, const u32 *x, const u32 *y) {
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
      u64 s = (u64)x[i] * y[j] + acc[i + j] + c;
      acc[i + j] = (u32)s;
      c = s >> 32;
    }
#pragma unroll
    for (int k = i + NY; k < NACC; ++k) {
      u64 s = (u64)acc[k] + c;
      acc[k] = (u32)s;
      c = s >> 32;
    }
  }
}

// 512-bit t -> t mod n
PA_HD void sc_reduce512(sc &r, const u32 t[16]) {
  const u32 PA_ND[5] = {PA_ND_LIMBS};
  u32 a1[14], a2[12], a3[9];
#pragma unroll
  for (int i = 0; i < 14; ++i) a1[i] = i < 8 ? t[i] : 0;
  mp_addmul<14, 8, 5>(a1, t + 8, PA_ND);  // < 2^386
#pragma unroll
  for (int i = 0; i < 12; ++i) a2[i] = i < 8 ? a1[i] : 0;
  mp_addmul<12, 6, 5>(a2, a1 + 8, PA_ND);  // < 2^260
#pragma unroll
  for (int i = 0; i < 9; ++i) a3[i] = i < 8 ? a2[i] : 0;
  mp_addmul<9, 4, 5>(a3, a2 + 8, PA_ND);  // < 2^256 + 2^134
  u32 top = a3[8];                        // 0 or 1
  a3[8] = 0;
  u32 one[1] = {top};
  mp_addmul<9, 1, 5>(a3, one, PA_ND);  // no carry out of limb 7 now
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = a3[i];
  sc_reduce(r);
}

PA_HD void sc_mul(sc &r, const sc &a, const sc &b) {
  u32 t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = 0;
  mp_addmul<16, 8, 8>(t, a.v, b.v);
  sc_reduce512(r, t);
}

PA_HD void sc_from_be(sc &r, const unsigned char *b) {  // reduces mod n
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const unsigned char *q = b + 4 * (7 - i);
    r.v[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
  }
  sc_reduce(r);
}
PA_HD void sc_to_be(unsigned char *b, const sc &a) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned char *q = b + 4 * (7 - i);
    q[0] = (unsigned char)(a.v[i] >> 24);
    q[1] = (unsigned char)(a.v[i] >> 16);
    q[2] = (unsigned char)(a.v[i] >> 8);
    q[3] = (unsigned char)(a.v[i]);
  }
}
