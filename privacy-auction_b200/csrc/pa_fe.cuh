// secp256k1 base-field arithmetic: p = 2^256 - 2^32 - 977  (OpenSSL NID 714,
// reference SEAL/params.h:4).  One field element per thread, 8 x 32-bit limbs in
// registers, little-endian limb order.
//
// This replaces what libcrypto's BN_mod_mul_montgomery / ec_GFp_mont_field_mul
// do underneath every EC_POINT_* call of the reference.
// Instead of Montgomery form the special shape of p is used:
// 2^256 = C (mod p), C = 2^32 + 977, so a 512-bit product folds to 256 bits
// with 8 extra multiply-accumulates (vs 72 for a CIOS Montgomery reduction).
//
// Values are kept WEAKLY reduced: any representative in [0, 2^256).  fe_canon /
// fe_is_zero / fe_eq give canonical answers.
#pragma once
#include "pa_ptx.cuh"

struct fe {
  u32 v[8];
};

#define PA_P0 0xFFFFFC2Fu
#define PA_P1 0xFFFFFFFEu
#define PA_C0 977u  // C = 2^32 + 977

PA_HD void fe_set_zero(fe &r) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0;
}
PA_HD void fe_set_one(fe &r) {
  fe_set_zero(r);
  r.v[0] = 1;
}

// r = a + b  (weak)
PA_HD void fe_add(fe &r, const fe &a, const fe &b) {
  u32 t[8];
  t[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) t[i] = addc_cc(a.v[i], b.v[i]);
  u32 m = 0u - addc(0, 0);  // all-ones if the sum passed 2^256: fold 2^256 -> C
  t[0] = add_cc(t[0], PA_C0 & m);
  t[1] = addc_cc(t[1], 1u & m);
#pragma unroll
  for (int i = 2; i < 8; ++i) t[i] = addc_cc(t[i], 0);
  m = 0u - addc(0, 0);  // only when both inputs were >= p; remainder < C then
  t[0] = add_cc(t[0], PA_C0 & m);
  t[1] = addc(t[1], 1u & m);
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}

// r = a - b  (weak)
PA_HD void fe_sub(fe &r, const fe &a, const fe &b) {
  u32 t[8];
  t[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) t[i] = subc_cc(a.v[i], b.v[i]);
  u32 m = subc(0, 0);  // 0 - 0 - borrow: all-ones if it borrowed (wrapped by 2^256 -> subtract C)
  t[0] = sub_cc(t[0], PA_C0 & m);
  t[1] = subc_cc(t[1], 1u & m);
#pragma unroll
  for (int i = 2; i < 8; ++i) t[i] = subc_cc(t[i], 0);
  // A second borrow happens only when the difference was below C (b > p + a): the value is then
  // 2^256 - (C - t) with C - t < 2^33, so taking C off once more cannot borrow past limb 1.
  m = subc(0, 0);
  t[0] = sub_cc(t[0], PA_C0 & m);
  t[1] = subc(t[1], 1u & m);
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}

// r = a - b - c [- d] (weak) with ONE fold: the subtractions run mod 2^256 while the borrows are
// counted (n <= 3), then n * C is taken off.  If that borrows, the value is 2^256 - e with
// e < n C < 2^34, and one more C stays inside limbs 0..1.
template <int N>
PA_HD void fe_sub_n(fe &r, const fe &a, const fe *const *s) {
  u32 t[8];
  u32 q = 0;  // minus the number of borrows
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = a.v[i];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    t[0] = sub_cc(t[0], s[k]->v[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) t[i] = subc_cc(t[i], s[k]->v[i]);
    q = subc(q, 0);
  }
  u32 n = 0u - q;
  t[0] = sub_cc(t[0], n * PA_C0);
  t[1] = subc_cc(t[1], n);
#pragma unroll
  for (int i = 2; i < 8; ++i) t[i] = subc_cc(t[i], 0);
  u32 m = subc(0, 0);
  t[0] = sub_cc(t[0], PA_C0 & m);
  t[1] = subc(t[1], 1u & m);
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}
PA_HD void fe_sub2(fe &r, const fe &a, const fe &b, const fe &c) {
  const fe *s[2] = {&b, &c};
  fe_sub_n<2>(r, a, s);
}
PA_HD void fe_sub3(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
  const fe *s[3] = {&b, &c, &d};
  fe_sub_n<3>(r, a, s);
}

// r = -a (weak): p - a = ~a - (C - 1) mod 2^256.  It borrows only for a > p (a = p + e, e < C);
// the wrapped value 2^256 - e then needs C taken off once more, which stays inside limbs 0..1.
PA_HD void fe_neg(fe &r, const fe &a) {
  u32 t[8];
  t[0] = sub_cc(~a.v[0], PA_C0 - 1u);
  t[1] = subc_cc(~a.v[1], 1u);
#pragma unroll
  for (int i = 2; i < 8; ++i) t[i] = subc_cc(~a.v[i], 0);
  u32 m = subc(0, 0);
  t[0] = sub_cc(t[0], PA_C0 & m);
  t[1] = subc(t[1], 1u & m);
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}
PA_HD void fe_dbl(fe &r, const fe &a) { fe_add(r, a, a); }

// t (8 limbs) += q * C for a small q (< 2^3), result weak.  If that wraps, the remainder is
// < q * C < 2^36, so one more C stays inside limbs 0..1.
PA_HD void fe_fold_small(fe &r, u32 t[8], u32 q) {
  t[0] = add_cc(t[0], q * PA_C0);
  t[1] = addc_cc(t[1], q);
#pragma unroll
  for (int i = 2; i < 8; ++i) t[i] = addc_cc(t[i], 0);
  u32 m = 0u - addc(0, 0);
  t[0] = add_cc(t[0], PA_C0 & m);
  t[1] = addc(t[1], 1u & m);
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = t[i];
}

// r = a * 2^K (weak), K = 1..3: shift, then fold the K bits that left the top (q < 2^K) as q * C.
// One pass instead of K additions (the 8C of a doubling: 23 instructions instead of 72).
template <int K>
PA_HD void fe_shl(fe &r, const fe &a) {
  u32 t[8];
  u32 q = a.v[7] >> (32 - K);
#pragma unroll
  for (int i = 7; i > 0; --i) t[i] = (a.v[i] << K) | (a.v[i - 1] >> (32 - K));
  t[0] = a.v[0] << K;
  fe_fold_small(r, t, q);
}

// r = 3a (weak) in one pass: a + (a << 1), the top (0..2) folded as q * C
PA_HD void fe_mul3(fe &r, const fe &a) {
  u32 t[8];
  t[0] = add_cc(a.v[0], a.v[0] << 1);
#pragma unroll
  for (int i = 1; i < 8; ++i) t[i] = addc_cc(a.v[i], (a.v[i] << 1) | (a.v[i - 1] >> 31));
  u32 q = addc(a.v[7] >> 31, 0);
  fe_fold_small(r, t, q);
}

// canonical representative in [0, p)
PA_HD void fe_canon(fe &r, const fe &a) {
  u32 t[8];
  t[0] = add_cc(a.v[0], PA_C0);
  t[1] = addc_cc(a.v[1], 1u);
#pragma unroll
  for (int i = 2; i < 8; ++i) t[i] = addc_cc(a.v[i], 0);
  u32 ge = addc(0, 0);  // a + C >= 2^256  <=>  a >= p
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = ge ? t[i] : a.v[i];
}

PA_HD bool fe_is_zero(const fe &a) {
  u32 z = a.v[0] | a.v[1], q = (a.v[0] ^ PA_P0) | (a.v[1] ^ PA_P1);
#pragma unroll
  for (int i = 2; i < 8; ++i) {
    z |= a.v[i];
    q |= ~a.v[i];
  }
  return z == 0 || q == 0;
}

PA_HD bool fe_eq(const fe &a, const fe &b) {
  fe d;
  fe_sub(d, a, b);
  return fe_is_zero(d);
}

// 16-limb product by two interleaved carry chains: `ev` collects the 64-bit
// partial products that start on an even limb, `od` the ones that start on an
// odd limb (stored shifted down by one limb), so inside a chain the products
// never overlap and every step is one 32x32+64 -> 64 multiply-accumulate with
// carry-in/out (IMAD.WIDE.U32.X).  64 such MACs in total.
PA_HD void mp_mul8(u32 t[16], const u32 a[8], const u32 b[8]) {
  u32 ev[16], od[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ev[i] = od[i] = 0;
  // After row i the partial product is below 2^(32 (i + 9)), so neither accumulator can pass limb
  // i + 8.  The chain that could carry out of limb i + 7 runs second, and its carry is added into
  // the OTHER accumulator's word for limb i + 8, which the first chain has just written as the high
  // half of its top pair: one add, and the words above stay untouched zeros, so the next row's top
  // product takes a zero addend (no carry word and no zero register to pair it with).
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if ((i & 1) == 0) {
      // a[odd] * b[i] -> limb i+j (odd) = od[i+j-1 .. i+j]; top word od[i+7] is limb i+8
      od[i + 0] = mad_lo_cc(a[1], b[i], od[i + 0]);
      od[i + 1] = madc_hi_cc(a[1], b[i], od[i + 1]);
#pragma unroll
      for (int j = 3; j < 8; j += 2) {
        od[i + j - 1] = madc_lo_cc(a[j], b[i], od[i + j - 1]);
        od[i + j] = madc_hi_cc(a[j], b[i], od[i + j]);
      }
      // a[even] * b[i] -> ev[i+j .. i+j+1]; carry out of limb i+7 goes to od's limb i+8
      ev[i + 0] = mad_lo_cc(a[0], b[i], ev[i + 0]);
      ev[i + 1] = madc_hi_cc(a[0], b[i], ev[i + 1]);
#pragma unroll
      for (int j = 2; j < 8; j += 2) {
        ev[i + j] = madc_lo_cc(a[j], b[i], ev[i + j]);
        ev[i + j + 1] = madc_hi_cc(a[j], b[i], ev[i + j + 1]);
      }
      od[i + 7] = addc(od[i + 7], 0);
    } else {
      // a[odd] * b[i] -> ev[i+j .. i+j+1]; top word ev[i+8] is limb i+8
      ev[i + 1] = mad_lo_cc(a[1], b[i], ev[i + 1]);
      ev[i + 2] = madc_hi_cc(a[1], b[i], ev[i + 2]);
#pragma unroll
      for (int j = 3; j < 8; j += 2) {
        ev[i + j] = madc_lo_cc(a[j], b[i], ev[i + j]);
        ev[i + j + 1] = madc_hi_cc(a[j], b[i], ev[i + j + 1]);
      }
      // a[even] * b[i] -> limb i+j (odd) = od[i+j-1 .. i+j]; carry out of limb i+7 goes to ev's limb i+8
      od[i - 1] = mad_lo_cc(a[0], b[i], od[i - 1]);
      od[i + 0] = madc_hi_cc(a[0], b[i], od[i + 0]);
#pragma unroll
      for (int j = 2; j < 8; j += 2) {
        od[i + j - 1] = madc_lo_cc(a[j], b[i], od[i + j - 1]);
        od[i + j] = madc_hi_cc(a[j], b[i], od[i + j]);
      }
      ev[i + 8] = addc(ev[i + 8], 0);
    }
  }
  t[0] = ev[0];
  t[1] = add_cc(ev[1], od[0]);
#pragma unroll
  for (int k = 2; k < 16; ++k) t[k] = addc_cc(ev[k], od[k - 1]);
}

// fold a 512-bit value to a weak 256-bit representative: 2^256 = 2^32 + 977 (mod p)
PA_HD void fe_reduce512(fe &r, const u32 t[16]) {
  const u32 *hi = t + 8;
  u32 s[10];
  // s = lo + 977 * hi + (hi << 32) as two independent chains whose 64-bit addends are the naturally
  // aligned limb pairs of t (no re-pairing moves between the passes):
  //   e = lo + 977 * (hi[0], hi[2], hi[4], hi[6]) << (0, 64, 128, 192)
  //   o = hi + 977 * (hi[1], hi[3], hi[5], hi[7]) << (0, 64, 128, 192),  standing one limb up
  u32 e[9], o[9];
  e[0] = mad_lo_cc(hi[0], PA_C0, t[0]);
  e[1] = madc_hi_cc(hi[0], PA_C0, t[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    e[j] = madc_lo_cc(hi[j], PA_C0, t[j]);
    e[j + 1] = madc_hi_cc(hi[j], PA_C0, t[j + 1]);
  }
  e[8] = addc(0, 0);
  o[0] = mad_lo_cc(hi[1], PA_C0, hi[0]);
  o[1] = madc_hi_cc(hi[1], PA_C0, hi[1]);
#pragma unroll
  for (int j = 3; j < 8; j += 2) {
    o[j - 1] = madc_lo_cc(hi[j], PA_C0, hi[j - 1]);
    o[j] = madc_hi_cc(hi[j], PA_C0, hi[j]);
  }
  o[8] = addc(0, 0);
  s[0] = e[0];
  s[1] = add_cc(e[1], o[0]);
#pragma unroll
  for (int j = 2; j < 9; ++j) s[j] = addc_cc(e[j], o[j - 1]);
  s[9] = addc(o[8], 0);
  // second fold: q = s[9]:s[8] (< 2^34);  s[0..7] += q * (2^32 + 977) = w2:w1:w0 (< 2^67), one chain
  u32 m0 = mul_lo(s[8], PA_C0);
  u32 m1 = mul_hi(s[8], PA_C0) + s[9] * PA_C0;  // < 2^13
  u32 w1 = add_cc(m1, s[8]);
  u32 w2 = addc(s[9], 0);
  s[0] = add_cc(s[0], m0);
  s[1] = addc_cc(s[1], w1);
  s[2] = addc_cc(s[2], w2);
#pragma unroll
  for (int j = 3; j < 8; ++j) s[j] = addc_cc(s[j], 0);
  // if that wrapped, the remainder is < 2^67 and one more C stays inside limbs 0..2
  u32 m = 0u - addc(0, 0);
  s[0] = add_cc(s[0], PA_C0 & m);
  s[1] = addc_cc(s[1], 1u & m);
  s[2] = addc(s[2], 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = s[i];
}

PA_HD void fe_mul_inl(fe &r, const fe &a, const fe &b) {
  u32 t[16];
  mp_mul8(t, a.v, b.v);
  fe_reduce512(r, t);
}

// 16-limb square: the 28 cross products once, doubled, plus the 8 diagonal
// squares (36 MACs instead of 64).
PA_HD void mp_sqr8(u32 t[16], const u32 a[8]) {
  // cross products a[i]*a[j], i<j, accumulated with the same even/odd split
  u32 ev[16], od[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ev[i] = od[i] = 0;
  // Rows 0..i of the cross products sum to less than 2^(32 (i + 9)).  In row i the chain that
  // contains j = 7 ends on limbs (i+7, i+8) and cannot carry out; it runs first.  The other chain
  // ends on limbs (i+6, i+7) and its carry is added into the first chain's accumulator at limb
  // i + 8 (as in mp_mul8: no carry word, no zero register to pair it with).
#pragma unroll
  for (int i = 0; i < 7; ++i) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: landing parity of j = 7; pass 1: the other parity
      const int par = (pass == 0) ? ((i + 7) & 1) : ((i + 6) & 1);
      bool first = true;
#pragma unroll
      for (int j = i + 1; j < 8; ++j) {
        if (((i + j) & 1) != par) continue;
        if (par == 0) {
          ev[i + j] = first ? mad_lo_cc(a[j], a[i], ev[i + j]) : madc_lo_cc(a[j], a[i], ev[i + j]);
          ev[i + j + 1] = madc_hi_cc(a[j], a[i], ev[i + j + 1]);
        } else {
          od[i + j - 1] = first ? mad_lo_cc(a[j], a[i], od[i + j - 1]) : madc_lo_cc(a[j], a[i], od[i + j - 1]);
          od[i + j] = madc_hi_cc(a[j], a[i], od[i + j]);
        }
        first = false;
      }
      if (pass == 1 && !first) {  // carry out of limb i + 7 -> limb i + 8 of the other accumulator
        if (par == 0) od[i + 7] = addc(od[i + 7], 0);
        else ev[i + 8] = addc(ev[i + 8], 0);
      }
    }
  }
  // cross = ev + (od << 32)
  u32 x[16];
  x[0] = ev[0];
  x[1] = add_cc(ev[1], od[0]);
#pragma unroll
  for (int k = 2; k < 16; ++k) x[k] = addc_cc(ev[k], od[k - 1]);
  // t = 2*cross + sum a[i]^2 << 64i
  u32 top = 0;
#pragma unroll
  for (int k = 15; k > 0; --k) x[k] = (x[k] << 1) | (x[k - 1] >> 31);
  x[0] <<= 1;
  (void)top;
  t[0] = mad_lo_cc(a[0], a[0], x[0]);
  t[1] = madc_hi_cc(a[0], a[0], x[1]);
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    t[2 * i] = madc_lo_cc(a[i], a[i], x[2 * i]);
    t[2 * i + 1] = madc_hi_cc(a[i], a[i], x[2 * i + 1]);
  }
}

PA_HD void fe_sqr_inl(fe &r, const fe &a) {
  u32 t[16];
  mp_sqr8(t, a.v);
  fe_reduce512(r, t);
}

// On the device the multiplier and the squarer are real (non-inlined)
// functions taking and returning the 8 limbs in registers: a scalar
// multiplication contains ~3000 of them, and inlining each (~150 SASS
// instructions) produces hundreds of KB of code that thrashes the 32 KB
// instruction cache.  Define PA_FE_INLINE to inline them instead.
#if defined(__CUDA_ARCH__) && !defined(PA_FE_INLINE)
static __device__ __noinline__ fe fe_mul_call(fe a, fe b) {
  fe r;
  fe_mul_inl(r, a, b);
  return r;
}
static __device__ __noinline__ fe fe_sqr_call(fe a) {
  fe r;
  fe_sqr_inl(r, a);
  return r;
}
// The caller copies every operand into the callee's argument registers and every result out of
// them, and ptxas emits those copies as IMAD.MOV — on the multiplier pipe, the one these kernels are
// bound by (r01 ncu: 21 % of its cycles).  PA_OPQ_MODE makes the copies explicit as `or x, 0` with a
// zero ptxas cannot see through (a __constant__ word), which can only issue on the ALU pipe (LOP3):
// 1 = operands, 2 = results, 3 = both, 0 = leave the copies to ptxas.
#ifndef PA_OPQ_MODE
#define PA_OPQ_MODE 0
#endif
PA_D fe fe_opq(const fe &a) {
  fe r;
  u32 z = pa_opq_z;
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("or.b32 %0, %1, %2;" : "=r"(r.v[i]) : "r"(a.v[i]), "r"(z));
  return r;
}
PA_D void fe_mul(fe &r, const fe &a, const fe &b) {
  if (PA_OPQ_MODE == 3) r = fe_opq(fe_mul_call(fe_opq(a), fe_opq(b)));
  else if (PA_OPQ_MODE == 2) r = fe_opq(fe_mul_call(a, b));
  else if (PA_OPQ_MODE == 1) r = fe_mul_call(fe_opq(a), fe_opq(b));
  else r = fe_mul_call(a, b);
}
PA_D void fe_sqr(fe &r, const fe &a) {
  if (PA_OPQ_MODE == 3) r = fe_opq(fe_sqr_call(fe_opq(a)));
  else if (PA_OPQ_MODE == 2) r = fe_opq(fe_sqr_call(a));
  else if (PA_OPQ_MODE == 1) r = fe_sqr_call(fe_opq(a));
  else r = fe_sqr_call(a);
}
#else
PA_HD void fe_mul(fe &r, const fe &a, const fe &b) { fe_mul_inl(r, a, b); }
PA_HD void fe_sqr(fe &r, const fe &a) { fe_sqr_inl(r, a); }
#endif

// Two independent products per call.  The point formulas (pa_ec.cuh) are arranged in stages of
// two independent products.  Computing a pair inside ONE non-inlined function (PA_FE_PAIRS) lets
// ptxas interleave the two multiply-add chains; measured on B200 this lowers the lone-warp latency
// a little (n = 1000 auction 204 -> 191 ms of kernel time) but costs throughput (2^20 variable-base
// mults 19.9 -> 22.1 ms, more registers live), so the default issues the two calls back to back.
struct fe2 {
  fe a, b;
};
#if defined(__CUDA_ARCH__) && !defined(PA_FE_PAIRS) && !defined(PA_FE_INLINE)  // default: the two products as two calls
PA_D void fe_mul2(fe &r0, const fe &a0, const fe &b0, fe &r1, const fe &a1, const fe &b1) {
  fe x, y;
  fe_mul(x, a0, b0);
  fe_mul(y, a1, b1);
  r0 = x;
  r1 = y;
}
PA_D void fe_sqr2(fe &r0, const fe &a0, fe &r1, const fe &a1) {
  fe x, y;
  fe_sqr(x, a0);
  fe_sqr(y, a1);
  r0 = x;
  r1 = y;
}
PA_D void fe_sqrmul(fe &r0, const fe &a0, fe &r1, const fe &a1, const fe &b1) {
  fe x, y;
  fe_sqr(x, a0);
  fe_mul(y, a1, b1);
  r0 = x;
  r1 = y;
}
#elif defined(__CUDA_ARCH__) && !defined(PA_FE_INLINE)
static __device__ __noinline__ fe2 fe_mul2_call(fe a0, fe b0, fe a1, fe b1) {
  fe2 r;
  u32 t0[16], t1[16];
  mp_mul8(t0, a0.v, b0.v);
  mp_mul8(t1, a1.v, b1.v);
  fe_reduce512(r.a, t0);
  fe_reduce512(r.b, t1);
  return r;
}
static __device__ __noinline__ fe2 fe_sqr2_call(fe a0, fe a1) {
  fe2 r;
  u32 t0[16], t1[16];
  mp_sqr8(t0, a0.v);
  mp_sqr8(t1, a1.v);
  fe_reduce512(r.a, t0);
  fe_reduce512(r.b, t1);
  return r;
}
static __device__ __noinline__ fe2 fe_sqrmul_call(fe a0, fe a1, fe b1) {
  fe2 r;
  u32 t0[16], t1[16];
  mp_sqr8(t0, a0.v);
  mp_mul8(t1, a1.v, b1.v);
  fe_reduce512(r.a, t0);
  fe_reduce512(r.b, t1);
  return r;
}
// r0 = a0*b0, r1 = a1*b1
PA_D void fe_mul2(fe &r0, const fe &a0, const fe &b0, fe &r1, const fe &a1, const fe &b1) {
  fe2 t = fe_mul2_call(a0, b0, a1, b1);
  r0 = t.a;
  r1 = t.b;
}
// r0 = a0^2, r1 = a1^2
PA_D void fe_sqr2(fe &r0, const fe &a0, fe &r1, const fe &a1) {
  fe2 t = fe_sqr2_call(a0, a1);
  r0 = t.a;
  r1 = t.b;
}
// r0 = a0^2, r1 = a1*b1
PA_D void fe_sqrmul(fe &r0, const fe &a0, fe &r1, const fe &a1, const fe &b1) {
  fe2 t = fe_sqrmul_call(a0, a1, b1);
  r0 = t.a;
  r1 = t.b;
}
#else
PA_HD void fe_mul2(fe &r0, const fe &a0, const fe &b0, fe &r1, const fe &a1, const fe &b1) {
  fe x, y;
  fe_mul_inl(x, a0, b0);
  fe_mul_inl(y, a1, b1);
  r0 = x;
  r1 = y;
}
PA_HD void fe_sqr2(fe &r0, const fe &a0, fe &r1, const fe &a1) {
  fe x, y;
  fe_sqr_inl(x, a0);
  fe_sqr_inl(y, a1);
  r0 = x;
  r1 = y;
}
PA_HD void fe_sqrmul(fe &r0, const fe &a0, fe &r1, const fe &a1, const fe &b1) {
  fe x, y;
  fe_sqr_inl(x, a0);
  fe_mul_inl(y, a1, b1);
  r0 = x;
  r1 = y;
}
#endif

PA_HD void fe_sqr_n(fe &r, const fe &a, int n) {
  fe t = a;
  for (int i = 0; i < n; ++i) fe_sqr(t, t);
  r = t;
}

// r = a^(p-2): 255 squarings + 15 multiplications.  p - 2 in binary is
// 223 ones, 0, 22 ones, 0000, 1, 0, 11, 0, 1.
PA_HD void fe_inv(fe &r, const fe &a) {
  fe x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
  fe_sqr(t, a);        fe_mul(x2, t, a);
  fe_sqr(t, x2);       fe_mul(x3, t, a);
  fe_sqr_n(t, x3, 3);  fe_mul(x6, t, x3);
  fe_sqr_n(t, x6, 3);  fe_mul(x9, t, x3);
  fe_sqr_n(t, x9, 2);  fe_mul(x11, t, x2);
  fe_sqr_n(t, x11, 11); fe_mul(x22, t, x11);
  fe_sqr_n(t, x22, 22); fe_mul(x44, t, x22);
  fe_sqr_n(t, x44, 44); fe_mul(x88, t, x44);
  fe_sqr_n(t, x88, 88); fe_mul(x176, t, x88);
  fe_sqr_n(t, x176, 44); fe_mul(x220, t, x44);
  fe_sqr_n(t, x220, 3); fe_mul(x223, t, x3);
  fe_sqr_n(t, x223, 23); fe_mul(t, t, x22);
  fe_sqr_n(t, t, 5);   fe_mul(t, t, a);
  fe_sqr_n(t, t, 3);   fe_mul(t, t, x2);
  fe_sqr_n(t, t, 2);   fe_mul(r, t, a);
}

// 32-byte big-endian <-> limbs
PA_HD void fe_from_be(fe &r, const unsigned char *b) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const unsigned char *q = b + 4 * (7 - i);
    r.v[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
  }
}
PA_HD void fe_to_be(unsigned char *b, const fe &a) {  // a must be canonical
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned char *q = b + 4 * (7 - i);
    q[0] = (unsigned char)(a.v[i] >> 24);
    q[1] = (unsigned char)(a.v[i] >> 16);
    q[2] = (unsigned char)(a.v[i] >> 8);
    q[3] = (unsigned char)(a.v[i]);
  }
}
