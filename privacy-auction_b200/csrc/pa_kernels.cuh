// CUDA kernels of the engine (sm_100a).  One thread owns one curve point /
// one scalar multiplication; all 256-bit arithmetic stays in registers
// (pa_fe.cuh), per-thread window tables live in local memory (L1-resident),
// the generator's comb table is read through the read-only path from L2.
//
// Intermediate results travel between kernels as Jacobian triples in a scratch
// buffer (24 x u32 per point, little-endian limbs) and are converted to the
// 64-byte affine wire form by k_normalize, which shares one field inversion
// among the points of a thread (Montgomery's trick) — the reference pays one
// inversion per point inside EC_POINT_point2oct (25 us each).
#pragma once
#include "pa_smul.cuh"

#define PA_BLOCK 128

// ---- loads / stores ---------------------------------------------------------
PA_D u32 pa_bswap(u32 x) { return __byte_perm(x, 0, 0x0123); }

// 32 big-endian bytes (16-byte aligned) -> 8 little-endian limbs
PA_D void ld_be32(u32 v[8], const unsigned char *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 hi = q[0], lo = q[1];
  v[7] = pa_bswap(hi.x); v[6] = pa_bswap(hi.y); v[5] = pa_bswap(hi.z); v[4] = pa_bswap(hi.w);
  v[3] = pa_bswap(lo.x); v[2] = pa_bswap(lo.y); v[1] = pa_bswap(lo.z); v[0] = pa_bswap(lo.w);
}
PA_D void st_be32(unsigned char *p, const u32 v[8]) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(pa_bswap(v[7]), pa_bswap(v[6]), pa_bswap(v[5]), pa_bswap(v[4]));
  q[1] = make_uint4(pa_bswap(v[3]), pa_bswap(v[2]), pa_bswap(v[1]), pa_bswap(v[0]));
}
PA_D void ld_aff(aff &a, const unsigned char *p) {
  ld_be32(a.x.v, p);
  ld_be32(a.y.v, p + 32);
}
PA_D void st_aff(unsigned char *p, const aff &a) {
  st_be32(p, a.x.v);
  st_be32(p + 32, a.y.v);
}
PA_D void ld_sc(sc &k, const unsigned char *p) {
  ld_be32(k.v, p);
  sc_reduce(k);
}
PA_D void st_sc(unsigned char *p, const sc &k) { st_be32(p, k.v); }
PA_D void ld_point_jac(jac &P, const unsigned char *p) {
  aff a;
  ld_aff(a, p);
  jac_from_aff(P, a);
}
PA_D void st_jac(u32 *dst, const jac &r) {
  uint4 *q = reinterpret_cast<uint4 *>(dst);
  q[0] = make_uint4(r.X.v[0], r.X.v[1], r.X.v[2], r.X.v[3]);
  q[1] = make_uint4(r.X.v[4], r.X.v[5], r.X.v[6], r.X.v[7]);
  q[2] = make_uint4(r.Y.v[0], r.Y.v[1], r.Y.v[2], r.Y.v[3]);
  q[3] = make_uint4(r.Y.v[4], r.Y.v[5], r.Y.v[6], r.Y.v[7]);
  q[4] = make_uint4(r.Z.v[0], r.Z.v[1], r.Z.v[2], r.Z.v[3]);
  q[5] = make_uint4(r.Z.v[4], r.Z.v[5], r.Z.v[6], r.Z.v[7]);
}
PA_D void ld_fe(fe &f, const u32 *src) {
  const uint4 *q = reinterpret_cast<const uint4 *>(src);
  uint4 a = q[0], b = q[1];
  f.v[0] = a.x; f.v[1] = a.y; f.v[2] = a.z; f.v[3] = a.w;
  f.v[4] = b.x; f.v[5] = b.y; f.v[6] = b.z; f.v[7] = b.w;
}
PA_D void st_fe(u32 *dst, const fe &f) {
  uint4 *q = reinterpret_cast<uint4 *>(dst);
  q[0] = make_uint4(f.v[0], f.v[1], f.v[2], f.v[3]);
  q[1] = make_uint4(f.v[4], f.v[5], f.v[6], f.v[7]);
}
PA_D void ld_jac(jac &r, const u32 *src) {
  ld_fe(r.X, src);
  ld_fe(r.Y, src + 8);
  ld_fe(r.Z, src + 16);
}

// ---- comb table -------------------------------------------------------------
__global__ void k_comb_base(u32 *bases) {  // 32 threads: B_w = 2^(8w) G
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= PA_COMB_WINDOWS) return;
  aff G, B;
  aff_set_generator(G);
  comb_base(B, w, G);
  st_fe(bases + 16 * w, B.x);
  st_fe(bases + 16 * w + 8, B.y);
}
__global__ void k_comb_entries(const u32 *bases, u32 *tab) {  // 32*256 threads
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= PA_COMB_WINDOWS * PA_COMB_ENTRIES) return;
  int w = t / PA_COMB_ENTRIES;
  u32 d = t % PA_COMB_ENTRIES;
  aff B, e;
  ld_fe(B.x, bases + 16 * w);
  ld_fe(B.y, bases + 16 * w + 8);
  if (d == 0) {
    aff_set_inf(e);
  } else {
    comb_entry(e, d, B);
  }
  st_fe(tab + (size_t)t * 16, e.x);
  st_fe(tab + (size_t)t * 16 + 8, e.y);
}

// ---- scalar multiplication -----------------------------------------------------
__global__ void __launch_bounds__(PA_BLOCK)
k_fixed_base(const unsigned char *scalars, const u32 *__restrict__ tab, u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc k;
  ld_sc(k, scalars + 32 * (size_t)i);
  jac r;
  fixed_base_mul(r, k, tab);
  st_jac(jout + 24 * (size_t)i, r);
}

__global__ void __launch_bounds__(PA_BLOCK)
k_var_base(const unsigned char *points, const unsigned char *scalars, u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac P, r;
  sc k;
  ld_point_jac(P, points + 64 * (size_t)i);
  ld_sc(k, scalars + 32 * (size_t)i);
  var_base_mul(r, P, k);
  st_jac(jout + 24 * (size_t)i, r);
}

__global__ void __launch_bounds__(PA_BLOCK)
k_double_mul(const unsigned char *a, const unsigned char *points, const unsigned char *b,
             const u32 *__restrict__ tab, u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac P, r, g;
  sc k;
  ld_point_jac(P, points + 64 * (size_t)i);
  ld_sc(k, b + 32 * (size_t)i);
  var_base_mul(r, P, k);
  ld_sc(k, a + 32 * (size_t)i);
  fixed_base_mul(g, k, tab);
  jac_add(r, r, g);
  st_jac(jout + 24 * (size_t)i, r);
}

__global__ void __launch_bounds__(PA_BLOCK)
k_lincomb2(const unsigned char *p, const unsigned char *a, const unsigned char *q, const unsigned char *b,
           u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac P, Q, r;
  sc ka, kb;
  ld_point_jac(P, p + 64 * (size_t)i);
  ld_point_jac(Q, q + 64 * (size_t)i);
  ld_sc(ka, a + 32 * (size_t)i);
  ld_sc(kb, b + 32 * (size_t)i);
  strauss<2>(r, P, ka, Q, kb);
  st_jac(jout + 24 * (size_t)i, r);
}

__global__ void __launch_bounds__(PA_BLOCK)
k_point_add(const unsigned char *p, const unsigned char *q, u32 *jout, int n, int sub) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac P;
  aff Q;
  ld_point_jac(P, p + 64 * (size_t)i);
  ld_aff(Q, q + 64 * (size_t)i);
  if (sub) aff_neg(Q, Q);
  jac_madd(P, P, Q);
  st_jac(jout + 24 * (size_t)i, P);
}

// ---- Jacobian -> 64-byte affine, one inversion per thread ------------------------
// Thread t owns points t, t + T, t + 2T, ... (coalesced across the warp).
__global__ void __launch_bounds__(PA_BLOCK)
k_normalize(const u32 *jin, u32 *prefix, unsigned char *out, int n, int T) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  fe acc, z;
  fe_set_one(acc);
  int last = -1;
  for (int idx = t; idx < n; idx += T) {
    ld_fe(z, jin + 24 * (size_t)idx + 16);
    if (!fe_is_zero(z)) fe_mul(acc, acc, z);
    st_fe(prefix + 8 * (size_t)idx, acc);
    last = idx;
  }
  if (last < 0) return;
  fe inv;
  fe_inv(inv, acc);
  for (int idx = last; idx >= 0; idx -= T) {
    jac p;
    ld_jac(p, jin + 24 * (size_t)idx);
    aff a;
    if (fe_is_zero(p.Z)) {
      aff_set_inf(a);
    } else {
      fe zi;
      if (idx - T >= 0) {
        fe prev;
        ld_fe(prev, prefix + 8 * (size_t)(idx - T));
        fe_mul(zi, inv, prev);
      } else {
        zi = inv;
      }
      fe_mul(inv, inv, p.Z);
      jac_to_aff_with_zinv(a, p, zi);
    }
    st_aff(out + 64 * (size_t)idx, a);
  }
}

// ---- EC_POINT_point2oct ------------------------------------------------------------
__global__ void k_encode(const unsigned char *points, int n, int compressed, unsigned char *out, size_t stride,
                         u32 *lens) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char *p = points + 64 * (size_t)i;
  unsigned char *o = out + stride * (size_t)i;
  u32 nz = 0;
  for (int j = 0; j < 64; ++j) nz |= p[j];
  u32 len;
  if (!nz) {
    o[0] = 0;
    len = 1;
  } else if (compressed) {
    o[0] = 2 + (p[63] & 1);
    for (int j = 0; j < 32; ++j) o[1 + j] = p[j];
    len = 33;
  } else {
    o[0] = 4;
    for (int j = 0; j < 64; ++j) o[1 + j] = p[j];
    len = 65;
  }
  for (size_t j = len; j < stride; ++j) o[j] = 0;
  lens[i] = len;
}

// ---- integer-pipe microbenchmarks (register only) ------------------------------------
__global__ void k_peak_imad(u32 *sink, int iters, u32 a, u32 b) {
  u32 x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
      x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
  }
  u32 s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  if (s == 0x12345u) sink[0] = s;
}
__global__ void k_peak_imad_wide(u64 *sink, int iters, u32 a) {
  u64 x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = (u64)(u32)x0 * a + x0; x1 = (u64)(u32)x1 * a + x1; x2 = (u64)(u32)x2 * a + x2; x3 = (u64)(u32)x3 * a + x3;
      x4 = (u64)(u32)x4 * a + x4; x5 = (u64)(u32)x5 * a + x5; x6 = (u64)(u32)x6 * a + x6; x7 = (u64)(u32)x7 * a + x7;
    }
  }
  u64 s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  if (s == 0x12345u) sink[0] = s;
}
__global__ void k_peak_fe(u32 *sink, int iters, int sqr) {
  fe x, y;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x.v[i] = threadIdx.x * 2654435761u + i;
    y.v[i] = blockIdx.x * 40503u + i * 7u + 1u;
  }
  if (sqr) {
#pragma unroll 1
    for (int i = 0; i < iters; ++i) { fe_sqr(x, x); fe_sqr(y, y); }
  } else {
#pragma unroll 1
    for (int i = 0; i < iters; ++i) { fe_mul(x, x, y); fe_mul(y, y, x); }
  }
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x.v[i] ^ y.v[i];
  if (s == 0x12345u) sink[0] = s;
}
