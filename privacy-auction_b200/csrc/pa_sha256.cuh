// SHA-256 (FIPS 180-4) and the Fiat-Shamir challenge of the reference:
//   h = SHA-256( enc(p_0) || ... || enc(p_k) || LE64(id) )  as a big-endian
//   integer, reduced modulo the group order       (SEAL/hash.cpp:8-53 etc.)
// where enc() is EC_POINT_point2oct(..., POINT_CONVERSION_UNCOMPRESSED):
// 04 || X || Y, or the single byte 00 for the point at infinity,
// and p_0 is always the generator.  The engine's 64-byte wire form of a point
// is exactly the X || Y part, so hashing needs no field arithmetic at all (the
// reference pays a field inversion per hashed point inside point2oct).
//
// One thread hashes one message.  The message is streamed through a 64-bit shift register: bytes where the
// encoding demands it (the 04 / 00 prefix of a point, the id), whole big-endian words for the 64 coordinate
// bytes of a point (aligned 16-byte loads from the wire record), so a point costs 17 steps instead of 65.
// A challenge is 4..29 blocks next to ~10^6 multiply-adds of curve work, but with one proof per thread
// these kernels run at lone-warp latency, which is why the streaming cost matters.
#pragma once
#include "pa_sc.cuh"

struct sha256_state {
  u32 h[8];
  u32 w[16];   // the block being filled (indexed by pos: lives in local memory on the device)
  u32 pos;     // complete words in w
  u32 nb;      // pending bytes (0..3), held in the low bits of acc
  u64 acc;
  u32 blocks;
};

PA_HD u32 sha_rotr(u32 x, int n) { return (x >> n) | (x << (32 - n)); }

// On the device the compression function is a real function (about 2,000 instructions): it is reached from every
// place a block can fill up, and inlining it there would multiply the code.
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ void sha256_compress(u32 *h, const u32 *blk) {
#else
PA_HD void sha256_compress(u32 h[8], const u32 blk[16]) {
#endif
  const u32 K[64] = {
      0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
      0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
      0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
      0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
      0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
      0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
      0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
      0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
  u32 w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = blk[i];
  u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    u32 wi;
    if (i < 16) {
      wi = w[i];
    } else {
      u32 w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      u32 s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      u32 s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
      w[i & 15] = wi;
    }
    u32 S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    u32 ch = (e & f) ^ (~e & g);
    u32 t1 = hh + S1 + ch + K[i] + wi;
    u32 S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    u32 mj = (a & b) ^ (a & c) ^ (b & c);
    u32 t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

PA_HD void sha256_init(sha256_state &s) {
  const u32 iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
#pragma unroll
  for (int i = 0; i < 8; ++i) s.h[i] = iv[i];
  s.pos = 0;
  s.nb = 0;
  s.acc = 0;
  s.blocks = 0;
}

PA_HD void sha256_emit(sha256_state &s, u32 word) {
  s.w[s.pos] = word;
  if (++s.pos == 16) {
    sha256_compress(s.h, s.w);
    s.pos = 0;
    s.blocks++;
  }
}
// one message byte
PA_HD void sha256_put(sha256_state &s, u32 byte) {
  s.acc = (s.acc << 8) | (byte & 0xFFu);
  if (++s.nb == 4) {
    sha256_emit(s, (u32)s.acc);
    s.nb = 0;
  }
}
// four message bytes, given as the big-endian word they form
PA_HD void sha256_put_be32(sha256_state &s, u32 x) {
  s.acc = (s.acc << 32) | x;
  sha256_emit(s, (u32)(s.acc >> (8 * s.nb)));
}

PA_HD void sha256_put_bytes(sha256_state &s, const unsigned char *p, int n) {
  for (int i = 0; i < n; ++i) sha256_put(s, p[i]);
}

// one curve point in wire form (64 bytes X||Y at a 16-byte aligned address, zeros = infinity)
PA_HD void sha256_put_point(sha256_state &s, const unsigned char *p) {
  u32 v[16];
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t = q[i];
    v[4 * i] = __byte_perm(t.x, 0, 0x0123); v[4 * i + 1] = __byte_perm(t.y, 0, 0x0123);
    v[4 * i + 2] = __byte_perm(t.z, 0, 0x0123); v[4 * i + 3] = __byte_perm(t.w, 0, 0x0123);
  }
#else
  for (int i = 0; i < 16; ++i) v[i] = ((u32)p[4 * i] << 24) | ((u32)p[4 * i + 1] << 16) | ((u32)p[4 * i + 2] << 8) | (u32)p[4 * i + 3];
#endif
  u32 nz = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) nz |= v[i];
  if (!nz) {
    sha256_put(s, 0);  // EC_POINT_point2oct of infinity is the single byte 00
    return;
  }
  sha256_put(s, 4);
#pragma unroll
  for (int i = 0; i < 16; ++i) sha256_put_be32(s, v[i]);
}

PA_HD void sha256_final(sha256_state &s, u32 digest[8]) {
  u64 bits = ((u64)s.blocks * 64 + s.pos * 4 + s.nb) * 8;
  sha256_put(s, 0x80);
  while (s.nb != 0) sha256_put(s, 0);
  while (s.pos != 14) sha256_emit(s, 0);
  sha256_emit(s, (u32)(bits >> 32));
  sha256_emit(s, (u32)bits);
#pragma unroll
  for (int i = 0; i < 8; ++i) digest[i] = s.h[i];
}

// digest (big-endian words) -> scalar mod n   (BN_bin2bn + BN_mod, SEAL/hash.cpp:50-51)
PA_HD void sc_from_digest(sc &r, const u32 digest[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = digest[7 - i];
  sc_reduce(r);
}

// generator in wire form
PA_HD void wire_generator(unsigned char g[64]) {
  const unsigned char G[64] = {
      0x79, 0xBE, 0x66, 0x7E, 0xF9, 0xDC, 0xBB, 0xAC, 0x55, 0xA0, 0x62, 0x95, 0xCE, 0x87, 0x0B, 0x07,
      0x02, 0x9B, 0xFC, 0xDB, 0x2D, 0xCE, 0x28, 0xD9, 0x59, 0xF2, 0x81, 0x5B, 0x16, 0xF8, 0x17, 0x98,
      0x48, 0x3A, 0xDA, 0x77, 0x26, 0xA3, 0xC4, 0x65, 0x5D, 0xA4, 0xFB, 0xFC, 0x0E, 0x11, 0x08, 0xA8,
      0xFD, 0x17, 0xB4, 0x48, 0xA6, 0x85, 0x54, 0x19, 0x9C, 0x47, 0xD0, 0x8F, 0xFB, 0x10, 0xD4, 0xB8};
  for (int i = 0; i < 64; ++i) g[i] = G[i];
}

// h = H(g, pts[0], ..., pts[k-1], id): pts are pointers to 64-byte wire points
PA_HD void challenge_hash(sc &h, const unsigned char *const *pts, int k, u64 id) {
  sha256_state s;
  sha256_init(s);
#if defined(__CUDACC__)
  __align__(16) unsigned char g[64];
#else
  alignas(16) unsigned char g[64];
#endif
  wire_generator(g);
  sha256_put_point(s, g);
  for (int i = 0; i < k; ++i) sha256_put_point(s, pts[i]);
  for (int i = 0; i < 8; ++i) sha256_put(s, (u32)(id >> (8 * i)) & 0xFFu);  // raw little-endian size_t
  u32 d[8];
  sha256_final(s, d);
  sc_from_digest(h, d);
}

// PA deterministic draw stream (include/pa_engine.h, "seeded randomness"):
//   draw(seed, stream, ctr) = SHA-256("PAv1" || LE64 seed || LE64 stream || LE64 ctr)
// With a 32-byte entropy key installed (pa_ctx_set_entropy: what a deployment uses, the seeded
// form is for tests and benchmarks) the draw is
//   SHA-256("PAv2" || key[32] || LE64 seed || LE64 stream || LE64 ctr)
// so that nothing derivable from a published transcript reproduces a party's secrets.
struct pa_rng_cfg {
  u32 keyed;        // 1: key[] is mixed into every draw
  u32 reject_bits;  // TEST HOOK: a draw whose top `reject_bits` bits are all ones is redrawn, as if it were >= the
                    // group order (probability 2^-reject_bits instead of 2^-128), to exercise the redraw paths
  u32 key[8];
};
#if defined(__CUDACC__)
__constant__ pa_rng_cfg pa_rng_config;  // zero: the seeded test stream
#endif

PA_HD void pa_stream_draw(u32 digest[8], u64 seed, u64 stream, u64 ctr) {
  sha256_state s;
  sha256_init(s);
  sha256_put(s, 'P'); sha256_put(s, 'A'); sha256_put(s, 'v');
#if defined(__CUDA_ARCH__)
  if (pa_rng_config.keyed) {
    sha256_put(s, '2');
    for (int w = 0; w < 8; ++w)
      for (int b = 3; b >= 0; --b) sha256_put(s, (pa_rng_config.key[w] >> (8 * b)) & 0xFFu);
  } else
#endif
  {
    sha256_put(s, '1');
  }
  for (int i = 0; i < 8; ++i) sha256_put(s, (u32)(seed >> (8 * i)) & 0xFFu);
  for (int i = 0; i < 8; ++i) sha256_put(s, (u32)(stream >> (8 * i)) & 0xFFu);
  for (int i = 0; i < 8; ++i) sha256_put(s, (u32)(ctr >> (8 * i)) & 0xFFu);
  sha256_final(s, digest);
}
// BN_rand_range(., order) on the PA stream: redraw while the value is >= n
PA_HD void pa_stream_rand_range(sc &r, u64 seed, u64 stream, u64 &ctr) {
  for (;;) {
    u32 d[8];
    pa_stream_draw(d, seed, stream, ctr++);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = d[7 - i];
    bool reject = sc_ge_n(r.v);
#if defined(__CUDA_ARCH__)
    const u32 rb = pa_rng_config.reject_bits;
    if (rb) reject |= (r.v[7] >> (32 - rb)) == ((1u << rb) - 1u);
#endif
    if (!reject) return;
  }
}
