/*
 * TEST INFRASTRUCTURE — implementation of the seeded replacements declared in
 * pa_seed_shim.h.  Compiled WITHOUT the force-include, so BN_* below are the
 * real libcrypto functions.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>

#include <openssl/bn.h>
#include <openssl/sha.h>

extern "C" {
int pa_shim_rand_range(BIGNUM *rnd, const BIGNUM *range);
int pa_shim_rand(BIGNUM *rnd, int bits, int top, int bottom);
void pa_shim_select(uint64_t seed, uint64_t stream);
void pa_shim_inject_bid(uint64_t bid);
uint64_t pa_shim_draws(void);
}

namespace {
uint64_t g_seed = 0, g_stream = 0, g_draws = 0;
std::map<std::pair<uint64_t, uint64_t>, uint64_t> g_ctr;
bool g_have_bid = false;
uint64_t g_bid = 0;

void le64(unsigned char *p, uint64_t v) {
  for (int i = 0; i < 8; ++i) p[i] = (unsigned char)(v >> (8 * i));
}

/* draw(seed, stream, ctr) = SHA-256("PAv1" || LE64 seed || LE64 stream || LE64 ctr) */
void pa_draw(unsigned char out[32]) {
  unsigned char msg[28];
  memcpy(msg, "PAv1", 4);
  uint64_t &ctr = g_ctr[std::make_pair(g_seed, g_stream)];
  le64(msg + 4, g_seed);
  le64(msg + 12, g_stream);
  le64(msg + 20, ctr);
  ++ctr;
  ++g_draws;
  SHA256(msg, sizeof msg, out);
}
} // namespace

int pa_shim_rand_range(BIGNUM *rnd, const BIGNUM *range) {
  unsigned char d[32];
  if (BN_num_bits(range) != 256) {
    fprintf(stderr, "pa_shim_rand_range: only 256-bit ranges are supported\n");
    abort();
  }
  do {
    pa_draw(d);
    BN_bin2bn(d, 32, rnd);
  } while (BN_cmp(rnd, range) >= 0);
  return 1;
}

int pa_shim_rand(BIGNUM *rnd, int bits, int top, int bottom) {
  unsigned char d[32];
  if (bits != 256 || top != -1 || bottom != 0) {
    fprintf(stderr, "pa_shim_rand: only BN_rand(.,256,-1,0) is supported\n");
    abort();
  }
  pa_draw(d);
  BN_bin2bn(d, 32, rnd);
  return 1;
}

void pa_shim_select(uint64_t seed, uint64_t stream) {
  g_seed = seed;
  g_stream = stream;
}

void pa_shim_inject_bid(uint64_t bid) {
  g_have_bid = true;
  g_bid = bid;
}

uint64_t pa_shim_draws(void) { return g_draws; }

struct pa_shim_random_device {
  typedef unsigned int result_type;
  unsigned int operator()();
};
unsigned int pa_shim_random_device::operator()() { return 0x5EA1u; }

uint64_t pa_shim_next_bid(int lo, int hi) {
  (void)lo;
  (void)hi;
  if (!g_have_bid) {
    fprintf(stderr, "pa_shim: a Bidder was constructed without an injected bid\n");
    abort();
  }
  g_have_bid = false;
  return g_bid;
}
