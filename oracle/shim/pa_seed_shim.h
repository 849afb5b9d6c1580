/*
 * TEST INFRASTRUCTURE — force-included (-include) in front of the UNMODIFIED
 * reference sources (/root/reference/{SEAL,CCS22}/*.cpp) when they are compiled
 * into oracle/_ref/.  It does three things and nothing else:
 *
 *  1. pulls in the standard headers the reference forgets (<bitset>,
 *     <algorithm>; libc++ on the author's machine included them transitively,
 *     SURVEY.md §8c);
 *  2. renames the reference's sources of randomness at compile time so that a
 *     run is reproducible:
 *        BN_rand_range(bn, range)        -> pa_shim_rand_range   (SEAL/bidder.cpp:97 ...)
 *        BN_rand(bn, 256, -1, 0)         -> pa_shim_rand         (CCS22/bidder.cpp:170 ...)
 *        std::random_device              -> pa_shim_random_device (SEAL/bidder.cpp:27)
 *        std::uniform_int_distribution   -> pa_shim_bid_dist      (SEAL/bidder.cpp:29)
 *     A compile-time rename is used instead of RAND_set_rand_method because
 *     EC_POINT_mul itself draws blinding randomness from the global DRBG
 *     (SURVEY.md §0 trap 2), which would entangle the protocol stream with
 *     libcrypto internals;
 *  3. exposes pa_shim_* control functions to our own driver (select the stream
 *     of the party that is about to act, inject the bid of the next constructed
 *     Bidder).
 *
 * The deterministic stream is the "PA stream" that the engine's own
 * pa_rng_fill kernel implements (include/pa_engine.h):
 *     draw(seed, stream, ctr) = SHA-256("PAv1" || LE64(seed) || LE64(stream) || LE64(ctr))
 * read as a big-endian integer.  BN_rand_range redraws (ctr+1) while the value
 * is >= range; BN_rand(256 bits) takes the value unreduced.
 */
#ifndef PA_SEED_SHIM_H
#define PA_SEED_SHIM_H

#include <algorithm>
#include <atomic>
#include <bitset>
#include <cassert>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/rand.h>

extern "C" {
int pa_shim_rand_range(BIGNUM *rnd, const BIGNUM *range);
int pa_shim_rand(BIGNUM *rnd, int bits, int top, int bottom);
/* select (seed, stream); each stream keeps its own counter across switches */
void pa_shim_select(uint64_t seed, uint64_t stream);
/* the next Bidder constructed gets exactly this bid (consumed once) */
void pa_shim_inject_bid(uint64_t bid);
uint64_t pa_shim_draws(void); /* total draws so far (diagnostics) */
}

struct pa_shim_random_device {
  typedef unsigned int result_type;
  unsigned int operator()();
};

uint64_t pa_shim_next_bid(int lo, int hi);

template <class T> struct pa_shim_bid_dist {
  T lo_, hi_;
  pa_shim_bid_dist(T lo, T hi) : lo_(lo), hi_(hi) {}
  template <class G> T operator()(G &) {
    return (T)pa_shim_next_bid((int)lo_, (int)hi_);
  }
};

#define BN_rand_range pa_shim_rand_range
#define BN_rand pa_shim_rand
#define random_device pa_shim_random_device
#define uniform_int_distribution pa_shim_bid_dist

#endif
