// Process-wide handle on the CUDA engine (one pa_ctx on one GPU) shared by the
// host-side Bidder / BulletinBoard objects, plus the seeded draw stream that
// replaces the reference's BN_rand_range calls.  Everything cryptographic is
// delegated to the C ABI of include/pa_engine.h; if the engine cannot be
// created (no B200, library missing) the program aborts — there is no CPU path.
#ifndef PA_HOST_ENGINE_H
#define PA_HOST_ENGINE_H

#include "../../include/pa_engine.h"
#include "types.h"

#include <cstdint>
#include <vector>

namespace pa_host {

struct Config {
  // seeded = false (the default): every draw is keyed with 32 bytes from the operating system's entropy pool
  // (pa_ctx_set_entropy) and bids come from std::random_device, as the reference draws from OpenSSL's DRBG and
  // std::random_device (SEAL/bidder.cpp:27, :97) - nothing in a transcript reproduces a secret.
  // seeded = true (--seed S on the command lines; tests, benchmarks): bidder j draws from PA stream (seed, j),
  // reproducible and therefore NOT private.
  bool seeded = false;
  uint64_t seed = 1;
  int device = 0;
};
// a 64-bit value for the parties' bids / the evaluator's id: from the seed when seeded, else std::random_device
uint64_t bid_entropy(uint64_t salt);
Config &config();

pa_ctx *engine();  // created on first use
void check(int rc, const char *what);

// k consecutive BN_rand_range(., order) replacements from stream (seed, stream),
// starting at *counter, which is advanced
std::vector<Scalar> draw(uint64_t stream, uint64_t *counter, size_t k);

}  // namespace pa_host

#endif
