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
  uint64_t seed = 1;  // seed of the PA stream; bidder j draws from stream (seed, j)
  int device = 0;
};
Config &config();

pa_ctx *engine();  // created on first use
void check(int rc, const char *what);

// k consecutive BN_rand_range(., order) replacements from stream (seed, stream),
// starting at *counter, which is advanced
std::vector<Scalar> draw(uint64_t stream, uint64_t *counter, size_t k);

}  // namespace pa_host

#endif
