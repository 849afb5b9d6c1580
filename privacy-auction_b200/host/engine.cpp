#include "engine.h"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <sys/random.h>

namespace pa_host {

Config &config() {
  static Config c;
  return c;
}

pa_ctx *engine() {
  static pa_ctx *ctx = nullptr;
  if (!ctx) {
    int rc = pa_ctx_create(&ctx, config().device);
    if (rc != PA_OK) {
      fprintf(stderr, "pa_ctx_create failed (%d): %s\n", rc, pa_last_error(nullptr));
      exit(1);
    }
    if (!config().seeded) {  // a deployment: key the draw stream from the OS entropy pool; never written anywhere
      uint8_t key[32];
      if (getrandom(key, sizeof key, 0) != (ssize_t)sizeof key) {
        fprintf(stderr, "getrandom failed: refusing to run with predictable randomness (use --seed for a test run)\n");
        exit(1);
      }
      check(pa_ctx_set_entropy(ctx, key), "pa_ctx_set_entropy");
      for (auto &b : key) b = 0;
    }
  }
  return ctx;
}

uint64_t bid_entropy(uint64_t salt) {
  if (config().seeded) {
    std::mt19937_64 gen(config().seed * 0x9E3779B97F4A7C15ull + salt);
    return gen();
  }
  std::random_device rd;  // SEAL/bidder.cpp:27
  return ((uint64_t)rd() << 32) | rd();
}

void check(int rc, const char *what) {
  if (rc != PA_OK) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, pa_last_error(engine()));
    exit(1);
  }
}

std::vector<Scalar> draw(uint64_t stream, uint64_t *counter, size_t k) {
  std::vector<Scalar> out(k);
  if (k) check(pa_rng_fill(engine(), config().seed, &stream, counter, k, out[0].b, 1), "pa_rng_fill");
  return out;
}

}  // namespace pa_host
