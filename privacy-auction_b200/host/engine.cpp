#include "engine.h"

#include <cstdio>
#include <cstdlib>

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
  }
  return ctx;
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
