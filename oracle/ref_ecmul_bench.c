/*
 * TEST / BASELINE INFRASTRUCTURE — times the reference's CPU implementation of
 * the scalar-multiplication hot path: OpenSSL libcrypto's EC_POINT_mul on
 * secp256k1 in exactly the two call shapes the reference uses,
 *     EC_POINT_mul(group, r, k, NULL, NULL, ctx)   fixed base     SEAL/bidder.cpp:98
 *     EC_POINT_mul(group, r, NULL, P, k, ctx)      variable base  SEAL/bidder.cpp:129
 * on T independent threads (the reference itself is single-threaded; one
 * libcrypto context per thread, no sharing).  Used by bench.py for the
 * `cpu_baseline` object and the `--impl reference` arm.  The reference has no
 * source file of its own for this (its hot path IS these library calls), so
 * this driver is the smallest program that executes them.
 *
 * usage: ecmul_ref <threads> <pairs-per-thread> [seed]
 *   each thread performs <pairs> fixed-base and <pairs> variable-base mults
 * prints one JSON line: {"threads":T,"mults":M,"seconds":S,"mults_per_s":R,"xor":"..."}
 */
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/obj_mac.h>
#include <openssl/sha.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  int tid;
  long pairs;
  uint64_t seed;
  unsigned char acc[32];
  pthread_barrier_t *bar;
} job;

static void draw(unsigned char out[32], uint64_t seed, uint64_t stream, uint64_t ctr) {
  unsigned char msg[28];
  memcpy(msg, "PAv1", 4);
  for (int i = 0; i < 8; ++i) {
    msg[4 + i] = (unsigned char)(seed >> (8 * i));
    msg[12 + i] = (unsigned char)(stream >> (8 * i));
    msg[20 + i] = (unsigned char)(ctr >> (8 * i));
  }
  SHA256(msg, sizeof msg, out);
}

static void *worker(void *arg) {
  job *j = (job *)arg;
  EC_GROUP *group = EC_GROUP_new_by_curve_name(NID_secp256k1);
  BN_CTX *ctx = BN_CTX_new();
  EC_POINT *P = EC_POINT_new(group), *R = EC_POINT_new(group);
  BIGNUM *k = BN_new(), *x = BN_new();
  unsigned char d[32], xb[32];
  memset(j->acc, 0, 32);
  pthread_barrier_wait(j->bar);
  for (long i = 0; i < j->pairs; ++i) {
    draw(d, j->seed, (uint64_t)j->tid, (uint64_t)(2 * i));
    BN_bin2bn(d, 32, k);
    EC_POINT_mul(group, P, k, NULL, NULL, ctx); /* fixed base */
    draw(d, j->seed, (uint64_t)j->tid, (uint64_t)(2 * i + 1));
    BN_bin2bn(d, 32, k);
    EC_POINT_mul(group, R, NULL, P, k, ctx); /* variable base */
    if ((i & 1023) == 0) { /* keep the results observable */
      EC_POINT_get_affine_coordinates(group, R, x, NULL, ctx);
      BN_bn2binpad(x, xb, 32);
      for (int b = 0; b < 32; ++b) j->acc[b] ^= xb[b];
    }
  }
  pthread_barrier_wait(j->bar);
  BN_free(k); BN_free(x);
  EC_POINT_free(P); EC_POINT_free(R);
  BN_CTX_free(ctx);
  EC_GROUP_free(group);
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <threads> <pairs-per-thread> [seed]\n", argv[0]);
    return 2;
  }
  int T = atoi(argv[1]);
  long pairs = atol(argv[2]);
  uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;
  if (T < 1) T = 1;
  pthread_t *th = calloc(T, sizeof *th);
  job *jobs = calloc(T, sizeof *jobs);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, T + 1);
  for (int t = 0; t < T; ++t) {
    jobs[t].tid = t; jobs[t].pairs = pairs; jobs[t].seed = seed; jobs[t].bar = &bar;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  struct timespec t0, t1;
  pthread_barrier_wait(&bar);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&bar);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  unsigned char acc[32] = {0};
  for (int t = 0; t < T; ++t) {
    pthread_join(th[t], NULL);
    for (int b = 0; b < 32; ++b) acc[b] ^= jobs[t].acc[b];
  }
  double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  double mults = 2.0 * pairs * T;
  char hex[65];
  for (int b = 0; b < 32; ++b) sprintf(hex + 2 * b, "%02x", acc[b]);
  printf("{\"threads\":%d,\"mults\":%.0f,\"seconds\":%.4f,\"mults_per_s\":%.1f,\"xor\":\"%s\"}\n", T, mults, s, mults / s, hex);
  return 0;
}
