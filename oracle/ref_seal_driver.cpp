/*
 * TEST INFRASTRUCTURE ("Tier-A oracle") — drives the UNMODIFIED reference SEAL
 * classes (compiled from /root/reference/SEAL/{bidder,hash,bulletinBoard}.cpp
 * with oracle/shim/pa_seed_shim.h force-included) through their public API in
 * the order of the reference's own main (SEAL/main.cpp:32-120) and serialises
 * everything they publish into the PASEALT1 transcript format (DESIGN.md
 * "Transcript format").  Nothing in here is product code; the product never
 * links or executes it.
 *
 * usage: seal_ref <n> <c> <seed> <bids: b0,b1,...|-> <out-file|-> [all|none]
 *   bids "-"  : bid_j = LE64(SHA-256("PAbid"||LE64 seed||LE64 j)[0..8]) mod 2^min(c,31)
 *   all|none  : run every bidder's verify* as main.cpp does (default) or skip
 * Randomness: party j draws from PA stream (seed, j) — selected before every
 * call into Bidder j, see pa_seed_shim.h.
 */
#include "bidder.h"
#include "bulletinBoard.h"
#include "dataTracker.h"

#include <chrono>
#include <cstdio>
#include <openssl/sha.h>
#include <string>
#include <vector>

static const EC_GROUP *G = nullptr;
static BN_CTX *CTX = nullptr;
static std::vector<unsigned char> OUT;

static void put_u64(uint64_t v) {
  for (int i = 0; i < 8; ++i) OUT.push_back((unsigned char)(v >> (8 * i)));
}
static void put_u32(uint32_t v) {
  for (int i = 0; i < 4; ++i) OUT.push_back((unsigned char)(v >> (8 * i)));
}
static void put_point(const EC_POINT *P) { /* 64 B: X||Y big-endian, infinity = zeros */
  unsigned char buf[64];
  memset(buf, 0, sizeof buf);
  if (!EC_POINT_is_at_infinity(G, P)) {
    BIGNUM *x = BN_new(), *y = BN_new();
    EC_POINT_get_affine_coordinates(G, P, x, y, CTX);
    BN_bn2binpad(x, buf, 32);
    BN_bn2binpad(y, buf + 32, 32);
    BN_free(x);
    BN_free(y);
  }
  OUT.insert(OUT.end(), buf, buf + 64);
}
static void put_scalar(const BIGNUM *k) { /* 32 B big-endian */
  unsigned char buf[32];
  BN_bn2binpad(k, buf, 32);
  OUT.insert(OUT.end(), buf, buf + 32);
}
static void put_pok(const NIZKPoKDLog &p) {
  put_point(p.eps);
  put_scalar(p.rho);
}

static uint64_t derive_bid(uint64_t seed, uint64_t j, size_t c) {
  unsigned char msg[21], d[32];
  memcpy(msg, "PAbid", 5);
  for (int i = 0; i < 8; ++i) msg[5 + i] = (unsigned char)(seed >> (8 * i));
  for (int i = 0; i < 8; ++i) msg[13 + i] = (unsigned char)(j >> (8 * i));
  SHA256(msg, sizeof msg, d);
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v |= (uint64_t)d[i] << (8 * i);
  size_t bits = c < 31 ? c : 31;
  return v & ((1ull << bits) - 1);
}

int main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s <n> <c> <seed> <bids|-> <out|-> [all|none]\n", argv[0]);
    return 2;
  }
  size_t n = std::stoul(argv[1]), c = std::stoul(argv[2]);
  uint64_t seed = std::stoull(argv[3]);
  std::string bidarg = argv[4], outarg = argv[5];
  bool verify = !(argc > 6 && std::string(argv[6]) == "none");

  G = EC_GROUP_new_by_curve_name(CURVE);
  CTX = BN_CTX_new();

  std::vector<uint64_t> bids(n);
  if (bidarg == "-") {
    for (size_t j = 0; j < n; ++j) bids[j] = derive_bid(seed, j, c);
  } else {
    size_t pos = 0;
    for (size_t j = 0; j < n; ++j) {
      size_t e = bidarg.find(',', pos);
      bids[j] = std::stoull(bidarg.substr(pos, e == std::string::npos ? e : e - pos));
      pos = e == std::string::npos ? bidarg.size() : e + 1;
    }
  }

  /* silence the reference's PRINT_MESSAGE chatter on stdout */
  std::cout.setstate(std::ios_base::failbit);

  auto t0 = std::chrono::steady_clock::now();
  double t_prove = 0, t_verify = 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };

  OUT.insert(OUT.end(), {'P', 'A', 'S', 'E', 'A', 'L', 'T', '1'});
  put_u64(n);
  put_u64(c);
  put_u64(seed);

  std::vector<Bidder> bidders;
  bidders.reserve(n);
  BulletinBoard bb(n, c);
  for (size_t j = 0; j < n; ++j) {
    pa_shim_select(seed, j);
    pa_shim_inject_bid(bids[j]);
    bidders.push_back(Bidder(j, n, c));
    put_u64(bidders[j].getBid());
  }

  bool ok = true;
  /* commit phase, SEAL/main.cpp:46-48 */
  auto t = now();
  for (size_t j = 0; j < n; ++j) {
    pa_shim_select(seed, j);
    bb.addCommitmentMsg(bidders[j].commitBid(), j);
  }
  t_prove += secs(t, now());
  const std::vector<CommitmentPub> &all_cp = bb.getCommitments(); /* one read of the board for the dump */
  for (size_t j = 0; j < n; ++j) {
    const CommitmentPub &cp = all_cp[j];
    for (size_t i = 0; i < c; ++i) {
      put_point(cp[i].phi);
      put_point(cp[i].A);
      put_point(cp[i].B);
      put_pok(cp[i].pokdlogA);
      put_pok(cp[i].pokdlogB);
      put_point(cp[i].powfcom.eps11);
      put_point(cp[i].powfcom.eps12);
      put_point(cp[i].powfcom.eps21);
      put_point(cp[i].powfcom.eps22);
      put_scalar(cp[i].powfcom.rho1);
      put_scalar(cp[i].powfcom.rho2);
      put_scalar(cp[i].powfcom.ch2);
    }
  }
  t = now();
  for (size_t j = 0; j < n; ++j) {
    bool v = verify ? bidders[j].verifyCommitment(bb.getCommitments()) : true;
    ok &= v;
    OUT.push_back(v ? 1 : 0);
  }
  t_verify += secs(t, now());

  for (size_t step = 0; step < c; ++step) {
    /* round one, SEAL/main.cpp:73-75 */
    t = now();
    for (size_t j = 0; j < n; ++j) {
      pa_shim_select(seed, j);
      bb.addRoundOneMsg(bidders[j].roundOne(step), j);
    }
    t_prove += secs(t, now());
    const std::vector<RoundOnePub> &all_r1 = bb.getRoundOnePubs();
    for (size_t j = 0; j < n; ++j) {
      const RoundOnePub &p = all_r1[j];
      put_point(p.X);
      put_point(p.R);
      put_pok(p.pokdlogX);
      put_pok(p.pokdlogR);
    }
    t = now();
    for (size_t j = 0; j < n; ++j) {
      bool v = verify ? bidders[j].verifyRoundOne(bb.getRoundOnePubs()) : true;
      ok &= v;
      OUT.push_back(v ? 1 : 0);
    }
    t_verify += secs(t, now());

    /* round two, SEAL/main.cpp:93-95 */
    t = now();
    for (size_t j = 0; j < n; ++j) {
      pa_shim_select(seed, j);
      bb.addRoundTwoMsg(bidders[j].roundTwo(bb.getRoundOneXs(), step), j);
    }
    t_prove += secs(t, now());
    const std::vector<RoundTwoPub> &all_r2 = bb.getRoundTwoPubs();
    for (size_t j = 0; j < n; ++j) {
      const RoundTwoPub &p = all_r2[j];
      put_u32(p.stage == STAGE1 ? 1 : 2);
      put_point(p.b);
      if (p.stage == STAGE1) {
        const NIZKPoWFStage1 &q = p.powf.powfstage1;
        const EC_POINT *pts[] = {q.eps11, q.eps12, q.eps13, q.eps14, q.eps21, q.eps22, q.eps23, q.eps24};
        for (auto e : pts) put_point(e);
        const BIGNUM *scs[] = {q.rho11, q.rho12, q.rho21, q.rho22, q.ch2};
        for (auto s : scs) put_scalar(s);
      } else {
        const NIZKPoWFStage2 &q = p.powf.powfstage2;
        const EC_POINT *pts[] = {q.eps11, q.eps12, q.eps13, q.eps11prime, q.eps12prime, q.eps13prime,
                                 q.eps21, q.eps22, q.eps23, q.eps21prime, q.eps22prime, q.eps23prime,
                                 q.eps31, q.eps32, q.eps31prime, q.eps32prime};
        for (auto e : pts) put_point(e);
        const BIGNUM *scs[] = {q.rho11, q.rho12, q.rho13, q.rho21, q.rho22, q.rho23, q.rho31, q.rho32, q.ch2, q.ch3};
        for (auto s : scs) put_scalar(s);
      }
    }
    t = now();
    for (size_t j = 0; j < n; ++j) {
      bool v = verify ? bidders[j].verifyRoundTwo(bb.getRoundTwoPubs(), step) : true;
      ok &= v;
      OUT.push_back(v ? 1 : 0);
    }
    t_verify += secs(t, now());

    /* round three, SEAL/main.cpp:113-115 */
    t = now();
    for (size_t j = 0; j < n; ++j) OUT.push_back((unsigned char)bidders[j].roundThree(bb.getRoundTwoBs(), step));
    t_prove += secs(t, now());
  }
  uint64_t truemax = *std::max_element(bids.begin(), bids.end());
  for (size_t j = 0; j < n; ++j) {
    put_u64(bidders[j].getMaxBid());
    if (bidders[j].getMaxBid() != truemax) ok = false;
  }

  if (outarg != "-") {
    FILE *f = fopen(outarg.c_str(), "wb");
    if (!f) { perror("fopen"); return 2; }
    fwrite(OUT.data(), 1, OUT.size(), f);
    fclose(f);
  }
  unsigned char dg[32];
  SHA256(OUT.data(), OUT.size(), dg);
  char hex[65];
  for (int i = 0; i < 32; ++i) sprintf(hex + 2 * i, "%02x", dg[i]);
  fprintf(stderr,
          "{\"impl\":\"reference-tierA\",\"n\":%zu,\"c\":%zu,\"seed\":%llu,\"ok\":%s,\"maxbid\":%llu,"
          "\"bytes\":%zu,\"sha256\":\"%s\",\"draws\":%llu,\"t_prove_s\":%.3f,\"t_verify_s\":%.3f,\"t_total_s\":%.3f,"
          "\"data_bidder\":%zu,\"data_verifier\":%zu,\"data_total\":%zu}\n",
          n, c, (unsigned long long)seed, ok ? "true" : "false", (unsigned long long)truemax, OUT.size(), hex,
          (unsigned long long)pa_shim_draws(), t_prove, t_verify, secs(t0, now()),
          DataTracker::getInstance().getCategoryDataSize(BIDDER_CATEGORY),
          DataTracker::getInstance().getCategoryDataSize(VERIFIER_CATEGORY), DataTracker::getInstance().getTotalDataSize());
  return ok ? 0 : 1;
}
