/*
 * TEST INFRASTRUCTURE ("Tier-A oracle", CCS22) — drives the UNMODIFIED reference CCS22
 * classes (compiled from /root/reference/CCS22/{bidder,evaluator,hash,bulletinBoard}.cpp with
 * oracle/shim/pa_seed_shim.h force-included) in the order of the reference's own main
 * (CCS22/main.cpp:16-130) and serialises everything published into the PACCS22T transcript.
 *
 * usage: ccs22_ref <n> <c> <seed> <evaluatorId> <bids: b0,b1,...|-> <out-file|->
 * Randomness: party i draws from PA stream (seed, i); the bulletin board (g1, h) from
 * stream (seed, 0xFFFFFFFF).  Selected before every call into a party, see pa_seed_shim.h.
 *
 * Transcript: "PACCS22T" | n c seed evaluatorId (LE64) | n bids (LE64, by id) | g1 h |
 *   per party (id order): Com, c public keys | per step: (n-1) x (T2, G, H), (n-1) x (z, C0, C1),
 *   1 byte d | n x LE64 max bid (id order).  Points 64 B (X||Y, zeros = infinity).
 */
#include "bidder.h"
#include "bulletinBoard.h"
#include "evaluator.h"

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
static void put_point(const EC_POINT *P) {
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
  if (argc < 7) {
    fprintf(stderr, "usage: %s <n> <c> <seed> <evaluatorId> <bids|-> <out|->\n", argv[0]);
    return 2;
  }
  size_t n = std::stoul(argv[1]), c = std::stoul(argv[2]);
  uint64_t seed = std::stoull(argv[3]);
  size_t evaluatorId = std::stoul(argv[4]);
  std::string bidarg = argv[5], outarg = argv[6];
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
  G = EC_GROUP_new_by_curve_name(CURVE);
  CTX = BN_CTX_new();
  std::cout.setstate(std::ios_base::failbit);

  pa_shim_select(seed, 0xFFFFFFFFull);
  BulletinBoard bb(n, c); /* draws g1 = g^rand256, h = g^rand256, CCS22/bulletinBoard.cpp:28-51 */
  auto pos = [evaluatorId](size_t i) { return i < evaluatorId ? i : i - 1; };

  OUT.insert(OUT.end(), {'P', 'A', 'C', 'C', 'S', '2', '2', 'T'});
  put_u64(n); put_u64(c); put_u64(seed); put_u64(evaluatorId);
  for (size_t j = 0; j < n; ++j) put_u64(bids[j]);
  const PubParams &pp = bb.getPubParams();
  put_point(pp.g1);
  put_point(pp.h);

  pa_shim_select(seed, evaluatorId);
  pa_shim_inject_bid(bids[evaluatorId]);
  Evaluator evaluator(evaluatorId, n, c, bb.getPubParams());
  std::vector<Bidder> bidders;
  bidders.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    if (i == evaluatorId) continue;
    pa_shim_select(seed, i);
    pa_shim_inject_bid(bids[i]);
    bidders.push_back(Bidder(i, n, c, bb.getPubParams()));
  }
  /* setup phase, CCS22/main.cpp:70-80 */
  for (size_t i = 0; i < n; ++i) {
    pa_shim_select(seed, i);
    if (i == evaluatorId) {
      evaluator.setup();
      bb.addCommitmentMsg(i, evaluator.getCommitments());
      bb.addPublicKeyMsg(i, evaluator.getPubKeys());
      put_point(evaluator.getCommitments());
      for (auto pk : evaluator.getPubKeys()) put_point(pk);
    } else {
      Bidder &b = bidders[pos(i)];
      b.setup();
      bb.addCommitmentMsg(i, b.getCommitments());
      bb.addPublicKeyMsg(i, b.getPubKeys());
      put_point(b.getCommitments());
      for (auto pk : b.getPubKeys()) put_point(pk);
    }
  }
  /* computation phase, CCS22/main.cpp:87-130 */
  for (size_t step = 0; step < c; ++step) {
    for (size_t i = 0; i < n; ++i) {
      pa_shim_select(seed, i);
      if (i == evaluatorId) evaluator.BESEncode(bb.getPublicKeysByStep(step), step);
      else bidders[pos(i)].BESEncode(bb.getPublicKeysByStep(step), step);
    }
    pa_shim_select(seed, evaluatorId);
    bb.addOTR1Vec(evaluator.OTReceive1(step));
    for (size_t j = 0; j + 1 < n; ++j) {
      OT_R1 r1 = bb.getOTR1(j);
      put_point(r1.T2); put_point(r1.G); put_point(r1.H);
    }
    for (size_t i = 0; i < n; ++i) {
      if (i == evaluatorId) continue;
      pa_shim_select(seed, i);
      bb.addOTS(pos(i), bidders[pos(i)].OTSend(step, bb.getOTR1(pos(i))));
    }
    OT_S_VEC sv = bb.getOTSVec();
    for (auto s : sv) { put_point(s->z); put_point(s->C0); put_point(s->C1); }
    pa_shim_select(seed, evaluatorId);
    bb.addd(evaluator.OTReceive2(step, bb.getOTSVec()));
    OUT.push_back((unsigned char)bb.getd());
    for (size_t i = 0; i < n; ++i)
      if (i != evaluatorId) bidders[pos(i)].checkIfEnterDeciderRound(step, bb.getd());
  }
  uint64_t truemax = *std::max_element(bids.begin(), bids.end());
  bool ok = true;
  for (size_t i = 0; i < n; ++i) {
    uint64_t m = i == evaluatorId ? evaluator.getMaxBid() : bidders[pos(i)].getMaxBid();
    put_u64(m);
    ok &= m == truemax;
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
  fprintf(stderr, "{\"impl\":\"reference-tierA-ccs22\",\"n\":%zu,\"c\":%zu,\"seed\":%llu,\"evaluator\":%zu,\"ok\":%s,\"maxbid\":%llu,\"bytes\":%zu,\"sha256\":\"%s\",\"draws\":%llu}\n",
          n, c, (unsigned long long)seed, evaluatorId, ok ? "true" : "false", (unsigned long long)truemax, OUT.size(), hex,
          (unsigned long long)pa_shim_draws());
  return ok ? 0 : 1;
}
