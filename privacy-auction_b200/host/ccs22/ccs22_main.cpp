// ./CCS22 <#bidders> <bit length of bids> [options] — the reference's command line
// (CCS22/main.cpp:16-199): same phases, same summary block, exit code 0 when every party
// computed the true maximum bid.  Options: --seed S, --bids a,b,.., --evaluator E,
// --transcript F (PACCS22T), --device D, --quiet.
#include "parties.h"

#include "../engine.h"
#include "../print.h"
#include "../trackers.h"

#include <algorithm>
#include <bitset>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

using namespace ccs22;

static std::vector<uint8_t> T;
static void put(const void *p, size_t n) { T.insert(T.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
static void put_u64(uint64_t v) {
  for (int i = 0; i < 8; ++i) T.push_back((uint8_t)(v >> (8 * i)));
}

int main(int argc, char *argv[]) {
  std::vector<std::string> pos_args;
  std::string bidarg, transcript;
  long evArg = -1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--seed" && i + 1 < argc) pa_host::config().seed = std::stoull(argv[++i]), pa_host::config().seeded = true;
    else if (a == "--bids" && i + 1 < argc) bidarg = argv[++i];
    else if (a == "--evaluator" && i + 1 < argc) evArg = std::stol(argv[++i]);
    else if (a == "--transcript" && i + 1 < argc) transcript = argv[++i];
    else if (a == "--device" && i + 1 < argc) pa_host::config().device = std::stoi(argv[++i]);
    else if (a == "--quiet") pa_host::quiet() = true;
    else pos_args.push_back(a);
  }
  if (pos_args.size() != 2) {
    PRINT_ERROR("Usage: " << argv[0] << " <#bidders> <bit length of bids> [--seed S] [--bids a,b,..] [--evaluator E] [--transcript F]");
    exit(1);
  }
  size_t n = std::stoul(pos_args[0]);
  size_t c = std::stoul(pos_args[1]);
  bool flag = true;
  pa_host::engine();  // GPU context + comb table once, before any party timer

  std::vector<size_t> given;
  for (size_t p = 0; !bidarg.empty() && p <= bidarg.size();) {
    size_t e = bidarg.find(',', p);
    given.push_back(std::stoull(bidarg.substr(p, e == std::string::npos ? e : e - p)));
    if (e == std::string::npos) break;
    p = e + 1;
  }

  BulletinBoard bb(n, c);
  // =============== Initialization phase ============
  size_t evaluatorId = evArg >= 0 ? (size_t)evArg : (size_t)(pa_host::bid_entropy(0xE7A1ull) % n);
  auto pos = [evaluatorId](size_t i) {
    assert(i != evaluatorId);
    return i < evaluatorId ? i : i - 1;
  };
  PRINT_MESSAGE("#bidders: n = " << n << ", bit length of bids: c = " << c << "\nEvaluator: " << evaluatorId);

  bool haveBids = given.size() == n;
  Evaluator evaluator = haveBids ? Evaluator(evaluatorId, n, c, bb.getPubParams(), given[evaluatorId])
                                 : Evaluator(evaluatorId, n, c, bb.getPubParams());
  std::vector<Bidder> bidders;
  std::vector<size_t> bids(n);
  for (size_t i = 0; i < n; ++i) {
    if (i == evaluatorId) {
      bids[i] = evaluator.getBid();
    } else {
      bidders.push_back(haveBids ? Bidder(i, n, c, bb.getPubParams(), given[i]) : Bidder(i, n, c, bb.getPubParams()));
      bids[i] = bidders[pos(i)].getBid();
    }
  }
  auto maxBid = *std::max_element(bids.begin(), bids.end());
  PRINT_MESSAGE("Finished initialization.\nMax bid: " << maxBid << ", Max bid (in binary): "
                                                      << std::bitset<C_MAX>(maxBid).to_string().substr(C_MAX - c));
  put("PACCS22T", 8);
  put_u64(n), put_u64(c), put_u64(pa_host::config().seeded ? pa_host::config().seed : 0), put_u64(evaluatorId);
  for (size_t b : bids) put_u64(b);
  put(&bb.getPubParams().g1, 64);
  put(&bb.getPubParams().h, 64);

  // =============== Setup phase =====================
  for (size_t i = 0; i < n; ++i) {
    Bidder &p = i == evaluatorId ? (Bidder &)evaluator : bidders[pos(i)];
    p.setup();
    bb.addCommitmentMsg(i, p.getCommitments());
    bb.addPublicKeyMsg(i, p.getPubKeys());
    put(&p.getCommitments(), 64);
    put(p.getPubKeys().data(), 64 * c);
  }

  // =============== Computation phase ===============
  for (size_t step = 0; step < c; ++step) {
    for (size_t i = 0; i < n; ++i) {
      if (i == evaluatorId) evaluator.BESEncode(bb.getPublicKeysByStep(step), step);
      else bidders[pos(i)].BESEncode(bb.getPublicKeysByStep(step), step);
    }
    bb.addOTR1Vec(evaluator.OTReceive1(step));
    for (size_t j = 0; j + 1 < n; ++j) {
      OT_R1 r1 = bb.getOTR1(j);
      put(&r1, sizeof r1);
    }
    for (size_t i = 0; i < n; ++i)
      if (i != evaluatorId) bb.addOTS(pos(i), bidders[pos(i)].OTSend(step, bb.getOTR1(pos(i))));
    for (auto &s : bb.getOTSVec()) put(&s, sizeof s);
    bb.addd(evaluator.OTReceive2(step, bb.getOTSVec()));
    T.push_back((uint8_t)bb.getd());
    for (size_t i = 0; i < n; ++i)
      if (i != evaluatorId) bidders[pos(i)].checkIfEnterDeciderRound(step, bb.getd());
  }
  for (size_t i = 0; i < n; ++i) put_u64(i == evaluatorId ? evaluator.getMaxBid() : bidders[pos(i)].getMaxBid());

  // =======  Verification phase: TODO in the reference (CCS22/main.cpp:132-134), not invented here

  // ============== Print info =======================
  auto &tt = TimeTracker::getInstance();
  auto &dt = DataTracker::getInstance();
  double nb = n > 1 ? (double)(n - 1) : 1.0;
  PRINT_INFO("#bidders: n = " << n << ", bit length of bids: c = " << c << std::endl
             << "Time (one bidder): " << tt.getCategoryTimeInSeconds(BIDDER_CATEGORY) / nb << " s." << std::endl
             << "Time (one evaluator): " << tt.getCategoryTimeInSeconds(EVALUATOR_CATEGORY) << " s." << std::endl
             << "Data (one bidder): " << dt.getCategoryDataSizeInMB(BIDDER_CATEGORY) / nb + dt.getCategoryDataSizeInMB(BIDDER_AND_EVALUATOR_CATEGORY) / n << " MB" << std::endl
             << "Data (one evaluator): " << dt.getCategoryDataSizeInMB(EVALUATOR_CATEGORY) + dt.getCategoryDataSizeInMB(BIDDER_AND_EVALUATOR_CATEGORY) / n << " MB" << std::endl
             << "Data (total communication, #bidders=" << n - 1 << " ,#evaluators=" << 1 << "): " << dt.getTotalDataSizeInMB() << " MB");
  fprintf(stderr, "{\"impl\":\"b200\",\"protocol\":\"ccs22\",\"n\":%zu,\"c\":%zu,\"evaluator\":%zu,\"maxbid\":%zu,\"bytes\":%zu,\"t_bidder_s\":%.4f,\"t_evaluator_s\":%.4f}\n",
          n, c, evaluatorId, (size_t)maxBid, T.size(), tt.getCategoryTimeInSeconds(BIDDER_CATEGORY), tt.getCategoryTimeInSeconds(EVALUATOR_CATEGORY));
  if (!transcript.empty()) {
    FILE *f = fopen(transcript.c_str(), "wb");
    if (!f) {
      perror("fopen");
      exit(1);
    }
    fwrite(T.data(), 1, T.size(), f);
    fclose(f);
  }

  // ============== Test Correctness =================
  for (size_t i = 0; i < n; ++i) {
    size_t m = i == evaluatorId ? evaluator.getMaxBid() : bidders[pos(i)].getMaxBid();
    if (m != maxBid) {
      flag = false;
      PRINT_ERROR((i == evaluatorId ? "Evaluator " : "Bidder ") << i << " failed to calculate max bid.\n"
                                                                << std::bitset<C_MAX>(m).to_string().substr(C_MAX - c));
    }
  }
  if (!flag) exit(1);
  PRINT_MESSAGE("Finished auction, all bidder calculated max bid.\nMax bid: "
                << maxBid << ", Max bid (in binary): " << std::bitset<C_MAX>(maxBid).to_string().substr(C_MAX - c));
  return 0;
}
