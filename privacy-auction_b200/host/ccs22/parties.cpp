#include "parties.h"

#include "../engine.h"
#include "../print.h"
#include "../trackers.h"

#include <bitset>
#include <cassert>
#include <random>

using namespace pa_host;

namespace ccs22 {
namespace {
template <class T> const uint8_t *bytes(const std::vector<T> &v) { return v.empty() ? nullptr : reinterpret_cast<const uint8_t *>(v.data()); }
template <class T> uint8_t *bytes(std::vector<T> &v) { return v.empty() ? nullptr : reinterpret_cast<uint8_t *>(v.data()); }

const uint64_t BB_STREAM = 0xFFFFFFFFull;

void trackPoint(const std::string &cat, const Point &p) { DataTracker::getInstance().addData(cat, p.isInfinity() ? 1 : 65); }

Scalar scalarOf(uint64_t v) {
  Scalar s;
  memset(s.b, 0, 32);
  for (int i = 0; i < 8; ++i) s.b[31 - i] = (uint8_t)(v >> (8 * i));
  return s;
}
}  // namespace

// ---------------------------------------------------------------- BulletinBoard
// Reference: CCS22/bulletinBoard.cpp:28-51 — g1 = g^rand256, h = g^rand256 (two fixed-base mults)
BulletinBoard::BulletinBoard(size_t n, size_t c) : n_(n), c_(c), commitments_(n) {
  assert(c <= C_MAX);
  uint64_t ctr = 0, stream = BB_STREAM;
  Scalar k[2];
  Point gh[2];
  check(pa_rng_fill256(engine(), config().seed, &stream, &ctr, 2, k[0].b, 1), "pa_rng_fill256");
  check(pa_fixed_base_mul(engine(), k[0].b, gh[0].b, 2), "pa_fixed_base_mul");
  Scalar one = scalarOf(1);
  check(pa_fixed_base_mul(engine(), one.b, pubParams_.g.b, 1), "pa_fixed_base_mul");
  pubParams_.g1 = gh[0];
  pubParams_.h = gh[1];
  pubKeys_.assign(n, std::vector<Point>(c));
  ot_s_vec_.resize(n ? n - 1 : 0);
}

const PubParams &BulletinBoard::getPubParams() const {
#ifdef ENABLE_COMMUNICATION_TRACKING
  // group: p (32) + a (0, BN_num_bytes(0) = 0) + b (1) + 3 ints; CCS22/bulletinBoard.cpp:180-207
  DataTracker::getInstance().addData(BIDDER_AND_EVALUATOR_CATEGORY, 32 + 0 + 1 + 3 * sizeof(int));
  trackPoint(BIDDER_AND_EVALUATOR_CATEGORY, pubParams_.g);
  trackPoint(BIDDER_AND_EVALUATOR_CATEGORY, pubParams_.g1);
  trackPoint(BIDDER_AND_EVALUATOR_CATEGORY, pubParams_.h);
  DataTracker::getInstance().addData(BIDDER_AND_EVALUATOR_CATEGORY, 32);  // order
#endif
  return pubParams_;
}
void BulletinBoard::addCommitmentMsg(size_t id, const Point &com) {
  assert(id < n_);
#ifdef ENABLE_COMMUNICATION_TRACKING
  DataTracker::getInstance().addData(BIDDER_AND_EVALUATOR_CATEGORY, sizeof(id));
  trackPoint(BIDDER_AND_EVALUATOR_CATEGORY, com);
#endif
  commitments_[id] = com;
}
void BulletinBoard::addPublicKeyMsg(size_t id, const std::vector<Point> &pubKeys) {
  assert(id < n_ && pubKeys.size() == c_);
#ifdef ENABLE_COMMUNICATION_TRACKING
  DataTracker::getInstance().addData(BIDDER_AND_EVALUATOR_CATEGORY, sizeof(id));
  for (auto &pk : pubKeys) trackPoint(BIDDER_AND_EVALUATOR_CATEGORY, pk);
#endif
  pubKeys_[id] = pubKeys;
}
const std::vector<Point> BulletinBoard::getPublicKeysByStep(size_t step) const {
  assert(step < c_);
  std::vector<Point> pk(n_);
  for (size_t i = 0; i < n_; ++i) {
    pk[i] = pubKeys_[i][step];
#ifdef ENABLE_COMMUNICATION_TRACKING
    trackPoint(BIDDER_AND_EVALUATOR_CATEGORY, pk[i]);
#endif
  }
  return pk;
}
void BulletinBoard::addOTR1Vec(const OT_R1_VEC &v) {
  ot_r1_vec_ = v;
#ifdef ENABLE_COMMUNICATION_TRACKING
  for (auto &r : v) trackPoint(EVALUATOR_CATEGORY, r.T2), trackPoint(EVALUATOR_CATEGORY, r.G), trackPoint(EVALUATOR_CATEGORY, r.H);
#endif
}
OT_R1 BulletinBoard::getOTR1(size_t j) const {
#ifdef ENABLE_COMMUNICATION_TRACKING
  trackPoint(BIDDER_CATEGORY, ot_r1_vec_[j].T2), trackPoint(BIDDER_CATEGORY, ot_r1_vec_[j].G), trackPoint(BIDDER_CATEGORY, ot_r1_vec_[j].H);
#endif
  return ot_r1_vec_[j];
}
void BulletinBoard::addOTS(size_t j, const OT_S &s) {
#ifdef ENABLE_COMMUNICATION_TRACKING
  trackPoint(BIDDER_CATEGORY, s.C0), trackPoint(BIDDER_CATEGORY, s.C1), trackPoint(BIDDER_CATEGORY, s.z);
#endif
  ot_s_vec_[j] = s;
}
OT_S_VEC BulletinBoard::getOTSVec() const {
#ifdef ENABLE_COMMUNICATION_TRACKING
  for (auto &s : ot_s_vec_) trackPoint(EVALUATOR_CATEGORY, s.C0), trackPoint(EVALUATOR_CATEGORY, s.C1), trackPoint(EVALUATOR_CATEGORY, s.z);
#endif
  return ot_s_vec_;
}
void BulletinBoard::addd(size_t d) {
#ifdef ENABLE_COMMUNICATION_TRACKING
  DataTracker::getInstance().addData(EVALUATOR_CATEGORY, sizeof(d));
#endif
  d_ = d;
}
size_t BulletinBoard::getd() const {
#ifdef ENABLE_COMMUNICATION_TRACKING
  DataTracker::getInstance().addData(BIDDER_CATEGORY, sizeof(d_));
#endif
  return d_;
}

// ---------------------------------------------------------------- Bidder
Bidder::Bidder(size_t id, size_t n, size_t c, const PubParams &p) : id_(id), c_(c), n_(n), pp(p) {
  init((size_t)(bid_entropy(id + 1) & ((1ull << c) - 1)));
}
Bidder::Bidder(size_t id, size_t n, size_t c, const PubParams &p, size_t bid) : id_(id), c_(c), n_(n), pp(p) { init(bid); }

void Bidder::init(size_t bid) {
  assert(c_ <= C_MAX);
  bid_ = bid;
  drawCounter = 0;
  inRaceFlag = true;
  d = 0;
  maxBid = 0;
  privKeys.resize(c_);
  pubKeys.resize(c_);
  randomS.resize(c_);
  randomT.resize(c_);
  memset(&B, 0, sizeof B);
  memset(&Com, 0, sizeof Com);
  R = draw(1)[0];  // CCS22/bidder.cpp:21-27
  binaryBidStr = std::bitset<C_MAX>(bid_).to_string().substr(C_MAX - c_);
  PRINT_MESSAGE("Construct Bidder: " << id_ << "\nBid: " << bid_ << ", Bid (in binary): " << binaryBidStr);
}

size_t Bidder::getId() { return id_; }
size_t Bidder::getBid() { return bid_; }
size_t Bidder::getMaxBid() { return maxBid; }
const Point &Bidder::getCommitments() const { return Com; }
const std::vector<Point> &Bidder::getPubKeys() const { return pubKeys; }

std::vector<Scalar> Bidder::draw(size_t k) {
  uint64_t ctr = drawCounter;
  auto v = pa_host::draw(id_, &ctr, k);
  drawCounter = ctr;
  return v;
}
std::vector<Scalar> Bidder::draw256(size_t k) {
  std::vector<Scalar> out(k);
  uint64_t stream = id_, ctr = drawCounter;
  if (k) check(pa_rng_fill256(engine(), config().seed, &stream, &ctr, k, out[0].b, 1), "pa_rng_fill256");
  drawCounter = ctr;
  return out;
}

// H = SHA256inSetup(hashed...) and Com = g^bid * g1^H + h^R          CCS22/bidder.cpp:80-88
void Bidder::commit(const std::vector<Scalar> &hashed) {
  Scalar bid = scalarOf(bid_);
  Point params[2] = {pp.g1, pp.h};
  check(pa_ccs22_commit(engine(), bytes(hashed), hashed.size(), bid.b, R.b, params[0].b, H.b, Com.b, 1), "pa_ccs22_commit");
}

// Reference: CCS22/bidder.cpp:48-89 — per bit x, r, s, t and X = g^x
void Bidder::setupInner() {
  std::vector<Scalar> dr = draw(4 * c_), xs(c_), hashed(4 * c_);
  for (size_t i = 0; i < c_; ++i) {
    privKeys[i] = PrivKey{dr[4 * i], dr[4 * i + 1]};
    randomS[i] = dr[4 * i + 2];
    randomT[i] = dr[4 * i + 3];
    xs[i] = dr[4 * i];
    hashed[i] = dr[4 * i], hashed[c_ + i] = dr[4 * i + 1], hashed[2 * c_ + i] = dr[4 * i + 2], hashed[3 * c_ + i] = dr[4 * i + 3];
  }
  check(pa_fixed_base_mul(engine(), bytes(xs), bytes(pubKeys), c_), "pa_fixed_base_mul");
  commit(hashed);
}
void Bidder::setup() {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  setupInner();
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
}

// Reference: CCS22/bidder.cpp:118-147 — own Y (n-2 additions), B = Y^x or g^r
void Bidder::BESEncodeInner(const std::vector<Point> &pk, size_t step) {
  int bit = binaryBidStr[step] - '0';
  d = (inRaceFlag && bit == 1) ? 1 : 0;
  uint64_t id = id_;
  uint8_t veto = (uint8_t)d;
  check(pa_ccs22_bes_encode(engine(), bytes(pk), n_, &id, &veto, privKeys[step].x.b, privKeys[step].r.b, B.b, 1), "pa_ccs22_bes_encode");
}
void Bidder::BESEncode(const std::vector<Point> &pk, size_t step) {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  BESEncodeInner(pk, step);
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
}

// Reference: CCS22/bidder.cpp:155-198 — six scalar multiplications one after the other there; here one
// engine call in which every multiplication of the message runs in its own warp.
OT_S Bidder::OTSend(size_t step, const OT_R1 &r1) {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  OT_S out;
  Scalar m = draw256(1)[0];  // M1 = g^rand
  Point params[2] = {pp.g1, pp.h};
  Scalar st[2] = {randomS[step], randomT[step]};
  check(pa_ccs22_ot_send(engine(), r1.T2.b, params[0].b, B.b, st[0].b, m.b, out.z.b, 1), "pa_ccs22_ot_send");
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
  return out;
}

// Reference: CCS22/bidder.cpp:200-212
void Bidder::checkIfEnterDeciderRound(size_t step, size_t new_d) {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  assert(new_d == 0 || new_d == 1);
  if (new_d == 1) {
    inRaceFlag = d == 0 ? false : inRaceFlag;
    maxBid |= ((size_t)1 << (c_ - step - 1));
  }
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
}

// ---------------------------------------------------------------- Evaluator
Evaluator::Evaluator(size_t id, size_t n, size_t c, const PubParams &p) : Bidder(id, n, c, p) { randomBeta.assign(c, std::vector<Scalar>(n ? n - 1 : 0)); }
Evaluator::Evaluator(size_t id, size_t n, size_t c, const PubParams &p, size_t bid) : Bidder(id, n, c, p, bid) {
  randomBeta.assign(c, std::vector<Scalar>(n ? n - 1 : 0));
}

// Reference: CCS22/evaluator.cpp:22-63 — per bit x, r and n-1 betas
void Evaluator::setupInner() {
  size_t nb = n_ - 1;
  std::vector<Scalar> dr = draw((2 + nb) * c_), xs(c_), hashed((n_ + 1) * c_);
  for (size_t i = 0; i < c_; ++i) {
    const Scalar *row = &dr[(2 + nb) * i];
    privKeys[i] = PrivKey{row[0], row[1]};
    xs[i] = row[0];
    hashed[i] = row[0], hashed[c_ + i] = row[1];
    for (size_t j = 0; j < nb; ++j) randomBeta[i][j] = row[2 + j], hashed[2 * c_ + i * nb + j] = row[2 + j];
  }
  check(pa_fixed_base_mul(engine(), bytes(xs), bytes(pubKeys), c_), "pa_fixed_base_mul");
  commit(hashed);
}
void Evaluator::setup() {
  TimeTracker::getInstance().start(EVALUATOR_CATEGORY);
  setupInner();
  TimeTracker::getInstance().stop(EVALUATOR_CATEGORY);
}
void Evaluator::BESEncode(const std::vector<Point> &pk, size_t step) {
  TimeTracker::getInstance().start(EVALUATOR_CATEGORY);
  BESEncodeInner(pk, step);
  TimeTracker::getInstance().stop(EVALUATOR_CATEGORY);
}

// Reference: CCS22/evaluator.cpp:78-115 — 4 scalar mults per other bidder; one batched, fused call here
OT_R1_VEC Evaluator::OTReceive1(size_t step) {
  TimeTracker::getInstance().start(EVALUATOR_CATEGORY);
  size_t nb = n_ - 1;
  OT_R1_VEC out(nb);
  if (nb) {
    std::vector<Scalar> k = draw256(nb), alpha(nb, scalarOf(d));
    std::vector<Point> params(2 * nb);
    for (size_t j = 0; j < nb; ++j) params[2 * j] = pp.g1, params[2 * j + 1] = pp.h;
    check(pa_ccs22_ot_recv1(engine(), bytes(k), bytes(randomBeta[step]), bytes(alpha), bytes(params), (uint8_t *)out.data(), nb),
          "pa_ccs22_ot_recv1");
  }
  TimeTracker::getInstance().stop(EVALUATOR_CATEGORY);
  return out;
}

// Reference: CCS22/evaluator.cpp:117-156
size_t Evaluator::OTReceive2(size_t step, const OT_S_VEC &ots) {
  TimeTracker::getInstance().start(EVALUATOR_CATEGORY);
  size_t ret;
  if (d == 1) {
    maxBid |= ((size_t)1 << (c_ - step - 1));
    ret = 1;
  } else {
    size_t nb = n_ - 1;
    assert(ots.size() == nb);
    int isInf = 1;
    check(pa_ccs22_ot_recv2(engine(), (const uint8_t *)ots.data(), bytes(randomBeta[step]), B.b, nb, &isInf), "pa_ccs22_ot_recv2");
    if (!isInf) {
      inRaceFlag = false;
      maxBid |= ((size_t)1 << (c_ - step - 1));
      ret = 1;
    } else {
      ret = 0;
    }
  }
  TimeTracker::getInstance().stop(EVALUATOR_CATEGORY);
  return ret;
}

}  // namespace ccs22
