#include "bidder.h"

#include "engine.h"
#include "trackers.h"

#include <bitset>
#include <cassert>
#include <random>

using namespace pa_host;

namespace {
template <class T> const uint8_t *bytes(const std::vector<T> &v) { return v.empty() ? nullptr : reinterpret_cast<const uint8_t *>(v.data()); }
template <class T> uint8_t *bytes(std::vector<T> &v) { return v.empty() ? nullptr : reinterpret_cast<uint8_t *>(v.data()); }
}  // namespace

Bidder::Bidder(size_t id, size_t n, size_t c) : id_(id), c_(c), n_(n) {
  assert(c <= C_MAX);
  // uniform c-bit bid; 64-bit arithmetic, so c = 32 does not degenerate to 0 as the
  // reference's `(1 << c) - 1` on int does
  init((size_t)(bid_entropy(id + 1) & ((c >= 64 ? ~0ull : (1ull << c)) - 1)));
}

Bidder::Bidder(size_t id, size_t n, size_t c, size_t bid) : id_(id), c_(c), n_(n) {
  assert(c <= C_MAX);
  init(bid);
}

void Bidder::init(size_t bid) {
  bid_ = bid;
  maxBid = 0;
  junctionFlag = false;
  prevDecidingStep = 0;
  prevDecidingBit = 1;
  drawCounter = 0;
  commitments.resize(c_);
  keys.resize(c_);
  binaryBidStr = std::bitset<C_MAX>(bid_).to_string().substr(C_MAX - c_);  // MSB first
  PRINT_MESSAGE("Construct Bidder: " << id_ << "\nBid: " << bid_ << ", Bid (in binary): " << binaryBidStr);
  curInfo.assign(n_, AuxilaryInfoPerBidder());
  prevDecidingInfo.assign(n_, AuxilaryInfoPerBidder());
  for (auto *info : {&curInfo, &prevDecidingInfo})
    for (auto &e : *info) memset(&e, 0, sizeof e);  // EC_POINT_new() is the point at infinity
}

size_t Bidder::getId() { return id_; }
size_t Bidder::getBid() { return bid_; }
size_t Bidder::getMaxBid() { return maxBid; }

std::vector<Scalar> Bidder::draw(size_t k) { return pa_host::draw(id_, &drawCounter, k); }

// In the Commit phase a bidder publishes, per bit of the bid, the commitment
// (phi, A, B), two Schnorr proofs and the OR proof of well-formedness.
// Reference: SEAL/bidder.cpp:1109-1162 — c iterations of ~10.5 EC_POINT_mul each;
// here: one draw call and four engine calls over the c bits.
CommitmentPub Bidder::commitBid() {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  CommitmentPub pubs(c_);
  // draw order per bit: alpha, beta | v_A | v_B | r1, (ch2, rho2 | ch1, rho1)   
  std::vector<Scalar> d = draw(7 * c_);
  std::vector<Scalar> alpha(c_), beta(c_), vA(c_), vB(c_), rnd(3 * c_);
  std::vector<uint8_t> bits(c_);
  std::vector<uint64_t> ids(c_, id_);
  for (size_t i = 0; i < c_; ++i) {
    alpha[i] = d[7 * i], beta[i] = d[7 * i + 1], vA[i] = d[7 * i + 2], vB[i] = d[7 * i + 3];
    rnd[3 * i] = d[7 * i + 4], rnd[3 * i + 1] = d[7 * i + 5], rnd[3 * i + 2] = d[7 * i + 6];
    bits[i] = (uint8_t)(binaryBidStr[i] - '0');
  }
  struct Triple { Point phi, A, B; };
  std::vector<Triple> pts(c_);
  std::vector<Point> A(c_), B(c_);
  std::vector<NIZKPoKDLog> pokA(c_), pokB(c_);
  std::vector<NIZKPoWFCom> com(c_);
  pa_ctx *e = engine();
  check(pa_commit_points(e, bytes(alpha), bytes(beta), bits.data(), bytes(pts), c_), "pa_commit_points");
  for (size_t i = 0; i < c_; ++i) A[i] = pts[i].A, B[i] = pts[i].B;
  check(pa_pokdlog_prove(e, bytes(A), bytes(alpha), ids.data(), bytes(vA), bytes(pokA), c_), "pa_pokdlog_prove");
  check(pa_pokdlog_prove(e, bytes(B), bytes(beta), ids.data(), bytes(vB), bytes(pokB), c_), "pa_pokdlog_prove");
  // the prover with witnesses: this bidder knows alpha AND beta, so every point of the OR proof is one fixed-base
  // multiplication (include/pa_engine.h); the proof bytes are those of pa_powfcom_prove
  std::vector<Scalar> ab(2 * c_);
  for (size_t i = 0; i < c_; ++i) ab[2 * i] = alpha[i], ab[2 * i + 1] = beta[i];
  check(pa_powfcom_prove_w(e, bytes(pts), bytes(ab), bits.data(), ids.data(), bytes(rnd), bytes(com), c_), "pa_powfcom_prove_w");
  for (size_t i = 0; i < c_; ++i) {
    commitments[i] = Commitment{pts[i].phi, pts[i].A, pts[i].B, alpha[i], beta[i]};
    pubs[i].phi = pts[i].phi, pubs[i].A = pts[i].A, pubs[i].B = pts[i].B;
    pubs[i].pokdlogA = pokA[i], pubs[i].pokdlogB = pokB[i], pubs[i].powfcom = com[i];
  }
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
  return pubs;
}

// Reference: SEAL/bidder.cpp:1171-1195 — for every other bidder and every bit,
// two Schnorr verifications and one OR-proof verification (12 EC_POINT_mul).
// Here: the (n-1)*c items are three batched engine calls.
bool Bidder::verifyCommitment(const std::vector<CommitmentPub> &pubs) {
  TimeTracker::getInstance().start(VERIFIER_CATEGORY);
  commitmentsBB = pubs;
  struct Triple { Point phi, A, B; };
  std::vector<Triple> stmt;
  std::vector<Point> A, B;
  std::vector<NIZKPoKDLog> pokA, pokB;
  std::vector<NIZKPoWFCom> com;
  std::vector<uint64_t> ids;
  for (size_t i = 0; i < pubs.size(); ++i) {
    if (i == id_) continue;
    for (size_t j = 0; j < c_; ++j) {
      const CommitmentPerBit &p = pubs[i][j];
      stmt.push_back(Triple{p.phi, p.A, p.B});
      A.push_back(p.A), B.push_back(p.B);
      pokA.push_back(p.pokdlogA), pokB.push_back(p.pokdlogB), com.push_back(p.powfcom);
      ids.push_back(i);
    }
  }
  size_t m = ids.size();
  std::vector<uint8_t> v1(m), v2(m), v3(m);
  pa_ctx *e = engine();
  check(pa_pokdlog_verify(e, bytes(pokA), bytes(A), ids.data(), v1.data(), m), "pa_pokdlog_verify");
  check(pa_pokdlog_verify(e, bytes(pokB), bytes(B), ids.data(), v2.data(), m), "pa_pokdlog_verify");
  check(pa_powfcom_verify(e, bytes(com), bytes(stmt), ids.data(), v3.data(), m), "pa_powfcom_verify");
  bool ret = true;
  for (size_t k = 0; k < m; ++k) {
    if (!v1[k] || !v2[k]) PRINT_ERROR("NIZKPoKDLog verification failed for bidder " << ids[k]);
    if (!v3[k]) PRINT_ERROR("NIZKPoWFCom verification failed for bidder " << ids[k]);
    ret &= v1[k] && v2[k] && v3[k];
  }
  TimeTracker::getInstance().stop(VERIFIER_CATEGORY);
  return ret;
}

// Reference: SEAL/bidder.cpp:1203-1236 — x, r, X = g^x, R = g^r and two Schnorr proofs.
RoundOnePub Bidder::roundOne(size_t step) {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  RoundOnePub pub;
  std::vector<Scalar> d = draw(4);  // x, r, v_X, v_R
  Scalar sc[2] = {d[0], d[1]}, v[2] = {d[2], d[3]};
  Point XR[2];
  NIZKPoKDLog pok[2];
  uint64_t ids[2] = {id_, id_};
  pa_ctx *e = engine();
  check(pa_fixed_base_mul(e, sc[0].b, XR[0].b, 2), "pa_fixed_base_mul");
  check(pa_pokdlog_prove(e, XR[0].b, sc[0].b, ids, v[0].b, (uint8_t *)pok, 2), "pa_pokdlog_prove");
  keys[step] = Key{XR[0], XR[1], d[0], d[1]};
  pub.X = XR[0], pub.R = XR[1], pub.pokdlogX = pok[0], pub.pokdlogR = pok[1];
  curInfo[id_].X = XR[0];
  curInfo[id_].R = XR[1];
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
  return pub;
}

// Reference: SEAL/bidder.cpp:1245-1262
bool Bidder::verifyRoundOne(const std::vector<RoundOnePub> &pubs) {
  TimeTracker::getInstance().start(VERIFIER_CATEGORY);
  std::vector<Point> P;
  std::vector<NIZKPoKDLog> pok;
  std::vector<uint64_t> ids;
  for (size_t i = 0; i < pubs.size(); ++i) {
    if (i == id_) continue;
    P.push_back(pubs[i].X), pok.push_back(pubs[i].pokdlogX), ids.push_back(i);
    P.push_back(pubs[i].R), pok.push_back(pubs[i].pokdlogR), ids.push_back(i);
    curInfo[i].X = pubs[i].X;
    curInfo[i].R = pubs[i].R;
  }
  std::vector<uint8_t> v(ids.size());
  check(pa_pokdlog_verify(engine(), bytes(pok), bytes(P), ids.data(), v.data(), ids.size()), "pa_pokdlog_verify");
  bool ret = true;
  for (size_t k = 0; k < v.size(); ++k) {
    if (!v[k]) PRINT_ERROR("NIZKPoKDLog verification failed for bidder " << ids[k]);
    ret &= v[k] != 0;
  }
  TimeTracker::getInstance().stop(VERIFIER_CATEGORY);
  return ret;
}

// Reference: SEAL/bidder.cpp:1271-1336.  The n(n-1) point additions that rebuild
// every Y are one scan on the GPU; the cryptogram b and its OR proof follow.
RoundTwoPub Bidder::roundTwo(const std::vector<Point> &Xs, size_t step) {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  RoundTwoPub pub;
  memset(&pub, 0, sizeof pub);
  assert(Xs.size() == n_);
  pa_ctx *e = engine();
  int bit = binaryBidStr[step] - '0';

  std::vector<Point> Ys(n_);
  check(pa_y_scan(e, bytes(Xs), bytes(Ys), n_), "pa_y_scan");
  for (size_t i = 0; i < n_; ++i) curInfo[i].Y = Ys[i];

  Point b;
  if ((!junctionFlag && bit == 0) || (junctionFlag && (bit == 0 || prevDecidingBit == 0))) {
    check(pa_var_base_mul(e, curInfo[id_].Y.b, keys[step].x.b, b.b, 1), "pa_var_base_mul");  // b = Y^x
    bit = 0;
  } else {
    check(pa_var_base_mul(e, keys[step].R.b, keys[step].x.b, b.b, 1), "pa_var_base_mul");  // b = R^x
    bit = 1;
  }
  pub.b = b;
  curInfo[id_].b = b;
  uint64_t id = id_;
  uint8_t bi = (uint8_t)bit;
  const Commitment &cm = commitments[step];
  if (!junctionFlag) {
    pub.stage = STAGE1;
    Point stmt[7] = {b, keys[step].X, curInfo[id_].Y, keys[step].R, cm.phi, cm.A, cm.B};
    Scalar sec[4] = {keys[step].x, cm.alpha, keys[step].r, cm.beta};  // extended secrets: prover with witnesses
    std::vector<Scalar> rnd = draw(5);
    check(pa_stage1_prove_w(e, stmt[0].b, sec[0].b, &bi, &id, bytes(rnd), (uint8_t *)&pub.powf.powfstage1, 1), "pa_stage1_prove_w");
  } else {
    pub.stage = STAGE2;
    const Key &kj = keys[prevDecidingStep];
    Point stmt[11] = {b, keys[step].X, keys[step].R, prevDecidingInfo[id_].b, kj.X, kj.R, cm.phi, cm.A, cm.B,
                      curInfo[id_].Y, prevDecidingInfo[id_].Y};
    Scalar sec[6] = {keys[step].x, kj.x, cm.alpha, keys[step].r, kj.r, cm.beta};
    uint8_t bj = (uint8_t)prevDecidingBit, cbit = (uint8_t)(binaryBidStr[step] - '0');  // cbit: the bit committed to in Ci
    std::vector<Scalar> rnd = draw(11);
    check(pa_stage2_prove_w(e, stmt[0].b, sec[0].b, &bi, &bj, &cbit, &id, bytes(rnd), (uint8_t *)&pub.powf.powfstage2, 1), "pa_stage2_prove_w");
  }
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
  return pub;
}

// Reference: SEAL/bidder.cpp:1346-1377 — 16 (stage 1) or 32 (stage 2) EC_POINT_mul
// per other bidder; here one batched call.
bool Bidder::verifyRoundTwo(const std::vector<RoundTwoPub> &pubs, size_t step) {
  TimeTracker::getInstance().start(VERIFIER_CATEGORY);
  std::vector<Point> stmt;
  std::vector<uint8_t> proofs;
  std::vector<uint64_t> ids;
  for (size_t i = 0; i < pubs.size(); ++i) {
    if (i == id_) continue;
    curInfo[i].b = pubs[i].b;
    const CommitmentPerBit &cm = commitmentsBB[i][step];
    if (!junctionFlag) {
      assert(pubs[i].stage == STAGE1);
      for (const Point &p : {curInfo[i].b, curInfo[i].X, curInfo[i].Y, curInfo[i].R, cm.phi, cm.A, cm.B}) stmt.push_back(p);
      const uint8_t *q = (const uint8_t *)&pubs[i].powf.powfstage1;
      proofs.insert(proofs.end(), q, q + sizeof(NIZKPoWFStage1));
    } else {
      assert(pubs[i].stage == STAGE2);
      for (const Point &p : {curInfo[i].b, curInfo[i].X, curInfo[i].R, prevDecidingInfo[i].b, prevDecidingInfo[i].X,
                             prevDecidingInfo[i].R, cm.phi, cm.A, cm.B, curInfo[i].Y, prevDecidingInfo[i].Y})
        stmt.push_back(p);
      const uint8_t *q = (const uint8_t *)&pubs[i].powf.powfstage2;
      proofs.insert(proofs.end(), q, q + sizeof(NIZKPoWFStage2));
    }
    ids.push_back(i);
  }
  std::vector<uint8_t> v(ids.size());
  if (!junctionFlag)
    check(pa_stage1_verify(engine(), proofs.data(), bytes(stmt), ids.data(), v.data(), ids.size()), "pa_stage1_verify");
  else
    check(pa_stage2_verify(engine(), proofs.data(), bytes(stmt), ids.data(), v.data(), ids.size()), "pa_stage2_verify");
  bool ret = true;
  for (size_t k = 0; k < v.size(); ++k) {
    if (!v[k]) PRINT_ERROR((junctionFlag ? "NIZKPoWFStage2" : "NIZKPoWFStage1") << " verification failed for bidder " << ids[k]);
    ret &= v[k] != 0;
  }
  TimeTracker::getInstance().stop(VERIFIER_CATEGORY);
  return ret;
}

// Reference: SEAL/bidder.cpp:1386-1421
size_t Bidder::roundThree(const std::vector<Point> &Bs, size_t step) {
  TimeTracker::getInstance().start(BIDDER_CATEGORY);
  assert(Bs.size() == n_);
  int isInf = 1;
  check(pa_point_sum_is_inf(engine(), bytes(Bs), Bs.size(), &isInf), "pa_point_sum_is_inf");
  size_t ret = 0;
  if (!isInf) {
    // somebody encoded a 1 in this step: a deciding step
    junctionFlag = true;
    prevDecidingStep = step;
    prevDecidingBit &= (size_t)(binaryBidStr[step] - '0');
    maxBid |= ((size_t)1 << (c_ - step - 1));  // 64-bit shift (the reference shifts an int)
    prevDecidingInfo = curInfo;
    ret = 1;
  }
  TimeTracker::getInstance().stop(BIDDER_CATEGORY);
  return ret;
}
