#include "bulletinBoard.h"

#include "trackers.h"

namespace {
#ifdef ENABLE_COMMUNICATION_TRACKING
// a wire record is npts points followed by nsc scalars
void track(const std::string &cat, const void *rec, size_t npts, size_t nsc) {
  const uint8_t *p = static_cast<const uint8_t *>(rec);
  size_t total = 0;
  for (size_t i = 0; i < npts; ++i, p += 64) {
    uint8_t nz = 0;
    for (int k = 0; k < 64; ++k) nz |= p[k];
    total += nz ? 65 : 1;  // EC_POINT_point2oct(..., UNCOMPRESSED) length
  }
  for (size_t i = 0; i < nsc; ++i, p += 32) {
    size_t lead = 0;
    while (lead < 32 && p[lead] == 0) ++lead;
    total += 32 - lead;  // BN_num_bytes
  }
  DataTracker::getInstance().addData(cat, total);
}
void trackCommitment(const std::string &cat, const CommitmentPerBit &c) {
  track(cat, &c.phi, 3, 0);
  track(cat, &c.pokdlogA, 1, 1);
  track(cat, &c.pokdlogB, 1, 1);
  track(cat, &c.powfcom, 4, 3);
}
void trackRoundOne(const std::string &cat, const RoundOnePub &p) {
  track(cat, &p.X, 2, 0);
  track(cat, &p.pokdlogX, 1, 1);
  track(cat, &p.pokdlogR, 1, 1);
}
void trackRoundTwo(const std::string &cat, const RoundTwoPub &p) {
  track(cat, &p.b, 1, 0);
  DataTracker::getInstance().addData(cat, sizeof(p.stage));
  if (p.stage == STAGE1)
    track(cat, &p.powf.powfstage1, 8, 5);
  else
    track(cat, &p.powf.powfstage2, 16, 10);
}
#endif
}  // namespace

BulletinBoard::BulletinBoard(size_t n, size_t c) : n_(n), c_(c) {
  commitments_.resize(n_);
  roundOnePubs_.resize(n_);
  roundTwoPubs_.resize(n_);
  for (auto &cp : commitments_) cp.resize(c_);
}

void BulletinBoard::addCommitmentMsg(const CommitmentPub &commitment, size_t id) {
  commitments_[id] = commitment;
#ifdef ENABLE_COMMUNICATION_TRACKING
  DataTracker::getInstance().addData(BIDDER_CATEGORY, sizeof(id));
  for (auto &c : commitment) trackCommitment(BIDDER_CATEGORY, c);
#endif
}

void BulletinBoard::addRoundOneMsg(const RoundOnePub &pub, size_t id) {
  roundOnePubs_[id] = pub;
#ifdef ENABLE_COMMUNICATION_TRACKING
  DataTracker::getInstance().addData(BIDDER_CATEGORY, sizeof(id));
  trackRoundOne(BIDDER_CATEGORY, pub);
#endif
}

void BulletinBoard::addRoundTwoMsg(const RoundTwoPub &pub, size_t id) {
  roundTwoPubs_[id] = pub;
#ifdef ENABLE_COMMUNICATION_TRACKING
  DataTracker::getInstance().addData(BIDDER_CATEGORY, sizeof(id));
  trackRoundTwo(BIDDER_CATEGORY, pub);
#endif
}

const std::vector<Point> BulletinBoard::getRoundOneXs() const {
  std::vector<Point> Xs;
  Xs.reserve(roundOnePubs_.size());
  for (auto &p : roundOnePubs_) {
    Xs.push_back(p.X);
#ifdef ENABLE_COMMUNICATION_TRACKING
    track(BIDDER_CATEGORY, &p.X, 1, 0);
#endif
  }
  return Xs;
}

const std::vector<Point> BulletinBoard::getRoundTwoBs() const {
  std::vector<Point> Bs;
  Bs.reserve(roundTwoPubs_.size());
  for (auto &p : roundTwoPubs_) {
    Bs.push_back(p.b);
#ifdef ENABLE_COMMUNICATION_TRACKING
    track(BIDDER_CATEGORY, &p.b, 1, 0);
#endif
  }
  return Bs;
}

const std::vector<CommitmentPub> &BulletinBoard::getCommitments() const {
#ifdef ENABLE_COMMUNICATION_TRACKING
  for (auto &cp : commitments_)
    for (auto &c : cp) trackCommitment(VERIFIER_CATEGORY, c);
#endif
  return commitments_;
}

const std::vector<RoundOnePub> &BulletinBoard::getRoundOnePubs() const {
#ifdef ENABLE_COMMUNICATION_TRACKING
  for (auto &p : roundOnePubs_) trackRoundOne(VERIFIER_CATEGORY, p);
#endif
  return roundOnePubs_;
}

const std::vector<RoundTwoPub> &BulletinBoard::getRoundTwoPubs() const {
#ifdef ENABLE_COMMUNICATION_TRACKING
  for (auto &p : roundTwoPubs_) trackRoundTwo(VERIFIER_CATEGORY, p);
#endif
  return roundTwoPubs_;
}
