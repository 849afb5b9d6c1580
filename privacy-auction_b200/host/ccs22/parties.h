// class Bidder / class Evaluator / class BulletinBoard of the CCS22 protocol with the public
// interfaces of the reference's CCS22/bidder.h:16-31, CCS22/evaluator.h:12-23 and
// CCS22/bulletinBoard.h:15-30, on the CUDA engine's C ABI.
#ifndef PA_HOST_CCS22_PARTIES_H
#define PA_HOST_CCS22_PARTIES_H

#include "params.h"
#include "types.h"

#include <cstddef>
#include <string>
#include <vector>

namespace ccs22 {

class BulletinBoard {
public:
  BulletinBoard(size_t n, size_t c);

  const PubParams &getPubParams() const;
  void addCommitmentMsg(size_t id, const Point &com);
  void addPublicKeyMsg(size_t id, const std::vector<Point> &pubKeys);
  const std::vector<Point> getPublicKeysByStep(size_t step) const;

  void addOTR1Vec(const OT_R1_VEC &);
  OT_R1 getOTR1(size_t id_wo_e) const;
  void addOTS(size_t id_wo_e, const OT_S &);
  OT_S_VEC getOTSVec() const;

  void addd(size_t d);
  size_t getd() const;

private:
  size_t n_, c_;
  PubParams pubParams_;
  std::vector<Point> commitments_;
  std::vector<std::vector<Point>> pubKeys_;
  OT_R1_VEC ot_r1_vec_;
  OT_S_VEC ot_s_vec_;
  size_t d_ = 0;
};

class Bidder {
public:
  Bidder(size_t id, size_t n, size_t c, const PubParams &);              // pseudo-random c-bit bid
  Bidder(size_t id, size_t n, size_t c, const PubParams &, size_t bid);  // extension: explicit bid
  virtual ~Bidder() = default;

  size_t getId();
  size_t getBid();
  size_t getMaxBid();

  virtual void setup();
  const Point &getCommitments() const;
  const std::vector<Point> &getPubKeys() const;

  virtual void BESEncode(const std::vector<Point> &, size_t step);
  OT_S OTSend(size_t step, const OT_R1 &);
  void checkIfEnterDeciderRound(size_t step, size_t d);

protected:
  struct PrivKey {
    Scalar x, r;
  };
  void init(size_t bid);
  std::vector<Scalar> draw(size_t k);
  std::vector<Scalar> draw256(size_t k);
  void commit(const std::vector<Scalar> &hashed);  // H and Com = g^bid g1^H + h^R
  virtual void setupInner();
  void BESEncodeInner(const std::vector<Point> &, size_t step);

  std::vector<PrivKey> privKeys;
  size_t id_, bid_, c_, n_;
  std::vector<Point> pubKeys;
  PubParams pp;
  Scalar R, H;
  Point Com, B;
  bool inRaceFlag;
  size_t d, maxBid;
  std::string binaryBidStr;
  unsigned long long drawCounter;

private:
  std::vector<Scalar> randomS, randomT;  // randomness in OT
};

class Evaluator : public Bidder {
public:
  Evaluator(size_t id, size_t n, size_t c, const PubParams &);
  Evaluator(size_t id, size_t n, size_t c, const PubParams &, size_t bid);

  void setup() override;
  void BESEncode(const std::vector<Point> &, size_t step) override;

  OT_R1_VEC OTReceive1(size_t step);
  size_t OTReceive2(size_t step, const OT_S_VEC &);

private:
  std::vector<std::vector<Scalar>> randomBeta;  // [step][other bidder]
  void setupInner() override;
};

}  // namespace ccs22
#endif
