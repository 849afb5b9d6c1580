// class Bidder — public interface of the reference's SEAL/bidder.h:25-42, on the
// CUDA engine.  Every method that the reference implements with loops of
// EC_POINT_mul / BN_mod_mul / SHA-256 (SEAL/bidder.cpp:90-1421) is one or a few
// batched calls into the engine's C ABI here; the class itself only keeps the
// protocol state machine (junction flag, previous deciding step, key and
// commitment bookkeeping) and decides WHAT to ask the engine for.
#ifndef PA_HOST_BIDDER_H
#define PA_HOST_BIDDER_H

#include "params.h"
#include "print.h"
#include "types.h"

#include <cstddef>
#include <string>
#include <vector>

class Bidder {
public:
  // c-bit pseudo-random bid derived from the configured seed (the reference draws
  // it from std::random_device, SEAL/bidder.cpp:27-30)
  Bidder(size_t id, size_t n, size_t c);
  // extension: explicit bid (used by the parity tests and `SEAL --bids`)
  Bidder(size_t id, size_t n, size_t c, size_t bid);

  size_t getId();
  size_t getBid();
  size_t getMaxBid();

  CommitmentPub commitBid();
  RoundOnePub roundOne(size_t step);
  RoundTwoPub roundTwo(const std::vector<Point> &Xs, size_t step);
  size_t roundThree(const std::vector<Point> &Bs, size_t step);

  bool verifyCommitment(const std::vector<CommitmentPub> &pubs);
  bool verifyRoundOne(const std::vector<RoundOnePub> &pubs);
  bool verifyRoundTwo(const std::vector<RoundTwoPub> &pubs, size_t step);

private:
  struct Commitment {
    Point phi, A, B;
    Scalar alpha, beta;
  };
  struct Key {
    Point X, R;
    Scalar x, r;
  };

  void init(size_t bid);
  std::vector<Scalar> draw(size_t k);

  size_t id_, bid_, c_, n_;
  size_t maxBid;
  std::string binaryBidStr;
  bool junctionFlag;        // whether the junction has been reached
  size_t prevDecidingStep;
  size_t prevDecidingBit;   // bit d in the paper
  uint64_t drawCounter;     // position in this bidder's PA stream

  std::vector<Commitment> commitments;       // own, private part included
  std::vector<Key> keys;                     // own, per step
  std::vector<CommitmentPub> commitmentsBB;  // everybody's, from the bulletin board
  AuxilaryInfo curInfo, prevDecidingInfo;
};

#endif
