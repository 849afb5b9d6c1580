// The public bulletin board of a SEAL auction: an in-memory store of what every bidder has
// published, one slot per bidder and message type.  Method names follow the reference's
// BulletinBoard so that main() reads the same; storage is by value (wire records), not pointers
// into OpenSSL objects.
//
// Byte accounting (DataTracker) reproduces the reference's rules exactly — 65 bytes per point
// (1 for infinity), the minimal big-endian length per scalar, the size of the id / stage tag —
// charged to the "bidder" category when a message is added and to the "verifier" category every
// time a verifier fetches the board; `tests/test_gpu_cli.py` compares the totals with the
// reference's own numbers.
#ifndef PA_HOST_BULLETIN_BOARD_H
#define PA_HOST_BULLETIN_BOARD_H

#include "params.h"
#include "types.h"

#include <cstddef>
#include <vector>

class BulletinBoard {
public:
  BulletinBoard(std::size_t bidders, std::size_t bits);

  // ---- publishing (charged to the publishing bidder) ----
  void addCommitmentMsg(const CommitmentPub &perBitCommitments, std::size_t bidderId);
  void addRoundOneMsg(const RoundOnePub &keysAndProofs, std::size_t bidderId);
  void addRoundTwoMsg(const RoundTwoPub &cryptogramAndProof, std::size_t bidderId);

  // ---- what a bidder needs for its own next move (charged to the bidder) ----
  const std::vector<Point> getRoundOneXs() const;  // every X of the current step, id order
  const std::vector<Point> getRoundTwoBs() const;  // every cryptogram b of the current step

  // ---- what a verifier reads (charged to the verifier, every call) ----
  const std::vector<CommitmentPub> &getCommitments() const;
  const std::vector<RoundOnePub> &getRoundOnePubs() const;
  const std::vector<RoundTwoPub> &getRoundTwoPubs() const;

private:
  std::size_t n_, c_;
  std::vector<CommitmentPub> commitments_;  // [bidder][bit]
  std::vector<RoundOnePub> roundOnePubs_;   // [bidder], overwritten every step
  std::vector<RoundTwoPub> roundTwoPubs_;   // [bidder], overwritten every step
};

#endif
