// class BulletinBoard — the in-memory message store of the reference's
// SEAL/bulletinBoard.h:13-41 with the same accessors and the same byte
// accounting rules (65 bytes per point, 1 for infinity; minimal big-endian
// length per scalar; charged on add and on every get, SEAL/bulletinBoard.cpp:26-48, 275-288).
#ifndef PA_HOST_BULLETIN_BOARD_H
#define PA_HOST_BULLETIN_BOARD_H

#include "params.h"
#include "types.h"

#include <string>
#include <vector>

class BulletinBoard {
public:
  BulletinBoard(size_t n, size_t c);

  void addCommitmentMsg(const CommitmentPub &, size_t id);
  void addRoundOneMsg(const RoundOnePub &, size_t id);
  void addRoundTwoMsg(const RoundTwoPub &, size_t id);

  const std::vector<Point> getRoundOneXs() const;
  const std::vector<Point> getRoundTwoBs() const;

  const std::vector<CommitmentPub> &getCommitments() const;
  const std::vector<RoundOnePub> &getRoundOnePubs() const;
  const std::vector<RoundTwoPub> &getRoundTwoPubs() const;

private:
  size_t n_, c_;
  std::vector<CommitmentPub> commitments_;
  std::vector<RoundOnePub> roundOnePubs_;
  std::vector<RoundTwoPub> roundTwoPubs_;
};

#endif
