// Message types of the SEAL protocol with the names and field order of the
// reference's SEAL/types.h:13-144.  The reference holds raw EC_POINT* / BIGNUM*;
// here every point is its 64-byte wire form (X || Y big-endian, zeros = infinity)
// and every scalar 32 bytes big-endian, so each struct IS its wire record: a
// vector of them can be handed to the engine's C ABI (include/pa_engine.h) or
// written to a transcript without any conversion.
#ifndef PA_HOST_TYPES_H
#define PA_HOST_TYPES_H

#include <cstdint>
#include <cstring>
#include <vector>

struct Point {
  uint8_t b[64];
  bool isInfinity() const {
    uint8_t z = 0;
    for (int i = 0; i < 64; ++i) z |= b[i];
    return z == 0;
  }
};
struct Scalar {
  uint8_t b[32];
};

/** Non-interactive zero-knowledge proof of knowledge of discrete logarithm */
struct NIZKPoKDLog {
  Point eps;
  Scalar rho;
};

/** NIZK proof of well-formedness of commitments */
struct NIZKPoWFCom {
  Point eps11, eps12, eps21, eps22;
  Scalar rho1, rho2;
  Scalar ch2;  // only ch2 is sent
};

/** NIZK proof of well-formedness of cryptograms in Stage 1 (before junction) */
struct NIZKPoWFStage1 {
  Point eps11, eps12, eps13, eps14, eps21, eps22, eps23, eps24;
  Scalar rho11, rho12, rho21, rho22;
  Scalar ch2;
};

/** NIZK proof of well-formedness of cryptograms in Stage 2 (after junction) */
struct NIZKPoWFStage2 {
  Point eps11, eps12, eps13, eps11prime, eps12prime, eps13prime;
  Point eps21, eps22, eps23, eps21prime, eps22prime, eps23prime;
  Point eps31, eps32, eps31prime, eps32prime;
  Scalar rho11, rho12, rho13, rho21, rho22, rho23, rho31, rho32;
  Scalar ch2, ch3;
};

struct CommitmentPerBit {
  Point phi, A, B;
  NIZKPoKDLog pokdlogA, pokdlogB;
  NIZKPoWFCom powfcom;
};
typedef std::vector<CommitmentPerBit> CommitmentPub;

struct RoundOnePub {
  Point X, R;
  NIZKPoKDLog pokdlogX, pokdlogR;
};

struct AuxilaryInfoPerBidder {
  Point b, Y, X, R;
};
typedef std::vector<AuxilaryInfoPerBidder> AuxilaryInfo;

typedef enum { STAGE1, STAGE2 } stagetype_t;

struct RoundTwoPub {
  Point b;
  stagetype_t stage;
  union {
    NIZKPoWFStage1 powfstage1;
    NIZKPoWFStage2 powfstage2;
  } powf;
};

static_assert(sizeof(NIZKPoKDLog) == 96 && sizeof(NIZKPoWFCom) == 352 && sizeof(NIZKPoWFStage1) == 672 &&
                  sizeof(NIZKPoWFStage2) == 1344 && sizeof(CommitmentPerBit) == 736 && sizeof(RoundOnePub) == 320,
              "message structs must equal their wire records");

#endif
