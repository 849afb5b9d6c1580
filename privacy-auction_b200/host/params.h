// Compile-time parameters, as the reference's SEAL/params.h:4-13.
#ifndef PA_HOST_PARAMS_H
#define PA_HOST_PARAMS_H

#define CURVE 714      // OpenSSL NID_secp256k1: the engine implements exactly this curve
#define HASH "sha256"  // Fiat-Shamir hash

#define C_MAX 32  // max length of a bid in bits

#define BIDDER_CATEGORY "bidder"
#define VERIFIER_CATEGORY "verifier"

#define ENABLE_COMMUNICATION_TRACKING
#define ENABLE_VERIFICATION

#endif
