// Protocol parameters of the SEAL host programs.
//
// The names are the ones the reference's sources use (its SEAL/params.h), because the class code
// and the command line are meant to be source-compatible with it; the values are what the engine
// implements: secp256k1 only (OpenSSL numbers it NID 714), SHA-256 challenges, bids of at most 32
// bits in the CLI (the engine itself takes up to 64).
#ifndef PA_HOST_PARAMS_H
#define PA_HOST_PARAMS_H

#include <cstddef>

enum : int { CURVE = 714 };           // the curve the engine is built for; there is no other
constexpr const char *HASH = "sha256";  // informational, as in the reference (never read)
constexpr std::size_t C_MAX = 32;       // widest bid std::bitset<C_MAX> prints

// accounting categories of TimeTracker / DataTracker
constexpr const char *BIDDER_CATEGORY = "bidder";
constexpr const char *VERIFIER_CATEGORY = "verifier";

// feature switches (tested with #ifdef, hence macros)
#define ENABLE_COMMUNICATION_TRACKING 1  // count published bytes the way the reference's BulletinBoard does
#define ENABLE_VERIFICATION 1            // main() calls the verify* methods after every phase

#endif
