// Message types of the CCS22 protocol, names as in the reference's CCS22/types.h:7-30, with
// 64-byte wire points instead of EC_POINT*.
#ifndef PA_HOST_CCS22_TYPES_H
#define PA_HOST_CCS22_TYPES_H
#include "../types.h"

struct PubParams {
  Point g, g1, h;  // the group and its order are fixed: secp256k1 (CURVE 714)
};
struct OT_R1 {
  // T1 = g1 (as in the reference, CCS22/types.h:16)
  Point T2, G, H;
};
typedef std::vector<OT_R1> OT_R1_VEC;
struct OT_S {
  Point z, C0, C1;
};
typedef std::vector<OT_S> OT_S_VEC;
static_assert(sizeof(OT_R1) == 192 && sizeof(OT_S) == 192, "OT messages must equal their wire records");
#endif
