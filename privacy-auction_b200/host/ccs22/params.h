// Protocol parameters of the CCS22 host programs (names as in the reference's CCS22/params.h,
// see ../params.h for the rationale).
#ifndef PA_HOST_CCS22_PARAMS_H
#define PA_HOST_CCS22_PARAMS_H

#include <cstddef>

namespace ccs22_params {
enum : int { CURVE = 714 };
constexpr std::size_t C_MAX = 32;
constexpr const char *BIDDER_CATEGORY = "bidder";
constexpr const char *EVALUATOR_CATEGORY = "evaluator";
constexpr const char *BIDDER_AND_EVALUATOR_CATEGORY = "bidder_and_evaluator";  // setup traffic both roles pay
constexpr const char *VERIFIER_CATEGORY = "verifier";                          // unused: CCS22 has no verification phase yet
}  // namespace ccs22_params
using namespace ccs22_params;

#define ENABLE_COMMUNICATION_TRACKING 1

#endif
