// As the reference's CCS22/params.h.
#ifndef PA_HOST_CCS22_PARAMS_H
#define PA_HOST_CCS22_PARAMS_H
#define CURVE 714
#define HASH "sha256"
#define C_MAX 32
#define BIDDER_CATEGORY "bidder"
#define EVALUATOR_CATEGORY "evaluator"
#define BIDDER_AND_EVALUATOR_CATEGORY "bidder_and_evaluator"
#define VERIFIER_CATEGORY "verifier"
#define ENABLE_COMMUNICATION_TRACKING
#define ENABLE_VERIFICATION
#endif
