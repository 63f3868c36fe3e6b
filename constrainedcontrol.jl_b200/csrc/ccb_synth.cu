// csrc/ccb_synth.cu -- linearsystem / dlqr / LQR + TrackingLQR synthesis entry points (K2, K3/K4)
#include "ccb_internal.h"
using namespace ccb;

extern "C" {

int ccb_linearize(ccb_mech*, int32_t, int32_t, const double*, const double*, const int32_t*, int32_t, const ccb_sim_opts*, double*, double*, double*,
                  double*, double*) {
    set_error("ccb_linearize: not implemented yet");
    return CCB_E_UNSUPPORTED;
}
int ccb_dlqr(int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, const double*, const double*, const double*, const double*, const double*, int64_t,
             const double*, int64_t, double*, int32_t*, int32_t*) {
    set_error("ccb_dlqr: not implemented yet");
    return CCB_E_UNSUPPORTED;
}
int ccb_lqr_synthesize(ccb_mech*, int32_t, int32_t, const double*, const double*, const int32_t*, int32_t, const double*, int64_t, const double*,
                       int64_t, int32_t, const ccb_sim_opts*, double*, int32_t*, int32_t*) {
    set_error("ccb_lqr_synthesize: not implemented yet");
    return CCB_E_UNSUPPORTED;
}
int ccb_tracking_synthesize(ccb_mech*, int32_t, int32_t, int32_t, const double*, const double*, const int32_t*, int32_t, const double*,
                            const double*, const ccb_sim_opts*, double*, int32_t*, int32_t*) {
    set_error("ccb_tracking_synthesize: not implemented yet");
    return CCB_E_UNSUPPORTED;
}
}
