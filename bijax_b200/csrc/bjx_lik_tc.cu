// placeholder until the tcgen05 kernel lands
#include "bjx_elbo.cuh"
namespace bjx {
bool tc_eligible(const BjxElboArgs&) { return false; }
int tc_plan(const BjxElboArgs&, Plan&) { return fail(BJX_ERR_UNSUPPORTED, "tcgen05 path not built"); }
size_t tc_theta_bytes(const BjxElboArgs&, const Plan&) { return 0; }
int launch_lik_tc(const BjxElboArgs&, const Plan&, char*, cudaStream_t) { return fail(BJX_ERR_UNSUPPORTED, "tcgen05 path not built"); }
}  // namespace bjx
