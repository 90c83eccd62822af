// placeholder until the tcgen05 kernel lands
#include "common.h"
#include "kernels.h"
namespace egn {
struct TcConvPlan {};
bool tc_conv_supported(const ConvArgs&) { return false; }
int tc_conv_plan_create(const ConvArgs&, const float*, TcConvPlan**) { set_error("tc path not built"); return EGN_ERR_INVALID; }
void tc_conv_plan_destroy(TcConvPlan*) {}
int launch_conv_tc(TcConvPlan*, const ConvArgs&, cudaStream_t) { set_error("tc path not built"); return EGN_ERR_INVALID; }
size_t tc_conv_plan_weight_bytes(const TcConvPlan*) { return 0; }
}
