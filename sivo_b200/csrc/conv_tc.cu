// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands in this file).
#include "segnet.h"
namespace sivo {
struct ConvTcPlan {};
bool conv_tc_supported(const Op&, const TensorView&, const TensorView&) { return false; }
std::shared_ptr<ConvTcPlan> conv_tc_plan(const Op&, const TensorView&, const TensorView&, const void*) { return nullptr; }
void conv_tc_launch(const ConvTcPlan&, const Op&, cudaStream_t) { fail(SIVO_EINVAL, "tcgen05 convolution is not built"); }
}  // namespace sivo
