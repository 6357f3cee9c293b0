// Kernel instances for model Linear1, arithmetic type float, 9..32 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_linear_f32_hi(const LaunchDesc& d) { return launch_model_hi<Linear1, float>(d); }
}  // namespace dfit
