// Kernel instances for model Linear1, arithmetic type double, 9..32 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_linear_f64_hi(const LaunchDesc& d) { return launch_model_hi<Linear1, double>(d); }
}  // namespace dfit
