// Kernel instances for model Linear1, arithmetic type double, 1..8 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_linear_f64_lo(const LaunchDesc& d) { return launch_model_lo<Linear1, double>(d); }
}  // namespace dfit
