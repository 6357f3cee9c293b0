// Kernel instances for model Linear1, arithmetic type float (all echo-count buckets).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_linear_f32(const LaunchDesc& d) { return launch_model<Linear1, float>(d); }
}  // namespace dfit
