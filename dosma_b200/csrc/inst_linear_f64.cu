// Kernel instances for model Linear1, arithmetic type double (all echo-count buckets).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_linear_f64(const LaunchDesc& d) { return launch_model<Linear1, double>(d); }
}  // namespace dfit
