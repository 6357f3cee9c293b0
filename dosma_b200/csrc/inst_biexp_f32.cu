// Kernel instances for model BiExp, arithmetic type float (all echo-count buckets).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f32(const LaunchDesc& d) { return launch_model<BiExp, float>(d); }
}  // namespace dfit
