// Kernel instances for model BiExp, arithmetic type double (all echo-count buckets).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f64(const LaunchDesc& d) { return launch_model<BiExp, double>(d); }
}  // namespace dfit
