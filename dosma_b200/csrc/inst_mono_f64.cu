// Kernel instances for model MonoExp, arithmetic type double (all echo-count buckets).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f64(const LaunchDesc& d) { return launch_model<MonoExp, double>(d); }
}  // namespace dfit
