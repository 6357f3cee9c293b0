// Kernel instances for model MonoExp, arithmetic type float (all echo-count buckets).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32(const LaunchDesc& d) { return launch_model<MonoExp, float>(d); }
}  // namespace dfit
