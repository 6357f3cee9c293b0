// Kernel instances for model MonoExp, arithmetic type float, 9..10 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_e9_10(const LaunchDesc& d) { return launch_range<MonoExp, float, 9, 10>(d); }
}  // namespace dfit
