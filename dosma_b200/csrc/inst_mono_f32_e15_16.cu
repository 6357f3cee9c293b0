// Kernel instances for model MonoExp, arithmetic type float, 15..16 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_e15_16(const LaunchDesc& d) { return launch_range<MonoExp, float, 15, 16>(d); }
}  // namespace dfit
