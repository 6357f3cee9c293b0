// Kernel instances for model MonoExp, arithmetic type float, 7..8 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_e7_8(const LaunchDesc& d) { return launch_range<MonoExp, float, 7, 8>(d); }
}  // namespace dfit
