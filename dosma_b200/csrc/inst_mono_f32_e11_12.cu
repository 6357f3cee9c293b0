// Kernel instances for model MonoExp, arithmetic type float, 11..12 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_e11_12(const LaunchDesc& d) { return launch_range<MonoExp, float, 11, 12>(d); }
}  // namespace dfit
