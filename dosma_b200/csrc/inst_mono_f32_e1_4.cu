// Kernel instances for model MonoExp, arithmetic type float, 1..4 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_e1_4(const LaunchDesc& d) { return launch_range<MonoExp, float, 1, 4>(d); }
}  // namespace dfit
