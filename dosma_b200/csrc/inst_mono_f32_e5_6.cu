// Kernel instances for model MonoExp, arithmetic type float, 5..6 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_e5_6(const LaunchDesc& d) { return launch_range<MonoExp, float, 5, 6>(d); }
}  // namespace dfit
