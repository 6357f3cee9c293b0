// Kernel instances for model MonoExp, arithmetic type float, 1..8 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_lo(const LaunchDesc& d) { return launch_model_lo<MonoExp, float>(d); }
}  // namespace dfit
