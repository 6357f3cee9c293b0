// Kernel instances for model MonoExp, arithmetic type float, 9..32 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_hi(const LaunchDesc& d) { return launch_model_hi<MonoExp, float>(d); }
}  // namespace dfit
