// Kernel instances for model MonoExp, arithmetic type float, 17..32 echoes (one predicated instance).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_many(const LaunchDesc& d) { return launch_many<MonoExp, float>(d); }
}  // namespace dfit
