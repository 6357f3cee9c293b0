// Kernel instances for model MonoExp, arithmetic type double, 9..32 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f64_hi(const LaunchDesc& d) { return launch_model_hi<MonoExp, double>(d); }
}  // namespace dfit
