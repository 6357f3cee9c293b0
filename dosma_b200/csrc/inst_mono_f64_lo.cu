// Kernel instances for model MonoExp, arithmetic type double, 1..8 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f64_lo(const LaunchDesc& d) { return launch_model_lo<MonoExp, double>(d); }
}  // namespace dfit
