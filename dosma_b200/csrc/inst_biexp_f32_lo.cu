// Kernel instances for model BiExp, arithmetic type float, 1..8 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f32_lo(const LaunchDesc& d) { return launch_model_lo<BiExp, float>(d); }
}  // namespace dfit
