// Kernel instances for model BiExp, arithmetic type float, 9..32 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f32_hi(const LaunchDesc& d) { return launch_model_hi<BiExp, float>(d); }
}  // namespace dfit
