// Kernel instances for model BiExp, arithmetic type float, 1..8 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f32_e1_8(const LaunchDesc& d) { return launch_range<BiExp, float, 1, 8>(d); }
}  // namespace dfit
