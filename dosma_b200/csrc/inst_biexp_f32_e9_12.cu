// Kernel instances for model BiExp, arithmetic type float, 9..12 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f32_e9_12(const LaunchDesc& d) { return launch_range<BiExp, float, 9, 12>(d); }
}  // namespace dfit
