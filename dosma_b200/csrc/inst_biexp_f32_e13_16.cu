// Kernel instances for model BiExp, arithmetic type float, 13..16 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f32_e13_16(const LaunchDesc& d) { return launch_range<BiExp, float, 13, 16>(d); }
}  // namespace dfit
