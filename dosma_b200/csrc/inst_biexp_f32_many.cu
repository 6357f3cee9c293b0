// Kernel instances for model BiExp, arithmetic type float, 17..32 echoes (one predicated instance).
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f32_many(const LaunchDesc& d) { return launch_many<BiExp, float>(d); }
}  // namespace dfit
