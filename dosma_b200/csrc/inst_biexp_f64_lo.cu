// Kernel instances for model BiExp, arithmetic type double, 1..8 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f64_lo(const LaunchDesc& d) { return launch_model_lo<BiExp, double>(d); }
}  // namespace dfit
