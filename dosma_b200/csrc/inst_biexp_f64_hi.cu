// Kernel instances for model BiExp, arithmetic type double, 9..32 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_biexp_f64_hi(const LaunchDesc& d) { return launch_model_hi<BiExp, double>(d); }
}  // namespace dfit
