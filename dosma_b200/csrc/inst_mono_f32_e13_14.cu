// Kernel instances for model MonoExp, arithmetic type float, 13..14 echoes.
#include "fit_kernel.cuh"

namespace dfit {
cudaError_t launch_mono_f32_e13_14(const LaunchDesc& d) { return launch_range<MonoExp, float, 13, 14>(d); }
}  // namespace dfit
