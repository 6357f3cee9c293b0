// TEST TOOLING (kernel A/B runs): stand-ins for every kernel-instance translation unit except inst_mono_f32_lo.cu, so
// that a variant library of the headline kernel is a few MB instead of 70 (it has to travel to the GPU box).
#include "../../../dosma_b200/csrc/kernel_common.cuh"
namespace dfit {
#define STUB(name) cudaError_t name(const LaunchDesc&) { return cudaErrorNotSupported; }
STUB(launch_mono_f32_hi) STUB(launch_mono_f64_lo) STUB(launch_mono_f64_hi) STUB(launch_biexp_f32_lo) STUB(launch_biexp_f32_hi)
STUB(launch_biexp_f64_lo) STUB(launch_biexp_f64_hi) STUB(launch_linear_f32_lo) STUB(launch_linear_f32_hi)
STUB(launch_linear_f64_lo) STUB(launch_linear_f64_hi)
}  // namespace dfit
