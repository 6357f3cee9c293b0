// Private to the library: handle layout and error helpers shared by the translation units that
// implement the C-ABI (dfit_api.cu, qdess.cu, metrics.cu).
#pragma once

#include <cuda_runtime.h>
#include <sched.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/dfit.h"
#include "fit_kernel.cuh"

namespace dfit {

int fail(int code, const char* fmt, ...);
const char* last_error();

#define CUDA_TRY(expr)                                                                                            \
  do {                                                                                                            \
    cudaError_t _e = (expr);                                                                                      \
    if (_e != cudaSuccess)                                                                                        \
      return dfit::fail(_e == cudaErrorMemoryAllocation ? DFIT_ERR_OOM : DFIT_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                                              \
  } while (0)

constexpr int kSlots = 3;  // pipeline depth of the host entry points

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

int ensure(DevBuf& b, size_t bytes);

// Page-locked host staging (cudaHostAlloc): pageable caller buffers are copied through these by a few host
// threads so that the GPU side of dfit_fit_host always DMAs at the full PCIe rate.
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
};

int ensure_host(HostBuf& b, size_t bytes, const cpu_set_t* cpus = nullptr);

struct Slot {
  cudaStream_t stream = nullptr;
  DevBuf y, mask, p0, popt, r2, status, niter, index;
  DevBuf lm;          // LM tail of the dense fast-path kernel: two alternating counters (16 bytes) + the voxel list
  int lm_parity = 0;
  HostBuf hin, hpopt, hr2;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;  // staging consumed by the H2D copy / filled by the D2H copies
};

}  // namespace dfit

struct dfit_handle {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  dfit::Slot slots[dfit::kSlots];
  unsigned long long* counters = nullptr;  // device, kStatSlots x CNT_COUNT entries
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  bool ev_valid = false;
  cudaStream_t ev_stream = nullptr;
  int64_t last_n = 0;
  int last_launches = 0;
  float last_total_ms = 0.f;
  float host_kernel_ms = -1.f;
  dfit::GatherArgs g = {};     // fused all-gather (dfit_set_gather / dfit_set_gather_ex); g.world == 0: off
  int64_t gather_rows = 0;     // rows in every map
  // CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs local_cpulist): the host threads that fill / drain /
  // widen the staging blocks run there, and the page-locked staging is allocated from there (DFIT_HOST_AFFINITY=0: off)
  cpu_set_t local_cpus;
  bool have_local_cpus = false;
  dfit::DevBuf scratch;    // small device scratch (qDESS maxima, metrics partials)
  dfit::DevBuf index_buf;  // compacted voxel list of the masked device path
  dfit::DevBuf lm_buf;     // LM tail of the dense fast-path kernel (device path): counters + voxel list
  int lm_parity = 0;
};
