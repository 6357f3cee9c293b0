// Fit kernels and their launcher.
//
//   kernel_common.cuh  -- kernel arguments, loads, per-voxel store / fused epilogue / gather, statistics, TMA primitives
//   fit_kernels1.cuh   -- one voxel per lane: mask_compact_kernel, fit_kernel (every model, LM with the one-voxel fast
//                         path in front for the mono-exponential model), fit_kernel_tma
//   mono2_kernels.cuh  -- two voxels per lane, mono-exponential fast path: fit_kernel_mono2 (plain loads),
//                         fit_kernel_mono2_list (mask path), fit_kernel_mono2_tma (persistent, TMA-staged: the headline)
//   lmq_kernel.cuh     -- the LM in rounds with a per-warp stack of suspended fits (fit_kernel_lmq): every fp32 fit that
//                         goes straight to the LM (bi-exponential, linear, mono-exponential with fast_path = 0)
//   this file          -- launch_one: which kernel a launch description gets, and the per-echo-count instances
//
// Together they replace the N-voxel loop of dosma/core/fitting.py:855-868 and fuse what the reference does
// around it: dtype up-cast (:711, SciPy's asarray(float)), mask select/scatter (:199-215), the
// log-linear initial guess (:701-718), `_process_params` (:109-146) and rounding (:734-737).
#pragma once

#include "fit_kernels1.cuh"
#include "kernel_common.cuh"
#include "lmq_kernel.cuh"
#include "mono2_kernels.cuh"

namespace dfit {

#if defined(__CUDACC__)
template <class M, typename T, int EMAX, bool EXACT>
inline cudaError_t launch_one(const LaunchDesc& d) {
  KernelArgs<T, EMAX> a;
  fill_args<T, EMAX>(d, a);
  if constexpr (M::MONO && EXACT && sizeof(T) == 4 && EMAX >= 3) {
    // dense fast path, two voxels per lane (see fit_kernel_mono2 for what it needs)
    const bool dt_ok = d.y_dtype == DT_F32 || d.y_dtype == DT_I16 || d.y_dtype == DT_U16;
    const size_t pair_bytes = 2 * dtype_size(d.y_dtype);
    // With the fused all-gather only the TMA kernel qualifies, for the two row formats it stores as vectors: one
    // parameter + r2 (8-byte rows, the T2 map) or all three columns; the rows of a lane's voxel pair must be 8 / 16-byte
    // aligned (even first row).  Everything else gathers through the one-voxel kernel below.
    const bool gather_ok = d.g.world == 0 || (d.tmap2 != nullptr && d.y_dtype == DT_F32 && (d.g.ncols == 2 || d.g.ncols == 3) &&
                                              d.g.row0 % 2 == 0);
    if (d.fast_path == 1 && !a.vo.has_bounds && d.mask == nullptr && gather_ok && dt_ok && d.layout == LAYOUT_PLANAR &&
        (d.popt != nullptr || d.g.world > 0) && reinterpret_cast<uintptr_t>(d.y) % pair_bytes == 0 && d.ld % 2 == 0 &&
        reinterpret_cast<uintptr_t>(d.popt) % 16 == 0 && reinterpret_cast<uintptr_t>(d.r2) % 8 == 0) {
      if (d.tmap2 != nullptr) {  // persistent, tiles staged through shared memory by TMA
        auto kfn = d.g.world > 0         ? fit_kernel_mono2_tma<M, EMAX, true, float>
                   : d.y_dtype == DT_I16 ? fit_kernel_mono2_tma<M, EMAX, false, short>
                   : d.y_dtype == DT_U16 ? fit_kernel_mono2_tma<M, EMAX, false, unsigned short>
                                         : fit_kernel_mono2_tma<M, EMAX, false, float>;
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kM2Warps * 32, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        const int64_t n_tiles = (d.n_vox + kM2Tile - 1) / kM2Tile;
        int64_t g = (int64_t)d.sm_count * per_sm;
        const int64_t needed = (n_tiles + kM2Warps - 1) / kM2Warps;
        if (g > needed) g = needed;
        const bool tail = d.lm_list != nullptr && d.g.world == 0 && lmq_config(M::P).enabled;
        const int par = tail ? (*d.lm_parity & 1) : 0;
        if (tail) {
          a.lm_list = d.lm_list;
          a.lm_count = d.lm_head + par;
        }
        kfn<<<(unsigned)g, kM2Warps * 32, 0, d.stream>>>(a, *d.tmap2);
        e = cudaGetLastError();
        if (e != cudaSuccess || !tail) return e;
        // the LM tail: voxels neither the straight-line fit nor the Newton loop settled (a.lm_list), by the LM in rounds
        KernelArgs<T, EMAX> at = a;
        at.index = d.lm_list;
        at.index_count = d.lm_head + par;
        at.lm_list = nullptr;
        at.lm_count = nullptr;
        at.lm_count_next = d.lm_head + (par ^ 1);  // zeroed by the tail kernel: the next launch counts there
        *d.lm_parity = par ^ 1;
        return launch_lmq<M, T, EMAX>(d, at, true);
      }
      const int64_t per_cta = 2 * kBlock2;
      fit_kernel_mono2<M, EMAX><<<(unsigned)((d.n_vox + per_cta - 1) / per_cta), kBlock2, 0, d.stream>>>(a);
      return cudaGetLastError();
    }
  }
  if constexpr (EXACT && sizeof(T) == 4) {
    if (d.tmap != nullptr) {  // TMA-staged persistent variant (the C-ABI layer checked eligibility)
      int per_sm = 0;
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_kernel_tma<M, T, EMAX>,
                                                                    kTmaWarps * 32, 0);
      if (e != cudaSuccess) return e;
      if (per_sm < 1) per_sm = 1;
      const int64_t n_tiles = (d.n_vox + kTmaTile - 1) / kTmaTile;
      int64_t blocks = (int64_t)d.sm_count * per_sm;
      const int64_t needed = (n_tiles + kTmaWarps - 1) / kTmaWarps;
      if (blocks > needed) blocks = needed;
      fit_kernel_tma<M, T, EMAX><<<(unsigned)blocks, kTmaWarps * 32, 0, d.stream>>>(a, *d.tmap);
      return cudaGetLastError();
    }
  }
  const int64_t blocks = (d.n_vox + kBlock - 1) / kBlock;
  // fits that go straight to the LM (no fast path in front), fp32, one GPU: the LM in rounds (lmq_kernel.cuh)
  // (fp32 only: with fp64 state words the rounds kernel measured slower than the plain one, 13.5 against 12.3 ms on config 4)
  [[maybe_unused]] const bool lmq = EXACT && sizeof(T) == 4 && d.g.world == 0 && lmq_config().enabled &&
                                    (!M::MONO || d.fast_path == 0 || a.vo.has_bounds) && d.n_vox < ((int64_t)1 << 32);
  if (d.mask != nullptr && d.index != nullptr) {
    // mask path: compact + fill, then fit the list with a grid sized for the SMs (grid-stride)
    cudaError_t e = cudaMemsetAsync(d.index_count, 0, sizeof(unsigned), d.stream);
    if (e != cudaSuccess) return e;
    const int64_t per_cta = 256 * kCompactPerThread;
    mask_compact_kernel<M::P, T, EMAX><<<(unsigned)((d.n_vox + per_cta - 1) / per_cta), 256, 0, d.stream>>>(a, d.index, d.index_count);
    a.index = d.index;
    a.index_count = d.index_count;
    int64_t g = (int64_t)d.sm_count * 16;
    if (g > blocks) g = blocks;
    if constexpr (M::MONO && EXACT && sizeof(T) == 4 && EMAX >= 3) {
      if (d.fast_path == 1 && !a.vo.has_bounds) {  // two voxels per lane over the list
        if (d.g.world > 0) fit_kernel_mono2_list<M, EMAX, true><<<(unsigned)g, kBlock, 0, d.stream>>>(a);
        else fit_kernel_mono2_list<M, EMAX, false><<<(unsigned)g, kBlock, 0, d.stream>>>(a);
        return cudaGetLastError();
      }
    }
    if constexpr (EXACT && sizeof(T) == 4) {
      if (lmq) return launch_lmq<M, T, EMAX>(d, a);
    }
    if (d.g.world > 0) {
      if constexpr (sizeof(T) == 4) fit_kernel<M, T, EMAX, EXACT, true><<<(unsigned)g, kBlock, 0, d.stream>>>(a);
      else return cudaErrorNotSupported;
    } else {
      fit_kernel<M, T, EMAX, EXACT, false><<<(unsigned)g, kBlock, 0, d.stream>>>(a);
    }
    return cudaGetLastError();
  }
  if constexpr (EXACT && sizeof(T) == 4) {
    if (lmq) return launch_lmq<M, T, EMAX>(d, a);
  }
  // the fused all-gather epilogue is a separate instance so that single-GPU launches do not pay its registers
  if (d.g.world > 0) {
    if constexpr (sizeof(T) == 4) fit_kernel<M, T, EMAX, EXACT, true><<<(unsigned)blocks, kBlock, 0, d.stream>>>(a);
    else return cudaErrorNotSupported;  // the gathered map is fp32 (the C-ABI layer rejects this earlier)
  } else {
    fit_kernel<M, T, EMAX, EXACT, false><<<(unsigned)blocks, kBlock, 0, d.stream>>>(a);
  }
  return cudaGetLastError();
}

// Samples live in registers, so the echo count is a template parameter: exact instances for
// E <= 16 (fully unrolled, no predicates, paired FP32 arithmetic), one predicated 32-register
// instance above that.
template <class M, typename T, int E>
inline cudaError_t launch_exact(const LaunchDesc& d) {
  if constexpr (E < M::P) {
    return cudaErrorInvalidValue;
  } else {
    return launch_one<M, T, E, true>(d);
  }
}

// Echo-count ranges: the instances of one model / arithmetic type are spread over several translation units that
// compile in parallel (the fp32 mono-exponential ones -- two-voxel, TMA and rounds kernels per echo count -- took 8 of
// the build's 8.5 minutes as two units).  launch_range covers the exact instances LO..HI (HI <= 16);
// launch_many the predicated 32-register instance above 16 echoes.
template <class M, typename T, int LO, int HI>
inline cudaError_t launch_range(const LaunchDesc& d) {
  if constexpr (LO > HI) {
    return cudaErrorInvalidValue;
  } else {
    if (d.n_echo == LO) return launch_exact<M, T, LO>(d);
    return launch_range<M, T, LO + 1, HI>(d);
  }
}
template <class M, typename T>
inline cudaError_t launch_many(const LaunchDesc& d) {
  return launch_one<M, T, 32, false>(d);
}

template <class M, typename T>
inline cudaError_t launch_model_lo(const LaunchDesc& d) {  // 1..8 echoes
  return launch_range<M, T, 1, 8>(d);
}
template <class M, typename T>
inline cudaError_t launch_model_hi(const LaunchDesc& d) {  // 9..32 echoes
  return d.n_echo <= 16 ? launch_range<M, T, 9, 16>(d) : launch_many<M, T>(d);
}
#endif

// Implemented one per translation unit so the (large, fully unrolled) instances compile in parallel.
#define DFIT_DECLARE_LAUNCH(name)                    \
  cudaError_t launch_##name##_lo(const LaunchDesc& d); \
  cudaError_t launch_##name##_hi(const LaunchDesc& d); \
  inline cudaError_t launch_##name(const LaunchDesc& d) { return d.n_echo <= 8 ? launch_##name##_lo(d) : launch_##name##_hi(d); }
DFIT_DECLARE_LAUNCH(mono_f64)
DFIT_DECLARE_LAUNCH(biexp_f64)
DFIT_DECLARE_LAUNCH(linear_f32)
DFIT_DECLARE_LAUNCH(linear_f64)
#undef DFIT_DECLARE_LAUNCH

// the fp32 mono- and bi-exponential instances: finer parts (inst_mono_f32_*.cu, inst_biexp_f32_*.cu)
cudaError_t launch_mono_f32_e1_4(const LaunchDesc& d);
cudaError_t launch_mono_f32_e5_6(const LaunchDesc& d);
cudaError_t launch_mono_f32_e7_8(const LaunchDesc& d);
cudaError_t launch_mono_f32_e9_10(const LaunchDesc& d);
cudaError_t launch_mono_f32_e11_12(const LaunchDesc& d);
cudaError_t launch_mono_f32_e13_14(const LaunchDesc& d);
cudaError_t launch_mono_f32_e15_16(const LaunchDesc& d);
cudaError_t launch_mono_f32_many(const LaunchDesc& d);
inline cudaError_t launch_mono_f32(const LaunchDesc& d) {
  const int e = d.n_echo;
  return e <= 4 ? launch_mono_f32_e1_4(d) : e <= 6 ? launch_mono_f32_e5_6(d) : e <= 8 ? launch_mono_f32_e7_8(d)
       : e <= 10 ? launch_mono_f32_e9_10(d) : e <= 12 ? launch_mono_f32_e11_12(d) : e <= 14 ? launch_mono_f32_e13_14(d)
       : e <= 16 ? launch_mono_f32_e15_16(d) : launch_mono_f32_many(d);
}
cudaError_t launch_biexp_f32_e1_8(const LaunchDesc& d);
cudaError_t launch_biexp_f32_e9_12(const LaunchDesc& d);
cudaError_t launch_biexp_f32_e13_16(const LaunchDesc& d);
cudaError_t launch_biexp_f32_many(const LaunchDesc& d);
inline cudaError_t launch_biexp_f32(const LaunchDesc& d) {
  const int e = d.n_echo;
  return e <= 8 ? launch_biexp_f32_e1_8(d) : e <= 12 ? launch_biexp_f32_e9_12(d) : e <= 16 ? launch_biexp_f32_e13_16(d)
                                                                                           : launch_biexp_f32_many(d);
}

}  // namespace dfit
