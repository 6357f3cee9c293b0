// Shared by every fit kernel: kernel arguments, sample loads, the per-voxel store / fused epilogue and the
// all-gather epilogue, statistics, the TMA / mbarrier primitives, and the launch description the C-ABI
// layer fills.  See fit_kernel.cuh for the map of the kernels.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lm_core.cuh"
#include "mono_fast.cuh"

namespace dfit {


constexpr int kBlock = 128;

enum DType : int { DT_F32 = 0, DT_F64 = 1, DT_I16 = 2, DT_U16 = 3, DT_I32 = 4, DT_U8 = 5 };
enum Layout : int { LAYOUT_PLANAR = 0, LAYOUT_ECHO_FASTEST = 1 };
enum Counter : int { CNT_FITTED = 0, CNT_FAILED, CNT_NONFINITE, CNT_OOB, CNT_ITERS, CNT_MAXITER, CNT_DEFERRED, CNT_COUNT };
constexpr int kStatSlots = 1024;  // power of two
constexpr int kMaxPeers = 8;

// Fused all-gather epilogue (multi-GPU, one process per GPU): while world > 0 the kernels store each voxel's packed
// fp32 row [selected parameters..., r2] straight into the reassembled map of EVERY rank at row `row0 + v` -- peer-mapped
// over NVLink, or with one multicast store that the NVSwitch replicates (NVLS) -- so the collective overlaps the fit
// instead of following it.
struct GatherArgs {
  float* maps[kMaxPeers];  // rank r's map as THIS process addresses it (own allocation for r == self)
  float* mc;               // multicast address of the same buffers (every rank's copy is written by one store) or null
  int world, self;         // world == 0: no gather
  unsigned cols;           // bit i: parameter i is carried in the row (r2 always is, last)
  int ncols;               // floats per row = popcount(cols) + 1
  int64_t row0;            // row of this launch's voxel 0
  int split_list;          // masked fits of ONE volume by all ranks: every rank scans the same whole-volume mask, fills
                           // its own map outside it and fits the masked voxels of its span [fit_lo, fit_hi) only
  int64_t fit_lo, fit_hi;  // this rank's voxel span (the host balances the spans by masked-voxel count)
  int64_t y_voxel0;        // voxel index of the first sample of `y` (a rank holds only its span of samples)
  int mc_weak;             // diagnostic: multicast stores with .weak instead of .relaxed.sys semantics
};

template <typename T, int EMAX>
struct KernelArgs {
  XTab<T, EMAX> xt;
  VoxelOpts<T> vo;
  PostOpts po;
  const void* y;
  int64_t ld;
  int64_t n;
  int y_dtype, layout, E;
  const uint8_t* mask;
  const unsigned* index;        // compacted list of voxels to fit (mask path), or null: fit voxel v = thread id
  const unsigned* index_count;  // device counter: number of entries in `index`
  const void* p0v;  // [N, P] per-voxel initial guess or null
  int p0_dtype;
  unsigned p0_voxel_bits;  // bit i set: parameter i comes from p0v
  T p0s[4];
  void* popt;
  void* r2;
  int out_dtype;
  int sel;  // >= 0: only this parameter is written, popt is [N] (dfit_opts.out_param); -1: all, popt is [N, P]
  uint8_t* status;
  uint8_t* niter;
  double mask_fill;  // value written outside the mask: NaN or nan_to_num (fitting.py:207-212)
  double fill_q[4];  // the same per parameter, after the epilogue's rounding (what a voxel outside the mask reads)
  unsigned long long* counters;
  // LM tail (dense two-voxel TMA kernel): voxels that neither the straight-line fit nor the one-voxel Newton loop
  // settle are appended to this list instead of running the LM inside the kernel; the LM-in-rounds kernel
  // (fit_kernel_lmq) fits the list right after.  Null: the LM runs in place.  `lm_count_next` is the counter of the NEXT
  // launch, which the tail kernel zeroes (two counters alternate, so no launch pays a memset).
  unsigned* lm_list;
  unsigned* lm_count;
  unsigned* lm_count_next;
  GatherArgs g;
};

static inline size_t dtype_size(int dt) {
  switch (dt) {
    case DT_F32: case DT_I32: return 4;
    case DT_F64: return 8;
    case DT_I16: case DT_U16: return 2;
    default: return 1;
  }
}

#if defined(__CUDACC__)

template <typename T>
__device__ __forceinline__ T load_as(const void* __restrict__ base, int dtype, int64_t idx) {
  switch (dtype) {
    case DT_F32: return (T)__ldcs(reinterpret_cast<const float*>(base) + idx);
    case DT_F64: return (T)__ldcs(reinterpret_cast<const double*>(base) + idx);
    case DT_I16: return (T)__ldcs(reinterpret_cast<const short*>(base) + idx);
    case DT_U16: return (T)__ldcs(reinterpret_cast<const unsigned short*>(base) + idx);
    case DT_I32: return (T)__ldcs(reinterpret_cast<const int*>(base) + idx);
    default: return (T)__ldcs(reinterpret_cast<const unsigned char*>(base) + idx);
  }
}

template <typename T, typename S, int EMAX, bool EXACT>
__device__ __forceinline__ void load_strided(const S* __restrict__ src, int64_t stride, int E, T (&y)[EMAX]) {
#pragma unroll
  for (int e = 0; e < EMAX; ++e) y[e] = (EXACT || e < E) ? (T)__ldcs(src + (int64_t)e * stride) : (T)0;
}

template <typename T, int EMAX, bool EXACT>
__device__ __forceinline__ void load_samples(const KernelArgs<T, EMAX>& a, int64_t v, T (&y)[EMAX]) {
  if (a.layout == LAYOUT_ECHO_FASTEST && a.y_dtype == DT_F32 && (EMAX % 4 == 0) && (a.ld % 4 == 0) &&
      (EXACT || a.E == EMAX) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0)) {
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.y) + v * a.ld);
#pragma unroll
    for (int q = 0; q < EMAX / 4; ++q) {
      const float4 t = __ldcs(src + q);
      y[4 * q + 0] = (T)t.x;
      y[4 * q + 1] = (T)t.y;
      y[4 * q + 2] = (T)t.z;
      y[4 * q + 3] = (T)t.w;
    }
    return;
  }
  // Planar: lane l of a warp reads voxel v0 + l of every echo plane -- one fully coalesced line per
  // echo.  The element type is switched once, outside the echo loop.
  const bool planar = a.layout == LAYOUT_PLANAR;
  const int64_t first = planar ? v : v * a.ld, stride = planar ? a.ld : 1;
  switch (a.y_dtype) {
    case DT_F32: load_strided<T, float, EMAX, EXACT>(reinterpret_cast<const float*>(a.y) + first, stride, a.E, y); break;
    case DT_F64: load_strided<T, double, EMAX, EXACT>(reinterpret_cast<const double*>(a.y) + first, stride, a.E, y); break;
    case DT_I16: load_strided<T, short, EMAX, EXACT>(reinterpret_cast<const short*>(a.y) + first, stride, a.E, y); break;
    case DT_U16:
      load_strided<T, unsigned short, EMAX, EXACT>(reinterpret_cast<const unsigned short*>(a.y) + first, stride, a.E, y);
      break;
    case DT_I32: load_strided<T, int, EMAX, EXACT>(reinterpret_cast<const int*>(a.y) + first, stride, a.E, y); break;
    default:
      load_strided<T, unsigned char, EMAX, EXACT>(reinterpret_cast<const unsigned char*>(a.y) + first, stride, a.E, y);
      break;
  }
}

template <int P, typename T, int EMAX>
__device__ __forceinline__ void load_p0(const KernelArgs<T, EMAX>& a, int64_t v, T (&p)[P]) {
#pragma unroll
  for (int i = 0; i < P; ++i) {
    p[i] = a.p0s[i];
    if ((a.p0_voxel_bits >> i) & 1u) p[i] = load_as<T>(a.p0v, a.p0_dtype, v * P + i);
  }
}

template <int P, typename TO>
__device__ __forceinline__ void store_vec(TO* __restrict__ dst, const double (&q)[P]) {
  if constexpr (sizeof(TO) == 4 && P == 2) {
    __stcs(reinterpret_cast<float2*>(dst), make_float2((float)q[0], (float)q[1]));
  } else if constexpr (sizeof(TO) == 4 && P == 4) {
    __stcs(reinterpret_cast<float4*>(dst), make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]));
  } else if constexpr (sizeof(TO) == 8 && P == 2) {
    __stcs(reinterpret_cast<double2*>(dst), make_double2(q[0], q[1]));
  } else if constexpr (sizeof(TO) == 8 && P == 4) {
    __stcs(reinterpret_cast<double2*>(dst), make_double2(q[0], q[1]));
    __stcs(reinterpret_cast<double2*>(dst) + 1, make_double2(q[2], q[3]));
  } else {
#pragma unroll
    for (int i = 0; i < P; ++i) dst[i] = (TO)q[i];
  }
}

// N consecutive floats (N = 1, 2 or 4; the address aligned to N floats) into every rank's map at float offset `off`:
// one multicast store when the maps are bound to an NVLS multicast object, else one store per rank (local HBM for
// the own rank, NVLink for the peers).
template <int N>
__device__ __forceinline__ void gather_store(const GatherArgs& g, int64_t off, const float (&w)[N]) {
  static_assert(N == 1 || N == 2 || N == 4, "vector width");
  if (g.mc != nullptr && g.mc_weak) {
    float* dst = g.mc + off;
    if constexpr (N == 4)
      asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]) : "memory");
    else if constexpr (N == 2)
      asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(w[0]), "f"(w[1]) : "memory");
    else
      asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(dst), "f"(w[0]) : "memory");
    return;
  }
  if (g.mc != nullptr) {
    float* dst = g.mc + off;
    if constexpr (N == 4)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]) : "memory");
    else if constexpr (N == 2)
      asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(w[0]), "f"(w[1]) : "memory");
    else
      asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(w[0]) : "memory");
    return;
  }
  // (each rank starts with its own map and goes round the ring from there: at any moment the ranks address
  // different destinations instead of all converging on the same GPU)
#pragma unroll
  for (int k = 0; k < kMaxPeers; ++k) {
    if (k < g.world) {
      int r = g.self + k;
      if (r >= g.world) r -= g.world;
      float* dst = g.maps[r] + off;
      if constexpr (N == 4) *reinterpret_cast<float4*>(dst) = make_float4(w[0], w[1], w[2], w[3]);
      else if constexpr (N == 2) *reinterpret_cast<float2*>(dst) = make_float2(w[0], w[1]);
      else *dst = w[0];
    }
  }
}

// The row of one voxel: the selected parameters, then r2.  Returns the number of floats.
template <int P>
__device__ __forceinline__ int gather_row(const GatherArgs& g, const float (&q)[P], float r2, float (&row)[P + 1]) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < P; ++i)
    if ((g.cols >> i) & 1u) row[n++] = q[i];
  row[n++] = r2;
  return n;
}

// Epilogue + stores for one voxel.  `fitted` false: voxel outside the mask.
template <int P, typename T, int EMAX, bool GATHER = true>
__device__ __forceinline__ void store_voxel(const KernelArgs<T, EMAX>& a, int64_t v, const T (&p)[P], T r2, bool fitted,
                                            int st, int iters, bool warp_rows = false) {
  if constexpr (sizeof(T) == 4) {
    // fp32 parameters into fp32 maps: raw (curve_fit without an epilogue) or through the fp32-where-exact
    // epilogue -- no trip through double for r2 and the comparisons-only parameters.
    if (fitted && a.out_dtype == DT_F32 && a.popt != nullptr && !(GATHER && a.g.world > 0)) {
      float q[P];
#pragma unroll
      for (int i = 0; i < P; ++i) q[i] = post_param_f32(a.po, i, p[i], r2);
      float* dst = reinterpret_cast<float*>(a.popt) + v * P;
      if (a.sel >= 0) {  // one selected parameter: popt is [N]
        float qs = q[0];
#pragma unroll
        for (int i = 1; i < P; ++i)
          if (i == a.sel) qs = q[i];
        __stcs(reinterpret_cast<float*>(a.popt) + v, qs);
      } else if constexpr (P == 2) __stcs(reinterpret_cast<float2*>(dst), make_float2(q[0], q[1]));
      else if constexpr (P == 4) __stcs(reinterpret_cast<float4*>(dst), make_float4(q[0], q[1], q[2], q[3]));
      else {
#pragma unroll
        for (int i = 0; i < P; ++i) __stcs(dst + i, q[i]);
      }
      __stcs(reinterpret_cast<float*>(a.r2) + v, r2);
      if (a.status) a.status[v] = (uint8_t)st;
      if (a.niter) a.niter[v] = (uint8_t)(iters > 255 ? 255 : iters);
      return;
    }
  }
  double q[P];
  double r2o;
  if (fitted) {
    r2o = (double)r2;
#pragma unroll
    for (int i = 0; i < P; ++i) q[i] = post_param(a.po, i, (double)p[i], r2o);
  } else {
    r2o = a.mask_fill;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      q[i] = a.mask_fill;
      if (a.po.enabled && a.po.decimals[i] >= 0) q[i] = div_pow10(rint(q[i] * a.po.scale[i]), a.po.scale[i], a.po.inv_scale[i]);
    }
  }
  if (a.popt != nullptr) {
    if (a.sel >= 0) {  // one selected parameter: popt is [N]
      double qs = q[0];
#pragma unroll
      for (int i = 1; i < P; ++i)
        if (i == a.sel) qs = q[i];
      if (a.out_dtype == DT_F32) {
        __stcs(reinterpret_cast<float*>(a.popt) + v, (float)qs);
        __stcs(reinterpret_cast<float*>(a.r2) + v, (float)r2o);
      } else {
        __stcs(reinterpret_cast<double*>(a.popt) + v, qs);
        __stcs(reinterpret_cast<double*>(a.r2) + v, r2o);
      }
    } else if (a.out_dtype == DT_F32) {
      store_vec<P, float>(reinterpret_cast<float*>(a.popt) + v * P, q);
      __stcs(reinterpret_cast<float*>(a.r2) + v, (float)r2o);
    } else {
      store_vec<P, double>(reinterpret_cast<double*>(a.popt) + v * P, q);
      __stcs(reinterpret_cast<double*>(a.r2) + v, r2o);
    }
  }
  if (GATHER && a.g.world > 0) {
    constexpr int C = P + 1;
    float qf[P], row[C];
#pragma unroll
    for (int i = 0; i < P; ++i) qf[i] = (float)q[i];
    if (!fitted && a.g.split_list) {
      // split-list mode, outside the mask: the fill value goes into the OWN map only -- every rank compacts the same
      // whole-volume mask and fills its own copy, so nothing of it crosses NVLink
      const int n = gather_row<P>(a.g, qf, (float)r2o, row);
      float* dst = a.g.maps[a.g.self] + (a.g.row0 + v) * n;
      for (int i = 0; i < n; ++i) dst[i] = row[i];
    } else if (warp_rows && a.g.ncols == C && a.g.mc == nullptr) {
      // All columns, whole warp: the warp's 32 rows are one contiguous block of 32*C floats in every map.  Transpose it
      // through shuffles so that each of the C store instructions writes 128 contiguous bytes per warp: NVLink
      // carries full write packets instead of 4-byte fragments at a 4*C-byte stride.
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int i = 0; i < P; ++i) row[i] = qf[i];
      row[P] = (float)r2o;
      float word[C];
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const int w = k * 32 + lane;  // word of the block this lane stores in round k
        const int src = w / C, col = w - src * C;
        float val = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = __shfl_sync(0xffffffffu, row[c], src);
          if (c == col) val = t;
        }
        word[k] = val;
      }
      const int64_t block0 = (a.g.row0 + (v - lane)) * C;
#pragma unroll
      for (int k2 = 0; k2 < kMaxPeers; ++k2) {
        if (k2 < a.g.world) {
          int r = a.g.self + k2;  // (own map first, then round the ring: see gather_store)
          if (r >= a.g.world) r -= a.g.world;
          float* dst = a.g.maps[r] + block0 + lane;  // local HBM for r == own rank, a peer's over NVLink otherwise
#pragma unroll
          for (int k = 0; k < C; ++k) dst[k * 32] = word[k];
        }
      }
    } else {
      const int n = gather_row<P>(a.g, qf, (float)r2o, row);
      const int64_t off = (a.g.row0 + v) * n;
      if (n == 2) {  // one parameter + r2 (the T2 map): 8-byte rows, a warp's rows are contiguous
        const float w2[2] = {row[0], row[1]};
        gather_store<2>(a.g, off, w2);
      } else {
        for (int i = 0; i < n; ++i) {
          const float w1[1] = {row[i]};
          gather_store<1>(a.g, off + i, w1);
        }
      }
    }
  }
  if (a.status) a.status[v] = (uint8_t)st;
  if (a.niter) a.niter[v] = (uint8_t)(iters > 255 ? 255 : iters);
}

// A voxel outside the mask: the fill value into every output (fitting.py:205-215).  The single-GPU case is a few
// streaming stores; with a fused all-gather the general store_voxel decides where the row goes.
template <int P, typename T, int EMAX>
__device__ __forceinline__ void fill_voxel(const KernelArgs<T, EMAX>& a, int64_t v) {
  if (a.g.world > 0) {
    T p[P];
    store_voxel<P, T, EMAX>(a, v, p, (T)0, false, ST_SKIPPED, 0);
    return;
  }
  if (a.popt != nullptr) {
    if (a.out_dtype == DT_F32) {
      if (a.sel >= 0) {
        float qs = (float)a.fill_q[0];
#pragma unroll
        for (int i = 1; i < P; ++i)
          if (i == a.sel) qs = (float)a.fill_q[i];
        __stcs(reinterpret_cast<float*>(a.popt) + v, qs);
      } else {
        double q[P];
#pragma unroll
        for (int i = 0; i < P; ++i) q[i] = a.fill_q[i];
        store_vec<P, float>(reinterpret_cast<float*>(a.popt) + v * P, q);
      }
      __stcs(reinterpret_cast<float*>(a.r2) + v, (float)a.mask_fill);
    } else {
      if (a.sel >= 0) {
        double qs = a.fill_q[0];
#pragma unroll
        for (int i = 1; i < P; ++i)
          if (i == a.sel) qs = a.fill_q[i];
        __stcs(reinterpret_cast<double*>(a.popt) + v, qs);
      } else {
        double q[P];
#pragma unroll
        for (int i = 0; i < P; ++i) q[i] = a.fill_q[i];
        store_vec<P, double>(reinterpret_cast<double*>(a.popt) + v * P, q);
      }
      __stcs(reinterpret_cast<double*>(a.r2) + v, a.mask_fill);
    }
  }
  if (a.status) a.status[v] = (uint8_t)ST_SKIPPED;
  if (a.niter) a.niter[v] = 0;
}

// Statistics are reduced per warp (dense path) or per CTA (grid-stride paths) and added to one of
// kStatSlots slots of global counters (same-address atomics serialise in the L2 atomic unit; 1.8 M warps
// hammering six addresses cost more than the fit itself).  The host sums the slots in dfit_get_stats.
// Dense path: three warp reductions and two fire-and-forget global reductions
// per warp (no shared memory, no block barrier), spread over kStatSlots slots; the rare events (failures,
// non-finite or out-of-bounds voxels) take a separate branch.  Must be reached by all 32 lanes.
__device__ __forceinline__ void warp_stats(unsigned long long* cnt, int st, int iters, unsigned flags) {
  const unsigned full = 0xffffffffu;
  const unsigned fitted = __popc(__ballot_sync(full, st >= ST_CONV_F));
  const unsigned its = __reduce_add_sync(full, (unsigned)iters);
  const unsigned mx = __reduce_max_sync(full, (unsigned)iters);
  const unsigned rare = __ballot_sync(full, st >= ST_MAXITER || flags != 0u);
  if ((threadIdx.x & 31) == 0) {
    unsigned long long* dst =
        cnt + (size_t)((blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5)) & (kStatSlots - 1)) * CNT_COUNT;
    if (fitted) atomicAdd(dst + CNT_FITTED, (unsigned long long)fitted);
    if (its) atomicAdd(dst + CNT_ITERS, (unsigned long long)its);
    if (mx) atomicMax(dst + CNT_MAXITER, (unsigned long long)mx);
  }
  if (rare) {
    const unsigned nfail = __popc(__ballot_sync(full, st >= ST_MAXITER));
    const unsigned nnf = __popc(__ballot_sync(full, (flags & FLAG_NONFINITE) != 0u));
    const unsigned noob = __popc(__ballot_sync(full, (flags & FLAG_OOB) != 0u));
    if ((threadIdx.x & 31) == 0) {
      unsigned long long* dst =
          cnt + (size_t)((blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5)) & (kStatSlots - 1)) * CNT_COUNT;
      if (nfail) atomicAdd(dst + CNT_FAILED, (unsigned long long)nfail);
      if (nnf) atomicAdd(dst + CNT_NONFINITE, (unsigned long long)nnf);
      if (noob) atomicAdd(dst + CNT_OOB, (unsigned long long)noob);
    }
  }
}

// Grid-stride / persistent kernels: per-thread counts -> one reduction per CTA -> one of the counter slots.
// Must be reached by every thread of the CTA.
__device__ __forceinline__ void block_stats_counts(unsigned long long* cnt, unsigned n_fit, unsigned n_fail, unsigned n_nf,
                                                   unsigned n_oob, int it_sum, int it_max) {
  __shared__ unsigned s_c[4], s_iters, s_max;
  if (threadIdx.x < 4) s_c[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    s_iters = 0;
    s_max = 0;
  }
  __syncthreads();
  const unsigned full = 0xffffffffu;
  const unsigned c[4] = {__reduce_add_sync(full, n_fit), __reduce_add_sync(full, n_fail), __reduce_add_sync(full, n_nf),
                         __reduce_add_sync(full, n_oob)};
  const unsigned s_it = __reduce_add_sync(full, (unsigned)it_sum);
  const unsigned m_it = __reduce_max_sync(full, (unsigned)it_max);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c[k]) atomicAdd(&s_c[k], c[k]);
    atomicAdd(&s_iters, s_it);
    atomicMax(&s_max, m_it);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* dst = cnt + (size_t)(blockIdx.x & (kStatSlots - 1)) * CNT_COUNT;
    if (s_c[0]) atomicAdd(dst + CNT_FITTED, (unsigned long long)s_c[0]);
    if (s_c[1]) atomicAdd(dst + CNT_FAILED, (unsigned long long)s_c[1]);
    if (s_c[2]) atomicAdd(dst + CNT_NONFINITE, (unsigned long long)s_c[2]);
    if (s_c[3]) atomicAdd(dst + CNT_OOB, (unsigned long long)s_c[3]);
    if (s_iters) atomicAdd(dst + CNT_ITERS, (unsigned long long)s_iters);
    if (s_max) atomicMax(dst + CNT_MAXITER, (unsigned long long)s_max);
  }
}

// ---- TMA / mbarrier primitives --------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

#endif  // __CUDACC__

// Type-erased launch description filled by the C-ABI layer and consumed by the per-model
// translation units (inst_*.cu).
struct LaunchDesc {
  int model, compute_dtype, n_echo;
  int64_t n_vox;
  const double* x;  // host
  const void* y;
  int y_dtype, layout;
  int64_t ld;
  const uint8_t* mask;
  unsigned* index;        // device scratch for the compacted mask path (n_vox entries) or null
  unsigned* index_count;  // device counter
  const void* p0v;
  int p0_dtype;
  unsigned p0_voxel_bits;
  double p0s[4];
  void* popt;
  void* r2;
  int out_dtype;
  int sel;
  uint8_t* status;
  uint8_t* niter;
  unsigned long long* counters;
  // solver
  double ftol, xtol, lambda0, floor_rel, r2_eps, y_lo, y_hi;
  int maxfev, init_mode, init_linear, fast_path;
  PostOpts po;
  double mask_fill;
  int use_tma;
  cudaStream_t stream;
  GatherArgs g;
  const CUtensorMap* tmap;   // host pointer to an encoded 2-D map of the planar fp32 samples (box 32 x E), or null
  const CUtensorMap* tmap2;  // the same with a 64-voxel box, for the two-voxels-per-lane kernel, or null
  int sm_count;
  unsigned* lm_list;  // LM tail (see KernelArgs): device list of n_vox entries, or null: the LM runs inside the kernel
  unsigned* lm_head;  // device, two alternating counters
  int* lm_parity;     // host: which of the two the next launch counts in (flipped by launch_one when a tail was launched)
};

template <typename T, int EMAX>
inline void fill_args(const LaunchDesc& d, KernelArgs<T, EMAX>& a) {
  fill_xtab<T, EMAX>(a.xt, d.x, d.n_echo);
  a.vo.s.ftol = (T)d.ftol;
  a.vo.s.xtol = (T)d.xtol;
  a.vo.s.lambda0 = (T)d.lambda0;
  a.vo.s.floor_rel = (T)d.floor_rel;
  a.vo.s.maxfev = d.maxfev;
  a.vo.s.init_linear = d.init_linear;
  a.vo.y_lo = (T)d.y_lo;
  a.vo.y_hi = (T)d.y_hi;
  a.vo.r2_eps = (T)d.r2_eps;
  a.vo.init_mode = d.init_mode;
  a.vo.has_bounds = (d.y_lo > -1.7e308 || d.y_hi < 1.7e308) ? 1 : 0;
  a.vo.fast = d.fast_path;
  a.po = d.po;
  a.y = d.y;
  a.ld = d.ld;
  a.n = d.n_vox;
  a.y_dtype = d.y_dtype;
  a.layout = d.layout;
  a.E = d.n_echo;
  a.mask = d.mask;
  a.index = nullptr;
  a.index_count = nullptr;
  a.p0v = d.p0v;
  a.p0_dtype = d.p0_dtype;
  a.p0_voxel_bits = d.p0_voxel_bits;
  for (int i = 0; i < 4; ++i) a.p0s[i] = (T)d.p0s[i];
  a.popt = d.popt;
  a.r2 = d.r2;
  a.out_dtype = d.out_dtype;
  a.sel = d.sel;
  a.status = d.status;
  a.niter = d.niter;
  a.mask_fill = d.mask_fill;
  for (int i = 0; i < 4; ++i) {
    double q = d.mask_fill;
    if (d.po.enabled && d.po.decimals[i] >= 0) q = nearbyint(q * d.po.scale[i]) / d.po.scale[i];  // np.around of the fill
    a.fill_q[i] = q;
  }
  a.counters = d.counters;
  a.lm_list = nullptr;  // (set by launch_one for the kernel that fills it)
  a.lm_count = nullptr;
  a.lm_count_next = nullptr;
  a.g = d.g;
}

}  // namespace dfit
