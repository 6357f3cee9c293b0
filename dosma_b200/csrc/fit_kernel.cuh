// Fit kernels: one voxel per lane, samples and the whole LM state in registers.
//
//   fit_kernel      -- coalesced/vectorised global loads straight into registers
//   fit_kernel_tma  -- persistent CTAs, sample tiles staged through shared memory by TMA
//                      (cp.async.bulk.tensor) in a double-buffered mbarrier pipeline
//
// Both replace the N-voxel loop of dosma/core/fitting.py:855-868 and fuse what the reference does
// around it: dtype up-cast (:711, SciPy's asarray(float)), mask select/scatter (:199-215), the
// log-linear initial guess (:701-718), `_process_params` (:109-146) and rounding (:734-737).
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
enum Counter : int { CNT_FITTED = 0, CNT_FAILED, CNT_NONFINITE, CNT_OOB, CNT_ITERS, CNT_MAXITER, CNT_COUNT };
constexpr int kStatSlots = 1024;  // power of two
constexpr int kMaxPeers = 8;

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
  uint8_t* status;
  uint8_t* niter;
  double mask_fill;  // value written outside the mask: NaN or nan_to_num (fitting.py:207-212)
  unsigned long long* counters;
  // Fused all-gather epilogue: when gather_world > 0 every voxel's packed row [popt..., r2] (fp32) is
  // also stored straight into the reassembled map of EVERY rank (peer-mapped over NVLink) at row
  // gather_row0 + v -- the collective overlaps the fit instead of following it.
  float* gather[kMaxPeers];
  int gather_world;
  int64_t gather_row0;
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

// Epilogue + stores for one voxel.  `fitted` false: voxel outside the mask.
template <int P, typename T, int EMAX, bool GATHER = true>
__device__ __forceinline__ void store_voxel(const KernelArgs<T, EMAX>& a, int64_t v, const T (&p)[P], T r2, bool fitted,
                                            int st, int iters, bool warp_rows = false) {
  if constexpr (sizeof(T) == 4) {
    // fp32 parameters into fp32 maps: raw (curve_fit without an epilogue) or through the fp32-where-exact
    // epilogue -- no trip through double for r2 and the comparisons-only parameters.
    if (fitted && a.out_dtype == DT_F32 && a.popt != nullptr && !(GATHER && a.gather_world > 0)) {
      float q[P];
#pragma unroll
      for (int i = 0; i < P; ++i) q[i] = post_param_f32(a.po, i, p[i], r2);
      float* dst = reinterpret_cast<float*>(a.popt) + v * P;
      if constexpr (P == 2) __stcs(reinterpret_cast<float2*>(dst), make_float2(q[0], q[1]));
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
    if (a.out_dtype == DT_F32) {
      store_vec<P, float>(reinterpret_cast<float*>(a.popt) + v * P, q);
      __stcs(reinterpret_cast<float*>(a.r2) + v, (float)r2o);
    } else {
      store_vec<P, double>(reinterpret_cast<double*>(a.popt) + v * P, q);
      __stcs(reinterpret_cast<double*>(a.r2) + v, r2o);
    }
  }
  if (GATHER && a.gather_world > 0) {
    constexpr int C = P + 1;
    float row[C];
#pragma unroll
    for (int i = 0; i < P; ++i) row[i] = (float)q[i];
    row[P] = (float)r2o;
    if (warp_rows) {
      // The warp's 32 rows are one contiguous block of 32*C floats in every map.  Transpose it through
      // shuffles so that each of the C store instructions writes 128 contiguous bytes per warp: NVLink
      // carries full write packets instead of 4-byte fragments at a 4*C-byte stride.
      const int lane = threadIdx.x & 31;
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
      const int64_t block0 = (a.gather_row0 + (v - lane)) * C;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) {
        if (r < a.gather_world) {
          float* dst = a.gather[r] + block0 + lane;  // local HBM for r == own rank, a peer's over NVLink otherwise
#pragma unroll
          for (int k = 0; k < C; ++k) dst[k * 32] = word[k];
        }
      }
    } else {
      const int64_t off = (a.gather_row0 + v) * C;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) {
        if (r < a.gather_world) {
          float* dst = a.gather[r] + off;
#pragma unroll
          for (int i = 0; i < C; ++i) dst[i] = row[i];
        }
      }
    }
  }
  if (a.status) a.status[v] = (uint8_t)st;
  if (a.niter) a.niter[v] = (uint8_t)(iters > 255 ? 255 : iters);
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

// Mask path, step 1: one streaming pass over the mask that (a) appends the voxels to fit to a compact
// index list -- each warp claims a contiguous run with one atomic, so neighbours stay neighbours -- and
// (b) writes the fill value for every voxel outside the mask (fitting.py:205-215).  Step 2 is the fit
// kernel over the list: all 32 lanes of a warp fit, however thin the tissue mask is.
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

constexpr int kCompactPerThread = 8;  // voxels per thread in the compaction pass (2048 per CTA)

template <int P, typename T, int EMAX>
__global__ void __launch_bounds__(256) mask_compact_kernel(const __grid_constant__ KernelArgs<T, EMAX> a, unsigned* index,
                                                           unsigned* count) {
  __shared__ unsigned s_list[256 * kCompactPerThread];
  __shared__ unsigned s_n, s_base;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int64_t v0 = (int64_t)blockIdx.x * (256 * kCompactPerThread);
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kCompactPerThread; ++k) {
    const int64_t v = v0 + k * 256 + threadIdx.x;
    const bool in = v < a.n;
    const bool active = in && a.mask[v] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, active);
    unsigned base = 0;
    if (lane == 0 && ballot) base = atomicAdd(&s_n, (unsigned)__popc(ballot));  // shared-memory atomic
    base = __shfl_sync(0xffffffffu, base, 0);
    if (active) s_list[base + __popc(ballot & ((1u << lane) - 1u))] = (unsigned)v;
    if (in && !active) {
      T p[P];
      store_voxel<P, T, EMAX>(a, v, p, (T)0, false, ST_SKIPPED, 0);
    }
  }
  __syncthreads();
  const unsigned n = s_n;
  if (threadIdx.x == 0 && n) s_base = atomicAdd(count, n);  // ONE global atomic per 2048 voxels
  __syncthreads();
  for (unsigned i = threadIdx.x; i < n; i += 256) index[s_base + i] = s_list[i];
}

template <class M, typename T, int EMAX, bool EXACT, bool GATHER>
__global__ void __launch_bounds__(kBlock, sizeof(T) == 4 ? (EMAX <= 8 ? 8 : 3) : (EMAX <= 8 ? 4 : 1)) fit_kernel(const __grid_constant__ KernelArgs<T, EMAX> a) {
  constexpr int P = M::P;
  int st = -1, iters = 0;
  unsigned flags = 0;
  if (a.index == nullptr) {
    const int64_t v = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool whole_warp = (v | 31) < a.n;  // all 32 voxels of this warp exist: cooperative stores are legal
    if (v < a.n) {
      T p[P], r2 = 0, y[EMAX];
      if (a.layout == LAYOUT_PLANAR && a.y_dtype == DT_F32) {
        // the common case, kept free of the dtype / layout dispatch: the plane bases are CTA-uniform
        const float* __restrict__ base = reinterpret_cast<const float*>(a.y) + (int64_t)blockIdx.x * kBlock;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) y[e] = (EXACT || e < a.E) ? (T)__ldcs(base + (int64_t)e * a.ld + threadIdx.x) : (T)0;
      } else {
        load_samples<T, EMAX, EXACT>(a, v, y);
      }
      st = fit_voxel_fast<M, T, EMAX, EXACT>(y, a.xt, a.vo, p, r2, iters);
      if (st < 0) {  // the general path: LM from the caller's initial guess
        load_p0<P, T, EMAX>(a, v, p);
        st = fit_voxel<M, T, T, EMAX, EXACT>(y, a.xt, a.E, a.vo, p, r2, iters, flags);
      }
      if (GATHER && whole_warp) __syncwarp();
      store_voxel<P, T, EMAX, GATHER>(a, v, p, r2, true, st, iters, whole_warp);
    }
    __syncwarp();
    warp_stats(a.counters, st, iters, flags);
    return;
  } else {
    // compacted mask path: grid-stride over the index list (its length is only known on the device)
    const unsigned count = *a.index_count;
    int it_sum = 0;
    unsigned n_fit = 0, n_fail = 0, n_nf = 0, n_oob = 0;
    for (unsigned i = blockIdx.x * kBlock + threadIdx.x; i < count; i += gridDim.x * kBlock) {
      const int64_t v = (int64_t)a.index[i];
      T p[P], r2 = 0, y[EMAX];
      int it = 0;
      unsigned fl = 0;
      load_samples<T, EMAX, EXACT>(a, v, y);
      int s = fit_voxel_fast<M, T, EMAX, EXACT>(y, a.xt, a.vo, p, r2, it);
      if (s < 0) {
        load_p0<P, T, EMAX>(a, v, p);
        s = fit_voxel<M, T, T, EMAX, EXACT>(y, a.xt, a.E, a.vo, p, r2, it, fl);
      }
      store_voxel<P, T, EMAX, GATHER>(a, v, p, r2, true, s, it);
      it_sum += it;
      iters = it > iters ? it : iters;
      n_fit += (unsigned)(s >= ST_CONV_F);
      n_fail += (unsigned)(s >= ST_MAXITER);
      n_nf += (unsigned)((fl & FLAG_NONFINITE) != 0);
      n_oob += (unsigned)((fl & FLAG_OOB) != 0);
    }
    block_stats_counts(a.counters, n_fit, n_fail, n_nf, n_oob, it_sum, iters);
  }
}

// ------------------------------------------------------------------------------------------------
// Two voxels per lane: the dense mono-exponential fast path (uniform echo spacing, fp32 arithmetic, planar
// f32 / i16 / u16 samples).  Lane l of a CTA owns voxels 2 (128 b + l) and the next one: one 8-byte (4-byte
// for 16-bit samples) coalesced load per echo, every packed instruction works on both voxels, and the
// results leave as one 16-byte [a, b, a, b] store and one 8-byte r2 store.  Voxels the fast path declines
// run the general LM from the caller's initial guess, one at a time (rare).
constexpr int kBlock2 = 128;

template <typename S>
struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<short> { typedef short2 type; };
template <> struct Vec2<unsigned short> { typedef ushort2 type; };

template <typename S, int EMAX>
__device__ __forceinline__ void load_pairs(const void* __restrict__ yv, int64_t ld, int64_t v0, bool both,
                                           pair2<float> (&Y)[EMAX]) {
  const S* __restrict__ base = reinterpret_cast<const S*>(yv) + v0;
  if (both) {
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      const typename Vec2<S>::type t = __ldcs(reinterpret_cast<const typename Vec2<S>::type*>(base + (int64_t)e * ld));
      Y[e] = p2_make<float>((float)t.x, (float)t.y);
    }
  } else {  // odd tail: the missing voxel duplicates the last one and is never stored
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      const float t = (float)__ldcs(base + (int64_t)e * ld);
      Y[e] = p2_make<float>(t, t);
    }
  }
}

__device__ __forceinline__ void warp_stats2(unsigned long long* cnt, const int (&st)[2], const int (&iters)[2],
                                            unsigned flags) {
  const unsigned full = 0xffffffffu;
  const unsigned fitted = __popc(__ballot_sync(full, st[0] >= ST_CONV_F)) + __popc(__ballot_sync(full, st[1] >= ST_CONV_F));
  const unsigned its = __reduce_add_sync(full, (unsigned)(iters[0] + iters[1]));
  const unsigned mx = __reduce_max_sync(full, (unsigned)(iters[0] > iters[1] ? iters[0] : iters[1]));
  const unsigned rare = __ballot_sync(full, st[0] >= ST_MAXITER || st[1] >= ST_MAXITER || flags != 0u);
  const unsigned slot = (blockIdx.x * (kBlock2 / 32) + (threadIdx.x >> 5)) & (kStatSlots - 1);
  unsigned long long* dst = cnt + (size_t)slot * CNT_COUNT;
  if ((threadIdx.x & 31) == 0) {
    if (fitted) atomicAdd(dst + CNT_FITTED, (unsigned long long)fitted);
    if (its) atomicAdd(dst + CNT_ITERS, (unsigned long long)its);
    if (mx) atomicMax(dst + CNT_MAXITER, (unsigned long long)mx);
  }
  if (rare) {
    const unsigned nfail = __reduce_add_sync(full, (unsigned)(st[0] >= ST_MAXITER) + (unsigned)(st[1] >= ST_MAXITER));
    const unsigned nnf = __reduce_add_sync(full, (flags & 0xffu));
    const unsigned noob = __reduce_add_sync(full, (flags >> 8) & 0xffu);
    if ((threadIdx.x & 31) == 0) {
      if (nfail) atomicAdd(dst + CNT_FAILED, (unsigned long long)nfail);
      if (nnf) atomicAdd(dst + CNT_NONFINITE, (unsigned long long)nnf);
      if (noob) atomicAdd(dst + CNT_OOB, (unsigned long long)noob);
    }
  }
}

template <class M, int EMAX>
__global__ void __launch_bounds__(kBlock2, 5) fit_kernel_mono2(const __grid_constant__ KernelArgs<float, EMAX> a) {
  typedef float T;
  constexpr int P = 2;
  const int64_t v0 = ((int64_t)blockIdx.x * kBlock2 + threadIdx.x) * 2;
  int st[2] = {-1, -1}, iters[2] = {0, 0};
  unsigned nflags = 0;  // bits 0..7: non-finite voxels of this lane, bits 8..15: out-of-bounds voxels
  if (v0 < a.n) {
    const bool both = v0 + 1 < a.n;
    pair2<T> Y[EMAX], pa, pb, r2;
    if (a.y_dtype == DT_F32) load_pairs<float, EMAX>(a.y, a.ld, v0, both, Y);
    else if (a.y_dtype == DT_I16) load_pairs<short, EMAX>(a.y, a.ld, v0, both, Y);
    else load_pairs<unsigned short, EMAX>(a.y, a.ld, v0, both, Y);
    fit_voxel_fast2<M, T, EMAX, pair2<T>[EMAX]>(Y, a.xt, a.vo, pa, pb, r2, st, iters);
    if (st[0] < 0 || st[1] < 0) {  // the general path, one voxel at a time
#pragma unroll 1
      for (int hsel = 0; hsel < 2; ++hsel) {
        if ((hsel ? st[1] : st[0]) >= 0) continue;
        T ys[EMAX], p[P], r = 0;
        int it = 0;
        unsigned fl = 0;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) ys[e] = hsel ? Y[e].hi : Y[e].lo;
        load_p0<P, T, EMAX>(a, hsel && both ? v0 + 1 : v0, p);
        const int s1 = fit_voxel<M, T, T, EMAX, true>(ys, a.xt, a.E, a.vo, p, r, it, fl);
        if (hsel == 0 || both) nflags += ((fl & FLAG_NONFINITE) ? 1u : 0u) + ((fl & FLAG_OOB) ? 0x100u : 0u);
        if (hsel) {
          st[1] = s1; iters[1] = it; pa.hi = p[0]; pb.hi = p[1]; r2.hi = r;
        } else {
          st[0] = s1; iters[0] = it; pa.lo = p[0]; pb.lo = p[1]; r2.lo = r;
        }
      }
    }
    if (!both) {
      st[1] = -1;
      iters[1] = 0;
    }
    if (!a.po.enabled && a.out_dtype == DT_F32 && both) {
      __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(a.popt) + v0 * P), make_float4(pa.lo, pb.lo, pa.hi, pb.hi));
      __stcs(reinterpret_cast<float2*>(reinterpret_cast<float*>(a.r2) + v0), make_float2(r2.lo, r2.hi));
      if (a.status) {
        a.status[v0] = (uint8_t)st[0];
        a.status[v0 + 1] = (uint8_t)st[1];
      }
      if (a.niter) {
        a.niter[v0] = (uint8_t)iters[0];
        a.niter[v0 + 1] = (uint8_t)iters[1];
      }
    } else {
      const T p0_[P] = {pa.lo, pb.lo}, p1_[P] = {pa.hi, pb.hi};
      store_voxel<P, T, EMAX, false>(a, v0, p0_, r2.lo, true, st[0], iters[0]);
      if (both) store_voxel<P, T, EMAX, false>(a, v0 + 1, p1_, r2.hi, true, st[1], iters[1]);
    }
  }
  __syncwarp();
  warp_stats2(a.counters, st, iters, nflags);
}

// Two voxels per lane over the compacted voxel list of the mask path (any echo spacing, any sample type):
// lane i takes list entries 2i and 2i+1, gathers their samples and runs the same packed fast path; voxels
// it declines run the LM.  Grid-stride, because the list length is only known on the device.
template <class M, int EMAX>
__global__ void __launch_bounds__(kBlock, 5) fit_kernel_mono2_list(const __grid_constant__ KernelArgs<float, EMAX> a) {
  typedef float T;
  constexpr int P = 2;
  const unsigned count = *a.index_count;
  const unsigned npairs = (count + 1u) >> 1;
  int it_sum = 0, it_max = 0;
  unsigned n_fit = 0, n_fail = 0, n_nf = 0, n_oob = 0;
  for (unsigned i = blockIdx.x * kBlock + threadIdx.x; i < npairs; i += gridDim.x * kBlock) {
    const bool both = 2u * i + 1u < count;
    const int64_t vA = (int64_t)a.index[2u * i], vB = both ? (int64_t)a.index[2u * i + 1u] : vA;
    T yA[EMAX], yB[EMAX];
    load_samples<T, EMAX, true>(a, vA, yA);
    load_samples<T, EMAX, true>(a, vB, yB);
    pair2<T> Y[EMAX], pa, pb, r2;
#pragma unroll
    for (int e = 0; e < EMAX; ++e) Y[e] = p2_make<T>(yA[e], yB[e]);
    int st[2], iters[2];
    fit_voxel_fast2<M, T, EMAX, pair2<T>[EMAX]>(Y, a.xt, a.vo, pa, pb, r2, st, iters);
    if (st[0] < 0 || (st[1] < 0 && both)) {
#pragma unroll 1
      for (int hsel = 0; hsel < 2; ++hsel) {
        if ((hsel ? st[1] : st[0]) >= 0 || (hsel && !both)) continue;
        T ys[EMAX], p[P], r = 0;
        int it = 0;
        unsigned fl = 0;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) ys[e] = hsel ? Y[e].hi : Y[e].lo;
        load_p0<P, T, EMAX>(a, hsel ? vB : vA, p);
        const int s1 = fit_voxel<M, T, T, EMAX, true>(ys, a.xt, a.E, a.vo, p, r, it, fl);
        n_nf += (unsigned)((fl & FLAG_NONFINITE) != 0);
        n_oob += (unsigned)((fl & FLAG_OOB) != 0);
        if (hsel) {
          st[1] = s1; iters[1] = it; pa.hi = p[0]; pb.hi = p[1]; r2.hi = r;
        } else {
          st[0] = s1; iters[0] = it; pa.lo = p[0]; pb.lo = p[1]; r2.lo = r;
        }
      }
    }
    const T p0_[P] = {pa.lo, pb.lo}, p1_[P] = {pa.hi, pb.hi};
    store_voxel<P, T, EMAX, false>(a, vA, p0_, r2.lo, true, st[0], iters[0]);
    if (both) store_voxel<P, T, EMAX, false>(a, vB, p1_, r2.hi, true, st[1], iters[1]);
    else { st[1] = -1; iters[1] = 0; }
    it_sum += iters[0] + iters[1];
    it_max = iters[0] > it_max ? iters[0] : it_max;
    it_max = iters[1] > it_max ? iters[1] : it_max;
    n_fit += (unsigned)(st[0] >= ST_CONV_F) + (unsigned)(st[1] >= ST_CONV_F);
    n_fail += (unsigned)(st[0] >= ST_MAXITER) + (unsigned)(st[1] >= ST_MAXITER);
  }
  block_stats_counts(a.counters, n_fit, n_fail, n_nf, n_oob, it_sum, it_max);
}

// ------------------------------------------------------------------------------------------------
// TMA-staged variant.  Persistent warps: every warp owns a 2-stage shared-memory ring of
// [E][32-voxel] sample tiles that the Tensor Memory Accelerator fills (cp.async.bulk.tensor.2d over a
// 2-D tensor map of the planar (E, ld) array, box = 32 voxels x E echoes) while the warp is busy
// fitting the previous tile; completion is signalled on a per-stage mbarrier (complete_tx::bytes).
// There is no block-level synchronisation -- warps run their own pipelines, so a slow voxel only
// holds its own warp.  Used for fp32 planar samples whose row pitch is a multiple of 16 bytes.
constexpr int kTmaWarps = 8;          // warps per CTA
constexpr int tma_stages(int E) { return E <= 8 ? 4 : 2; }  // ring depth: 32 KB of tiles per CTA at 8 echoes
constexpr int kTmaTile = 32;          // voxels per warp tile (one per lane)

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

template <class M, typename T, int EMAX>
__global__ void __launch_bounds__(kTmaWarps * 32)
    fit_kernel_tma(const __grid_constant__ KernelArgs<T, EMAX> a, const __grid_constant__ CUtensorMap tmap) {
  constexpr int P = M::P;
  constexpr unsigned kTileBytes = EMAX * kTmaTile * sizeof(float);
  constexpr int kTmaStages = tma_stages(EMAX);
  __shared__ __align__(128) float tiles[kTmaWarps][kTmaStages][EMAX][kTmaTile];
  __shared__ __align__(8) uint64_t full[kTmaWarps][kTmaStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.n + kTmaTile - 1) / kTmaTile;
  const int64_t warp_global = (int64_t)blockIdx.x * kTmaWarps + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kTmaWarps;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // prologue: fill the ring
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) {
      const int64_t t = warp_global + (int64_t)s * warp_stride;
      if (t < n_tiles) {
        mbar_expect_tx(&full[warp][s], kTileBytes);
        tma_load_2d(&tiles[warp][s][0][0], &tmap, (int)(t * kTmaTile), 0, &full[warp][s]);
      }
    }
  }
  __syncwarp();

  int st_acc_fit = 0, st_acc_fail = 0, st_acc_nf = 0, st_acc_oob = 0, it_sum = 0, it_max = 0;
  int k = 0;
  for (int64_t t = warp_global; t < n_tiles; t += warp_stride, ++k) {
    const int s = k % kTmaStages;
    const unsigned parity = (unsigned)(k / kTmaStages) & 1u;
    mbar_wait(&full[warp][s], parity);
    T y[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) y[e] = (T)tiles[warp][s][e][lane];  // conflict-free: lane == bank
    __syncwarp();
    if (lane == 0) {  // the stage is drained: refill it with the tile two trips ahead
      const int64_t tn = t + (int64_t)kTmaStages * warp_stride;
      if (tn < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[warp][s], kTileBytes);
        tma_load_2d(&tiles[warp][s][0][0], &tmap, (int)(tn * kTmaTile), 0, &full[warp][s]);
      }
    }
    const int64_t v = t * kTmaTile + lane;
    int st = -1, iters = 0;
    unsigned flags = 0;
    if (v < a.n) {
      const bool active = a.mask == nullptr || a.mask[v] != 0;
      T p[P], r2 = 0;
      st = ST_SKIPPED;
      if (active) {
        st = fit_voxel_fast<M, T, EMAX, true>(y, a.xt, a.vo, p, r2, iters);
        if (st < 0) {
          load_p0<P, T, EMAX>(a, v, p);
          st = fit_voxel<M, T, T, EMAX, true>(y, a.xt, a.E, a.vo, p, r2, iters, flags);
        }
      }
      store_voxel<P, T, EMAX>(a, v, p, r2, active, st, iters);
    }
    st_acc_fit += st >= ST_CONV_F;
    st_acc_fail += st >= ST_MAXITER;
    st_acc_nf += (flags & FLAG_NONFINITE) != 0;
    st_acc_oob += (flags & FLAG_OOB) != 0;
    it_sum += iters;
    it_max = iters > it_max ? iters : it_max;
  }
  // statistics: per-thread accumulators -> one reduction per CTA
  {
    __shared__ unsigned s_cnt[CNT_COUNT];
    if (threadIdx.x < CNT_COUNT) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned full_mask = 0xffffffffu;
    const unsigned v0 = __reduce_add_sync(full_mask, (unsigned)st_acc_fit), v1 = __reduce_add_sync(full_mask, (unsigned)st_acc_fail);
    const unsigned v2 = __reduce_add_sync(full_mask, (unsigned)st_acc_nf), v3 = __reduce_add_sync(full_mask, (unsigned)st_acc_oob);
    const unsigned v4 = __reduce_add_sync(full_mask, (unsigned)it_sum), v5 = __reduce_max_sync(full_mask, (unsigned)it_max);
    if (lane == 0) {
      atomicAdd(&s_cnt[CNT_FITTED], v0);
      atomicAdd(&s_cnt[CNT_FAILED], v1);
      atomicAdd(&s_cnt[CNT_NONFINITE], v2);
      atomicAdd(&s_cnt[CNT_OOB], v3);
      atomicAdd(&s_cnt[CNT_ITERS], v4);
      atomicMax(&s_cnt[CNT_MAXITER], v5);
    }
    __syncthreads();
    if (threadIdx.x < CNT_COUNT) {
      const unsigned val = s_cnt[threadIdx.x];
      unsigned long long* dst = a.counters + (size_t)(blockIdx.x & (kStatSlots - 1)) * CNT_COUNT + threadIdx.x;
      if (val) {
        if (threadIdx.x == CNT_MAXITER) atomicMax(dst, (unsigned long long)val);
        else atomicAdd(dst, (unsigned long long)val);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Two voxels per lane + TMA staging (fp32 or raw 16-bit samples, converted on the way to registers):
// persistent warps, each with its own ring of [E][64-voxel] sample tiles in shared memory.  The Tensor Memory Accelerator fills a stage (one cp.async.bulk.tensor.2d over
// the 2-D map of the planar samples, box = 64 voxels x E echoes, completion on the stage's mbarrier) while
// the warp is fitting earlier tiles, so the HBM latency that the plain kernel exposes at the top of every
// CTA is hidden behind arithmetic.  Lanes read their two voxels of every echo as one conflict-free 8-byte
// shared load.  No block-level synchronisation inside the loop.
constexpr int kM2Warps = 4;
constexpr int kM2Tile = 64;
constexpr int m2_stages(int E) { return E <= 8 ? 4 : 2; }  // 32 KB of tiles per CTA

template <class M, int EMAX, bool GATHER, typename S>
__global__ void __launch_bounds__(kM2Warps * 32, 5)  // 5 CTAs/SM: 6 spills, 4 is slower (measured: 0.743 / 0.689 / 0.696 ms);
                                                     // reading the samples from the tile on every use to free 16 registers
                                                     // (6-7 CTAs/SM) was measured too: 3 % slower
    fit_kernel_mono2_tma(const __grid_constant__ KernelArgs<float, EMAX> a, const __grid_constant__ CUtensorMap tmap) {
  typedef float T;
  constexpr int P = 2;
  constexpr int kStages = m2_stages(EMAX);
  constexpr unsigned kTileBytes = EMAX * kM2Tile * sizeof(S);  // S = float, or the raw 16-bit DICOM sample type
  __shared__ __align__(128) S tiles[kM2Warps][kStages][EMAX][kM2Tile];
  __shared__ __align__(8) uint64_t full[kM2Warps][kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 32-bit indexing: the launcher admits fewer than 2^31 voxels
  const int n_vox = (int)a.n;
  const int n_tiles = (n_vox + kM2Tile - 1) / kM2Tile;
  const int warp_global = (int)blockIdx.x * kM2Warps + warp;
  const int warp_stride = (int)gridDim.x * kM2Warps;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int s = 0; s < kStages; ++s) {  // prologue: fill the ring
      const int t = warp_global + s * warp_stride;
      if (t < n_tiles) {
        mbar_expect_tx(&full[warp][s], kTileBytes);
        tma_load_2d(&tiles[warp][s][0][0], &tmap, t * kM2Tile, 0, &full[warp][s]);
      }
    }
  }
  __syncwarp();

  unsigned n_fit = 0, it_sum = 0, it_max = 0;
  unsigned long long* const stat_slot = a.counters + (size_t)(warp_global & (kStatSlots - 1)) * CNT_COUNT;
  int k = 0;
  for (int t = warp_global; t < n_tiles; t += warp_stride, ++k) {
    const int s = k % kStages;
    mbar_wait(&full[warp][s], (unsigned)(k / kStages) & 1u);
    pair2<T> Y[EMAX], pa, pb, r2;
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      const typename Vec2<S>::type v = *reinterpret_cast<const typename Vec2<S>::type*>(&tiles[warp][s][e][2 * lane]);
      Y[e] = p2_make<T>((T)v.x, (T)v.y);
    }
    __syncwarp();
    if (lane == 0) {  // the stage is drained: refill it with the tile kStages trips ahead
      const int tn = t + kStages * warp_stride;
      if (tn < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[warp][s], kTileBytes);
        tma_load_2d(&tiles[warp][s][0][0], &tmap, tn * kM2Tile, 0, &full[warp][s]);
      }
    }
    const int v0 = t * kM2Tile + 2 * lane;
    const bool validA = v0 < n_vox, validB = v0 + 1 < n_vox;
    int st[2], iters[2];
    fit_voxel_fast2<M, T, EMAX, pair2<T>[EMAX]>(Y, a.xt, a.vo, pa, pb, r2, st, iters);  // voxels past the end are zero-filled: declined
    if ((st[0] < 0 && validA) || (st[1] < 0 && validB)) {  // the general path, one voxel at a time
#pragma unroll 1
      for (int hsel = 0; hsel < 2; ++hsel) {
        if ((hsel ? st[1] : st[0]) >= 0 || !(hsel ? validB : validA)) continue;
        T ys[EMAX], p[P], r = 0;
        int it = 0;
        unsigned fl = 0;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) ys[e] = hsel ? Y[e].hi : Y[e].lo;
        load_p0<P, T, EMAX>(a, v0 + hsel, p);
        const int s1 = fit_voxel<M, T, T, EMAX, true>(ys, a.xt, a.E, a.vo, p, r, it, fl);
        // rare events go straight to the counters (the voxel has just paid for a full LM anyway)
        if (s1 >= ST_MAXITER) atomicAdd(stat_slot + CNT_FAILED, 1ull);
        if (fl & FLAG_NONFINITE) atomicAdd(stat_slot + CNT_NONFINITE, 1ull);
        if (fl & FLAG_OOB) atomicAdd(stat_slot + CNT_OOB, 1ull);
        if (hsel) {
          st[1] = s1; iters[1] = it; pa.hi = p[0]; pb.hi = p[1]; r2.hi = r;
        } else {
          st[0] = s1; iters[0] = it; pa.lo = p[0]; pb.lo = p[1]; r2.lo = r;
        }
      }
    }
    if (!validA) { st[0] = -1; iters[0] = 0; }
    if (!validB) { st[1] = -1; iters[1] = 0; }
    if constexpr (GATHER) {
      // Fused all-gather: the tile's 64 rows [a, b, r2] are one contiguous 768-byte block in every rank's map.
      // Stage them in shared memory (double-buffered) and let the TMA push the block to every rank with one
      // bulk store each (cp.async.bulk global <- shared): the SM's load/store path never waits on NVLink.  The
      // launcher admits this kernel only without the epilogue and with 16-byte-aligned rank blocks.
      __shared__ __align__(128) float rows[kM2Warps][2][kM2Tile * 3];
      float* sg = rows[warp][k & 1];
      if (t * kM2Tile + kM2Tile <= n_vox) {
        // the staging buffer used two tiles ago must have been read by its bulk copies
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float2*>(sg + 6 * lane) = make_float2(pa.lo, pb.lo);
        *reinterpret_cast<float2*>(sg + 6 * lane + 2) = make_float2(r2.lo, pa.hi);
        *reinterpret_cast<float2*>(sg + 6 * lane + 4) = make_float2(pb.hi, r2.hi);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          // one TMA bulk store of the 768-byte block per rank: local HBM for the own rank, NVLink otherwise
          const int64_t base = (a.gather_row0 + (int64_t)t * kM2Tile) * 3;
#pragma unroll
          for (int r = 0; r < kMaxPeers; ++r) {
            if (r < a.gather_world) {
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.gather[r] + base),
                           "r"(smem_u32(sg)), "n"(kM2Tile * 3 * 4)
                           : "memory");
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {  // ragged last tile: row by row
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
          if (r < a.gather_world) {
            float* dst = a.gather[r] + (a.gather_row0 + v0) * 3;
            if (validA) { dst[0] = pa.lo; dst[1] = pb.lo; dst[2] = r2.lo; }
            if (validB) { dst[3] = pa.hi; dst[4] = pb.hi; dst[5] = r2.hi; }
          }
        }
      }
    }
    if (GATHER && a.popt == nullptr) {
      // the maps are the only output
    } else if (!a.po.enabled && a.out_dtype == DT_F32 && validB) {
      __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(a.popt) + (int64_t)v0 * P), make_float4(pa.lo, pb.lo, pa.hi, pb.hi));
      __stcs(reinterpret_cast<float2*>(reinterpret_cast<float*>(a.r2) + v0), make_float2(r2.lo, r2.hi));
      if (a.status) {
        a.status[v0] = (uint8_t)st[0];
        a.status[v0 + 1] = (uint8_t)st[1];
      }
      if (a.niter) {
        a.niter[v0] = (uint8_t)iters[0];
        a.niter[v0 + 1] = (uint8_t)iters[1];
      }
    } else {
      const T p0_[P] = {pa.lo, pb.lo}, p1_[P] = {pa.hi, pb.hi};
      if (validA) store_voxel<P, T, EMAX, false>(a, v0, p0_, r2.lo, true, st[0], iters[0]);
      if (validB) store_voxel<P, T, EMAX, false>(a, v0 + 1, p1_, r2.hi, true, st[1], iters[1]);
    }
    n_fit += (unsigned)(st[0] >= ST_CONV_F) + (unsigned)(st[1] >= ST_CONV_F);
    it_sum += (unsigned)(iters[0] + iters[1]);
    const unsigned im = (unsigned)(iters[0] > iters[1] ? iters[0] : iters[1]);
    it_max = im > it_max ? im : it_max;
  }
  if constexpr (GATHER) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores of this warp are done
  }
  // statistics: per-thread accumulators -> one reduction per warp at the end of the kernel
  {
    const unsigned fm = 0xffffffffu;
    const unsigned v0 = __reduce_add_sync(fm, n_fit);
    const unsigned v4 = __reduce_add_sync(fm, it_sum), v5 = __reduce_max_sync(fm, it_max);
    if (lane == 0) {
      if (v0) atomicAdd(stat_slot + CNT_FITTED, (unsigned long long)v0);
      if (v4) atomicAdd(stat_slot + CNT_ITERS, (unsigned long long)v4);
      if (v5) atomicMax(stat_slot + CNT_MAXITER, (unsigned long long)v5);
    }
  }
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
  float* gather[kMaxPeers];
  int gather_world;
  int64_t gather_row0;
  const CUtensorMap* tmap;   // host pointer to an encoded 2-D map of the planar fp32 samples (box 32 x E), or null
  const CUtensorMap* tmap2;  // the same with a 64-voxel box, for the two-voxels-per-lane kernel, or null
  int sm_count;
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
  a.status = d.status;
  a.niter = d.niter;
  a.mask_fill = d.mask_fill;
  a.counters = d.counters;
  for (int r = 0; r < kMaxPeers; ++r) a.gather[r] = d.gather[r];
  a.gather_world = d.gather_world;
  a.gather_row0 = d.gather_row0;
}

#if defined(__CUDACC__)
template <class M, typename T, int EMAX, bool EXACT>
inline cudaError_t launch_one(const LaunchDesc& d) {
  KernelArgs<T, EMAX> a;
  fill_args<T, EMAX>(d, a);
  if constexpr (M::MONO && EXACT && sizeof(T) == 4 && EMAX >= 3) {
    // dense fast path, two voxels per lane (see fit_kernel_mono2 for what it needs)
    const bool dt_ok = d.y_dtype == DT_F32 || d.y_dtype == DT_I16 || d.y_dtype == DT_U16;
    const size_t pair_bytes = 2 * dtype_size(d.y_dtype);
    // with the fused all-gather only the TMA kernel qualifies (raw parameters, 16-byte-aligned rank blocks)
    // Measured (weak scaling, 384^3 x 8 echoes per GPU): the one-voxel kernel's warp-transposed peer stores reach
    // 696 / 680 GB/s of NVLink egress at 4 / 8 GPUs, the TMA bulk stores of this kernel 652 / 657 GB/s (equal at 2),
    // and the step is NVLink-bound there -- so the bulk-store gather is opt-in (use_tma = 1).
    const bool gather_ok = d.gather_world == 0 || (d.use_tma == 1 && d.tmap2 != nullptr && d.y_dtype == DT_F32 &&
                                                   !d.po.enabled && d.gather_row0 % 4 == 0);
    if (d.fast_path == 1 && !a.vo.has_bounds && d.mask == nullptr && gather_ok && dt_ok && d.layout == LAYOUT_PLANAR &&
        (d.popt != nullptr || d.gather_world > 0) && reinterpret_cast<uintptr_t>(d.y) % pair_bytes == 0 && d.ld % 2 == 0 &&
        reinterpret_cast<uintptr_t>(d.popt) % 16 == 0 && reinterpret_cast<uintptr_t>(d.r2) % 8 == 0) {
      if (d.tmap2 != nullptr) {  // persistent, tiles staged through shared memory by TMA
        auto kfn = d.gather_world > 0    ? fit_kernel_mono2_tma<M, EMAX, true, float>
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
        kfn<<<(unsigned)g, kM2Warps * 32, 0, d.stream>>>(a, *d.tmap2);
        return cudaGetLastError();
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
      if (d.fast_path == 1 && !a.vo.has_bounds && d.gather_world == 0) {  // two voxels per lane over the list
        fit_kernel_mono2_list<M, EMAX><<<(unsigned)g, kBlock, 0, d.stream>>>(a);
        return cudaGetLastError();
      }
    }
    if (d.gather_world > 0) {
      if constexpr (sizeof(T) == 4) fit_kernel<M, T, EMAX, EXACT, true><<<(unsigned)g, kBlock, 0, d.stream>>>(a);
      else return cudaErrorNotSupported;
    } else {
      fit_kernel<M, T, EMAX, EXACT, false><<<(unsigned)g, kBlock, 0, d.stream>>>(a);
    }
    return cudaGetLastError();
  }
  // the fused all-gather epilogue is a separate instance so that single-GPU launches do not pay its registers
  if (d.gather_world > 0) {
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

// Split in two halves so that each model/dtype compiles as two translation units in parallel.
template <class M, typename T>
inline cudaError_t launch_model_lo(const LaunchDesc& d) {  // 1..8 echoes
  switch (d.n_echo) {
    case 1: return launch_exact<M, T, 1>(d);
    case 2: return launch_exact<M, T, 2>(d);
    case 3: return launch_exact<M, T, 3>(d);
    case 4: return launch_exact<M, T, 4>(d);
    case 5: return launch_exact<M, T, 5>(d);
    case 6: return launch_exact<M, T, 6>(d);
    case 7: return launch_exact<M, T, 7>(d);
    default: return launch_exact<M, T, 8>(d);
  }
}

template <class M, typename T>
inline cudaError_t launch_model_hi(const LaunchDesc& d) {  // 9..32 echoes
  switch (d.n_echo) {
    case 9: return launch_exact<M, T, 9>(d);
    case 10: return launch_exact<M, T, 10>(d);
    case 11: return launch_exact<M, T, 11>(d);
    case 12: return launch_exact<M, T, 12>(d);
    case 13: return launch_exact<M, T, 13>(d);
    case 14: return launch_exact<M, T, 14>(d);
    case 15: return launch_exact<M, T, 15>(d);
    case 16: return launch_exact<M, T, 16>(d);
    default: return launch_one<M, T, 32, false>(d);
  }
}
#endif

// Implemented one per translation unit so the (large, fully unrolled) instances compile in parallel.
#define DFIT_DECLARE_LAUNCH(name)                    \
  cudaError_t launch_##name##_lo(const LaunchDesc& d); \
  cudaError_t launch_##name##_hi(const LaunchDesc& d); \
  inline cudaError_t launch_##name(const LaunchDesc& d) { return d.n_echo <= 8 ? launch_##name##_lo(d) : launch_##name##_hi(d); }
DFIT_DECLARE_LAUNCH(mono_f32)
DFIT_DECLARE_LAUNCH(mono_f64)
DFIT_DECLARE_LAUNCH(biexp_f32)
DFIT_DECLARE_LAUNCH(biexp_f64)
DFIT_DECLARE_LAUNCH(linear_f32)
DFIT_DECLARE_LAUNCH(linear_f64)
#undef DFIT_DECLARE_LAUNCH

}  // namespace dfit
