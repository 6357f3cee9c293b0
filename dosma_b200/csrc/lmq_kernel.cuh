// Levenberg-Marquardt in rounds, with a per-warp queue of suspended fits (fit_kernel_lmq).
//
// The LM's pass count varies widely from voxel to voxel -- bi-exponential benchmark volume: 4 .. 30 passes, mean
// 7.5, 2.4 % run into maxfev -- and with one voxel per lane for the lifetime of a warp (fit_kernel) the warp runs as
// long as its slowest voxel: ncu showed 12.5 of 32 lanes active per instruction (profiles/r02_ncu_biexp16_lm.md).
// Here the warps are persistent and the solver runs in ROUNDS: a warp takes 32 voxels, gives every lane a budget
// of model evaluations (lm_iterate), and when the round is over
//   * finished voxels are stored,
//   * unfinished ones are SUSPENDED: their whole solver state (samples, parameters, normal equations, damping,
//     the step about to be evaluated: LmState) is pushed onto the warp's stack in shared memory,
// and the next round runs either 32 suspended fits popped from that stack -- all lanes busy again -- or, while
// fewer than 32 are waiting, the next 32 fresh voxels.  All housekeeping happens once per round with the whole warp
// (no per-lane refill inside the iteration, which is what sank the lane-refill experiment of
// profiles/experiments/), the stack is private to the warp (no atomics, no block barrier), and per voxel the
// arithmetic is lm_solve's bit for bit -- a round boundary only decides WHICH lane evaluates the next trial point.
//
// Used for every fit that goes straight to the LM in fp32 on a single GPU: the bi-exponential and linear models,
// the mono-exponential model with fast_path = 0 or y_bounds; dense ranges and the compacted list of the mask path.
#pragma once

#include <cstdio>
#include <cstdlib>

#include "kernel_common.cuh"

namespace dfit {

#if defined(__CUDACC__)

constexpr int kLmqCap = 64;  // stack slots per warp: at most 31 left waiting + 32 pushed by one round

// integers ride in the state words bit for bit
template <typename T>
struct LmqWord;
template <>
struct LmqWord<float> {
  static __device__ __forceinline__ float from_int(int v) { return __int_as_float(v); }
  static __device__ __forceinline__ int to_int(float w) { return __float_as_int(w); }
};
template <>
struct LmqWord<double> {
  static __device__ __forceinline__ double from_int(int v) { return __longlong_as_double((long long)v); }
  static __device__ __forceinline__ int to_int(double w) { return (int)__double_as_longlong(w); }
};

template <int P, int EMAX>
struct LmqLayout {  // word offsets of one suspended fit; slot s of word w lives at [w][s] (lanes -> consecutive banks)
  static constexpr int NA = P * (P + 1) / 2;
  static constexpr int Y = 0, PAR = Y + EMAX, PT = PAR + P, A = PT + P, G = A + NA, D2 = G + P, F = D2 + P, LAM = F + 1,
                       NU = LAM + 1, YSQ = NU + 1, ZZ = YSQ + 1, PN = ZZ + 1, PRED = PN + 1, FEV = PRED + 1, ITERS = FEV + 1,
                       VOX = ITERS + 1, WORDS = VOX + 1;
  static constexpr size_t bytes(size_t word, int warps) { return (size_t)warps * WORDS * kLmqCap * word; }
};

// Warps per SM the register allocation is asked to leave room for (128 registers per thread at 16, 80 at 24) -- in ONE
// CTA per SM: with the same number of warps, larger CTAs measured faster throughout (bi-exponential, config 4: 4 CTAs of
// 4 warps 1.31 ms / 2 x 8 1.28 / 1 x 16 1.21; two-parameter LM on pure noise: 6 x 4 4.33 ms / 3 x 8 4.17 / 2 x 12 4.05 /
// 1 x 24 3.81).
#ifndef DFIT_LMQ_WARPS_SM4
#define DFIT_LMQ_WARPS_SM4 16  // 4-parameter models above 8 echoes: measured 12 warps (158 registers) 1.54 ms / 16 (128) 1.44 / 20 (96, spills) slower
#endif
#ifndef DFIT_LMQ_WARPS_SM2
#define DFIT_LMQ_WARPS_SM2 24  // one / two parameters, up to 8 echoes (28: 72 registers with spills, slower)
#endif
constexpr int lmq_cta_warps(int P, int E, int word) {
  return word == 8 ? (P >= 4 || E > 8 ? 8 : 12)  // fp64: twice the registers and twice the stack
                   : P >= 4 ? (E <= 8 ? 16 : DFIT_LMQ_WARPS_SM4) : (E <= 8 ? DFIT_LMQ_WARPS_SM2 : 16);
}

template <class M, typename T, int EMAX, bool UNI>
__global__ void __launch_bounds__(lmq_cta_warps(M::P, EMAX, sizeof(T)) * 32, 1)
    fit_kernel_lmq(const __grid_constant__ KernelArgs<T, EMAX> a, const int k_first, const int k_next) {
  constexpr int P = M::P;
  constexpr int kLmqWarps = lmq_cta_warps(P, EMAX, sizeof(T));
  constexpr int NA = P * (P + 1) / 2;
  typedef LmqLayout<P, EMAX> L;
  extern __shared__ __align__(16) unsigned char lmq_smem_raw[];
  T* const lmq_smem = reinterpret_cast<T*>(lmq_smem_raw);
  const unsigned full = 0xffffffffu;
  const int warp = __shfl_sync(full, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  T* const q = lmq_smem + (size_t)warp * L::WORDS * kLmqCap;  // this warp's stack: q[w * kLmqCap + slot]
  // launched as the LM tail of a dense fast-path kernel (programmatic stream serialisation): wait until that grid has
  // completed and its list is visible; a no-op in a plain launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (a.lm_count_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.lm_count_next = 0u;  // the next launch's counter
  const bool listed = a.index != nullptr;
  const int64_t total = listed ? (int64_t)*a.index_count : a.n;
  const int64_t n_batches = (total + 31) >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * kLmqWarps;
  int64_t batch = (int64_t)blockIdx.x * kLmqWarps + warp;  // warp-uniform: batches are dealt round-robin
  int n_q = 0;                                             // suspended fits on the stack (warp-uniform)
  int it_sum = 0, it_max = 0;
  unsigned n_fit = 0, n_fail = 0, n_nf = 0, n_oob = 0;

  for (;;) {
    T y[EMAX], p[P];
    LmState<P, T, T> s;
    int64_t v = 0;
    int st = ST_PENDING, budget;
    unsigned flags = 0;
    bool live = false, resume;
    if (n_q < 32 && batch < n_batches) {
      // ---- 32 fresh voxels ----
      const int64_t i = (batch << 5) + lane;
      batch += n_warps;
      resume = false;
      budget = k_first;
      if (i < total) {
        live = true;
        v = listed ? (int64_t)a.index[i] : i;
        load_samples<T, EMAX, true>(a, listed ? v - a.g.y_voxel0 : v, y);
        load_p0<P, T, EMAX>(a, v, p);
        st = voxel_prepare<M, T, EMAX, true>(y, a.xt, a.E, a.vo, p, flags);
        s.F = 0;
        s.iters = 0;
        if (st == ST_PENDING) st = lm_begin<M, T, T, EMAX, true, UNI>(p, y, a.xt, a.E, a.vo.s, s);
      }
    } else if (n_q > 0) {
      // ---- up to 32 suspended fits off the stack ----
      const int cnt = n_q < 32 ? n_q : 32;
      n_q -= cnt;
      resume = true;
      budget = k_next;
      if (lane < cnt) {
        live = true;
        const T* __restrict__ src = q + n_q + lane;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) y[e] = src[(L::Y + e) * kLmqCap];
#pragma unroll
        for (int i = 0; i < P; ++i) {
          p[i] = src[(L::PAR + i) * kLmqCap];
          s.pt[i] = src[(L::PT + i) * kLmqCap];
          s.g[i] = src[(L::G + i) * kLmqCap];
          s.D2[i] = src[(L::D2 + i) * kLmqCap];
        }
#pragma unroll
        for (int k = 0; k < NA; ++k) s.A[k] = src[(L::A + k) * kLmqCap];
        s.F = src[L::F * kLmqCap];
        s.lam = src[L::LAM * kLmqCap];
        s.nu = src[L::NU * kLmqCap];
        s.ysq = src[L::YSQ * kLmqCap];
        s.zz = src[L::ZZ * kLmqCap];
        s.pnorm2 = src[L::PN * kLmqCap];
        s.pred = src[L::PRED * kLmqCap];
        s.fev = LmqWord<T>::to_int(src[L::FEV * kLmqCap]);
        s.iters = LmqWord<T>::to_int(src[L::ITERS * kLmqCap]);
        v = (int64_t)(unsigned)LmqWord<T>::to_int(src[L::VOX * kLmqCap]);
      }
      __syncwarp();  // every slot has been read before this round's pushes may overwrite it
    } else {
      break;
    }

    // ---- one round of the solver ----
    if (live && st == ST_PENDING) st = lm_iterate<M, T, T, EMAX, true, UNI>(p, y, a.xt, a.E, a.vo.s, s, budget, resume);

    // ---- retire: store what is finished, suspend what is not ----
    const bool pending = live && st == ST_PENDING;
    if (live && !pending) {
      T r2;
      voxel_finish<M, T, EMAX, true>(st, y, a.E, a.vo, s.F, p, r2);
      store_voxel<P, T, EMAX, false>(a, v, p, r2, true, st, s.iters);
      it_sum += s.iters;
      it_max = s.iters > it_max ? s.iters : it_max;
      n_fit += (unsigned)(st >= ST_CONV_F);
      n_fail += (unsigned)(st >= ST_MAXITER);
      n_nf += (unsigned)((flags & FLAG_NONFINITE) != 0);
      n_oob += (unsigned)((flags & FLAG_OOB) != 0);
    }
    const unsigned m = __ballot_sync(full, pending);
    if (pending) {
      T* __restrict__ dst = q + n_q + __popc(m & below);
#pragma unroll
      for (int e = 0; e < EMAX; ++e) dst[(L::Y + e) * kLmqCap] = y[e];
#pragma unroll
      for (int i = 0; i < P; ++i) {
        dst[(L::PAR + i) * kLmqCap] = p[i];
        dst[(L::PT + i) * kLmqCap] = s.pt[i];
        dst[(L::G + i) * kLmqCap] = s.g[i];
        dst[(L::D2 + i) * kLmqCap] = s.D2[i];
      }
#pragma unroll
      for (int k = 0; k < NA; ++k) dst[(L::A + k) * kLmqCap] = s.A[k];
      dst[L::F * kLmqCap] = s.F;
      dst[L::LAM * kLmqCap] = s.lam;
      dst[L::NU * kLmqCap] = s.nu;
      dst[L::YSQ * kLmqCap] = s.ysq;
      dst[L::ZZ * kLmqCap] = s.zz;
      dst[L::PN * kLmqCap] = s.pnorm2;
      dst[L::PRED * kLmqCap] = s.pred;
      dst[L::FEV * kLmqCap] = LmqWord<T>::from_int(s.fev);
      dst[L::ITERS * kLmqCap] = LmqWord<T>::from_int(s.iters);
      dst[L::VOX * kLmqCap] = LmqWord<T>::from_int((int)(unsigned)v);
    }
    n_q += __popc(m);
    __syncwarp();
  }
  block_stats_counts(a.counters, n_fit, n_fail, n_nf, n_oob, it_sum, it_max);
}

// Evaluations per round: `k_first` for fresh voxels (after the start point's own pass), `k_next` for resumed ones.
// DFIT_LMQ=0 switches the kernel off (the plain one-voxel-per-lane kernel runs instead), DFIT_LMQ=a,b sets the budgets
// (for A/B runs).
struct LmqConfig {
  int enabled, k_first, k_next;
};
inline LmqConfig lmq_config(int P = 4) {  // (read at every launch: a getenv, so that tests can switch within one process)
  LmqConfig c{1, 5, 2};  // measured on config 4 (bi-exponential, 16 echoes): 4,3 1.50 ms / 4,2 1.46 / 5,2 1.44 / 6,2 1.50
  if (P < 4) c = LmqConfig{1, 6, 6};  // (pure-noise volume: 6,4 6.93 ms / 6,6 6.78 / 8,4 6.88 / 8,8 6.87 / 4,3 7.35)
  if (const char* e = std::getenv("DFIT_LMQ")) {
    int a = 0, b = 0;
    const int n = std::sscanf(e, "%d,%d", &a, &b);
    if (n >= 1 && a <= 0) c.enabled = 0;
    if (n >= 1 && a > 0) c.k_first = c.k_next = a;
    if (n >= 2 && b > 0) c.k_next = b;
  }
  return c;
}

// `tail`: launched behind the dense fast-path kernel whose LM list it fits, with programmatic stream serialisation: the
// launch is set up while that kernel still runs and griddepcontrol.wait orders the memory.  The dense kernel never
// releases its dependents early (no griddepcontrol.launch_dependents): measured on a pure-noise volume, tail CTAs that
// become resident while the dense grid is still running cost 1.5 ms (7.07 against 5.51 ms); on the benchmark volume the
// (empty) tail then costs 0.7 us per launch.
template <class M, typename T, int EMAX>
inline cudaError_t launch_lmq(const LaunchDesc& d, const KernelArgs<T, EMAX>& a, bool tail = false) {
  const LmqConfig cfg = lmq_config(M::P);
  // uniformly spaced echoes (the host decided: fill_xtab): the exponentials of the model come from a two-echo
  // recurrence instead of MUFU.EX2 (DFIT_LMQ_UNI=0 switches that off, for A/B runs)
  constexpr bool CAN_UNI = M::HAS_REC && EMAX >= 4 && sizeof(T) == 4;
  bool uni = CAN_UNI && a.xt.uniform != 0;
  if (const char* e = std::getenv("DFIT_LMQ_UNI")) uni = uni && e[0] != '0';
  auto kfn = uni ? fit_kernel_lmq<M, T, EMAX, CAN_UNI> : fit_kernel_lmq<M, T, EMAX, false>;
  constexpr int kLmqWarps = lmq_cta_warps(M::P, EMAX, sizeof(T));
  const size_t smem = LmqLayout<M::P, EMAX>::bytes(sizeof(T), kLmqWarps);
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kLmqWarps * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  int64_t g = (int64_t)d.sm_count * per_sm;
  const int64_t needed = (d.n_vox + kLmqWarps * 32 - 1) / (kLmqWarps * 32);
  if (g > needed) g = needed;
  if (g < 1) g = 1;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)g);
  lc.blockDim = dim3(kLmqWarps * 32);
  lc.dynamicSmemBytes = smem;
  lc.stream = d.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = tail ? 1 : 0;
  return cudaLaunchKernelEx(&lc, kfn, a, cfg.k_first, cfg.k_next);
}

#endif  // __CUDACC__

}  // namespace dfit
