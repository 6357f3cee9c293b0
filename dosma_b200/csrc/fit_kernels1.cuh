// One voxel per lane: mask compaction, the general kernel (every model, fp32 / fp64, any layout and sample
// type, fused all-gather epilogue) and its persistent TMA-staged variant.
#pragma once

#include "kernel_common.cuh"

namespace dfit {

#if defined(__CUDACC__)

constexpr int kCompactPerThread = 8;  // voxels per thread in the compaction pass (2048 per CTA)

// The fill value into every output of a run of voxels outside the mask (single GPU, popt and r2 16-byte aligned), as
// 16-byte streaming stores -- the pass over a thin tissue mask IS this fill (config 3: 805 MB of it against 1.6 M voxels
// to fit), and per-voxel 8- and 4-byte stores ran it at 2.7 TB/s.  Two shapes: a thread's own 8 voxels (v a multiple of
// 8, idx0 = 0, stride = 1), or a whole warp's 256 voxels with the lanes interleaved (v a multiple of 256, idx0 = lane,
// stride = 32) so that every store instruction writes 512 contiguous bytes -- a thread's own 64 bytes would be written
// as half sectors by consecutive instructions.
template <int P, typename T, int EMAX>
__device__ __forceinline__ void fill_run(const KernelArgs<T, EMAX>& a, int64_t v, int idx0, int stride) {
  const int pw = a.sel >= 0 ? 1 : P;  // 1, 2 or 4 values per voxel
  double pat[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double q = a.fill_q[0];
#pragma unroll
    for (int i = 1; i < P; ++i)
      if (i == (a.sel >= 0 ? a.sel : (k & (pw - 1)))) q = a.fill_q[i];
    pat[k] = q;
  }
  if (a.out_dtype == DT_F32) {
    const float4 q4 = make_float4((float)pat[0], (float)pat[1], (float)pat[2], (float)pat[3]);
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.popt) + v * pw) + idx0;
    for (int k = 0; k < 2 * pw; ++k) __stcs(dst + k * stride, q4);
    const float mf = (float)a.mask_fill;
    float4* dr = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.r2) + v) + idx0;
    __stcs(dr, make_float4(mf, mf, mf, mf));
    __stcs(dr + stride, make_float4(mf, mf, mf, mf));
  } else {
    // (a double2 holds values 2 i and 2 i + 1 of the run: which parameters depends on the parity of i for pw = 4)
    const double2 even = make_double2(pat[0], pat[1]), odd = make_double2(pat[2], pat[3]);
    double2* dst = reinterpret_cast<double2*>(reinterpret_cast<double*>(a.popt) + v * pw) + idx0;
    for (int k = 0; k < 4 * pw; ++k) __stcs(dst + k * stride, ((idx0 + k * stride) & 1) ? odd : even);
    double2* dr = reinterpret_cast<double2*>(reinterpret_cast<double*>(a.r2) + v) + idx0;
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(dr + k * stride, make_double2(a.mask_fill, a.mask_fill));
  }
  if (a.status) *(reinterpret_cast<uint2*>(a.status + v) + idx0) = make_uint2(0u, 0u);  // ST_SKIPPED
  if (a.niter) *(reinterpret_cast<uint2*>(a.niter + v) + idx0) = make_uint2(0u, 0u);
}

// The same for the multi-GPU split mode with 8-byte [parameter, r2] rows: outside the mask every rank fills its OWN
// reassembled map (nothing of the fill crosses NVLink); rows of a run of voxels are contiguous.
template <int P, typename T, int EMAX>
__device__ __forceinline__ void fill_run_map2(const KernelArgs<T, EMAX>& a, int64_t v, int idx0, int stride) {
  float q = (float)a.fill_q[0];
#pragma unroll
  for (int i = 1; i < P; ++i)
    if ((a.g.cols >> i) & 1u) q = (float)a.fill_q[i];
  const float mf = (float)a.mask_fill;
  float4* dst = reinterpret_cast<float4*>(a.g.maps[a.g.self] + (a.g.row0 + v) * 2) + idx0;
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k * stride] = make_float4(q, mf, q, mf);
}

// Mask path, step 1: one streaming pass over the mask that (a) appends the voxels to fit to a compact index list and
// (b) writes the fill value for every voxel outside the mask (fitting.py:205-215).  Step 2 is the fit kernel over the
// list: all 32 lanes of a warp fit, however thin the tissue mask is.  A thread takes 8 CONSECUTIVE voxels -- one 8-byte
// load of their mask bytes, 16-byte fill stores -- a warp claims its run of the CTA's list with one shared-memory atomic
// after a shuffle scan of the lanes' counts (neighbours stay neighbours), the CTA its run of the global list with ONE
// global atomic per 2048 voxels.
template <int P, typename T, int EMAX>
__global__ void __launch_bounds__(256) mask_compact_kernel(const __grid_constant__ KernelArgs<T, EMAX> a, unsigned* index,
                                                           unsigned* count) {
  __shared__ unsigned s_list[256 * kCompactPerThread];
  __shared__ unsigned s_n, s_base;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int64_t v0 = (int64_t)blockIdx.x * (256 * kCompactPerThread) + (int64_t)threadIdx.x * kCompactPerThread;
  const int lane = threadIdx.x & 31;
  const bool whole = v0 + kCompactPerThread <= a.n;
  // mask bytes of the thread's 8 voxels: in range (inb), inside the mask (msk), to fit by this rank (act)
  unsigned inb = 0, msk = 0;
  if (whole && (reinterpret_cast<uintptr_t>(a.mask) & 7) == 0) {
    const uint2 w = __ldcs(reinterpret_cast<const uint2*>(a.mask + v0));
    inb = 0xffu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      msk |= ((w.x >> (8 * k)) & 0xffu) ? (1u << k) : 0u;
      msk |= ((w.y >> (8 * k)) & 0xffu) ? (16u << k) : 0u;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kCompactPerThread; ++k) {
      if (v0 + k < a.n) {
        inb |= 1u << k;
        msk |= a.mask[v0 + k] != 0 ? (1u << k) : 0u;
      }
    }
  }
  unsigned act = msk;
  if (a.g.split_list) {  // (multi-GPU split mode: masked voxels outside this rank's span are a peer's to fit -- neither listed nor filled)
#pragma unroll
    for (int k = 0; k < kCompactPerThread; ++k)
      if (!(v0 + k >= a.g.fit_lo && v0 + k < a.g.fit_hi)) act &= ~(1u << k);
  }
  // the warp's run of the CTA's list
  const unsigned cnt = (unsigned)__popc(act);
  if (__any_sync(0xffffffffu, cnt != 0u)) {
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    unsigned base = 0;
    if (lane == 31) base = atomicAdd(&s_n, incl);  // shared-memory atomic
    base = __shfl_sync(0xffffffffu, base, 31);
    unsigned pos = base + incl - cnt;
#pragma unroll
    for (int k = 0; k < kCompactPerThread; ++k)
      if ((act >> k) & 1u) s_list[pos++] = (unsigned)(v0 + k);
  }
  // the fill value outside the mask
  const unsigned fillm = inb & ~msk;
  const bool wide = a.g.world == 0 && a.popt != nullptr && ((reinterpret_cast<uintptr_t>(a.popt) | reinterpret_cast<uintptr_t>(a.r2)) & 15) == 0 &&
                    ((reinterpret_cast<uintptr_t>(a.status) | reinterpret_cast<uintptr_t>(a.niter)) & 7) == 0;
  // (split mode of a multi-GPU masked fit whose only output is the [parameter, r2] map: the same, into the own map)
  const bool wide_map = a.g.world > 0 && a.g.split_list && a.popt == nullptr && a.g.ncols == 2 &&
                        a.status == nullptr && a.niter == nullptr && (a.g.row0 & 1) == 0 &&
                        (reinterpret_cast<uintptr_t>(a.g.maps[a.g.self]) & 15) == 0;
  const bool run8 = fillm == 0xffu && (wide || wide_map);
  if (__all_sync(0xffffffffu, run8)) {  // the warp's 256 voxels, lanes interleaved
    if (wide) fill_run<P, T, EMAX>(a, v0 - lane * kCompactPerThread, lane, 32);
    else fill_run_map2<P, T, EMAX>(a, v0 - lane * kCompactPerThread, lane, 32);
  } else if (run8) {
    if (wide) fill_run<P, T, EMAX>(a, v0, 0, 1);
    else fill_run_map2<P, T, EMAX>(a, v0, 0, 1);
  } else if (fillm != 0u) {
#pragma unroll 1
    for (int k = 0; k < kCompactPerThread; ++k)
      if ((fillm >> k) & 1u) fill_voxel<P, T, EMAX>(a, v0 + k);
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
    const unsigned first = 0, count = *a.index_count;
    int it_sum = 0;
    unsigned n_fit = 0, n_fail = 0, n_nf = 0, n_oob = 0;
    for (unsigned i = blockIdx.x * kBlock + threadIdx.x; i < count; i += gridDim.x * kBlock) {
      const int64_t v = (int64_t)a.index[first + i];
      T p[P], r2 = 0, y[EMAX];
      int it = 0;
      unsigned fl = 0;
      load_samples<T, EMAX, EXACT>(a, v - a.g.y_voxel0, y);
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
// TMA-staged variant.  Persistent warps: every warp owns a 2-stage shared-memory ring of
// [E][32-voxel] sample tiles that the Tensor Memory Accelerator fills (cp.async.bulk.tensor.2d over a
// 2-D tensor map of the planar (E, ld) array, box = 32 voxels x E echoes) while the warp is busy
// fitting the previous tile; completion is signalled on a per-stage mbarrier (complete_tx::bytes).
// There is no block-level synchronisation -- warps run their own pipelines, so a slow voxel only
// holds its own warp.  Used for fp32 planar samples whose row pitch is a multiple of 16 bytes.
constexpr int kTmaWarps = 8;          // warps per CTA
constexpr int tma_stages(int E) { return E <= 8 ? 4 : 2; }  // ring depth: 32 KB of tiles per CTA at 8 echoes
constexpr int kTmaTile = 32;          // voxels per warp tile (one per lane)

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

#endif  // __CUDACC__

}  // namespace dfit
