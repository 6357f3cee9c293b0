// Two voxels per lane: the mono-exponential fast-path kernels -- plain loads (fit_kernel_mono2), the compacted
// voxel list of the mask path (fit_kernel_mono2_list) and the persistent TMA-staged kernel that the headline
// number runs (fit_kernel_mono2_tma, optionally with the fused all-gather by TMA bulk stores).
#pragma once

#include "kernel_common.cuh"

namespace dfit {

#if defined(__CUDACC__)

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
    {
      bool ok[2];
      fit_voxel_fast2s<M, T, EMAX, pair2<T>[EMAX]>(Y, a.xt, a.vo, pa, pb, r2, ok);
      st[0] = ok[0] ? (int)ST_CONV_F : -1;
      st[1] = ok[1] ? (int)ST_CONV_F : -1;
      iters[0] = iters[1] = kFast2Passes;
    }
    if (st[0] < 0 || st[1] < 0) {  // turned down: the one-voxel path (generic Newton loop, then the LM), one at a time
#pragma unroll 1
      for (int hsel = 0; hsel < 2; ++hsel) {
        if ((hsel ? st[1] : st[0]) >= 0) continue;
        T ys[EMAX], p[P], r = 0;
        int it = 0;
        unsigned fl = 0;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) ys[e] = hsel ? Y[e].hi : Y[e].lo;
        int s1 = fit_voxel_fast<M, T, EMAX, true>(ys, a.xt, a.vo, p, r, it);
        if (s1 < 0) {
          load_p0<P, T, EMAX>(a, hsel && both ? v0 + 1 : v0, p);
          s1 = fit_voxel<M, T, T, EMAX, true>(ys, a.xt, a.E, a.vo, p, r, it, fl);
        }
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
    if (a.out_dtype == DT_F32 && both && a.popt != nullptr) {
      float4 q = make_float4(pa.lo, pb.lo, pa.hi, pb.hi);
      if (a.po.enabled) {  // fused epilogue, fp32 form (see post_pair_f32)
        const pair2<float> qa = post_pair_f32(a.po, 0, pa, r2), qb = post_pair_f32(a.po, 1, pb, r2);
        q = make_float4(qa.lo, qb.lo, qa.hi, qb.hi);
      }
      if (a.sel >= 0) __stcs(reinterpret_cast<float2*>(reinterpret_cast<float*>(a.popt) + v0), a.sel == 0 ? make_float2(q.x, q.z) : make_float2(q.y, q.w));
      else __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(a.popt) + v0 * P), q);
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
template <class M, int EMAX, bool GATHER = false>
__global__ void __launch_bounds__(kBlock, 5) fit_kernel_mono2_list(const __grid_constant__ KernelArgs<float, EMAX> a) {
  typedef float T;
  constexpr int P = 2;
  // (multi-GPU split mode: the list holds the masked voxels of this rank's span only, see mask_compact_kernel)
  const unsigned count = *a.index_count;
  const unsigned npairs = (count + 1u) >> 1;
  const unsigned* __restrict__ list = a.index;
  int it_sum = 0, it_max = 0;
  unsigned n_fit = 0, n_fail = 0, n_nf = 0, n_oob = 0;
  for (unsigned i = blockIdx.x * kBlock + threadIdx.x; i < npairs; i += gridDim.x * kBlock) {
    const bool both = 2u * i + 1u < count;
    const int64_t vA = (int64_t)list[2u * i], vB = both ? (int64_t)list[2u * i + 1u] : vA;
    T yA[EMAX], yB[EMAX];
    load_samples<T, EMAX, true>(a, vA - a.g.y_voxel0, yA);
    load_samples<T, EMAX, true>(a, vB - a.g.y_voxel0, yB);
    pair2<T> Y[EMAX], pa, pb, r2;
#pragma unroll
    for (int e = 0; e < EMAX; ++e) Y[e] = p2_make<T>(yA[e], yB[e]);
    int st[2], iters[2];
    {
      bool ok[2];
      fit_voxel_fast2s<M, T, EMAX, pair2<T>[EMAX]>(Y, a.xt, a.vo, pa, pb, r2, ok);
      st[0] = ok[0] ? (int)ST_CONV_F : -1;
      st[1] = ok[1] ? (int)ST_CONV_F : -1;
      iters[0] = iters[1] = kFast2Passes;
    }
    if (st[0] < 0 || (st[1] < 0 && both)) {
#pragma unroll 1
      for (int hsel = 0; hsel < 2; ++hsel) {
        if ((hsel ? st[1] : st[0]) >= 0 || (hsel && !both)) continue;
        T ys[EMAX], p[P], r = 0;
        int it = 0;
        unsigned fl = 0;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) ys[e] = hsel ? Y[e].hi : Y[e].lo;
        int s1 = fit_voxel_fast<M, T, EMAX, true>(ys, a.xt, a.vo, p, r, it);
        if (s1 < 0) {
          load_p0<P, T, EMAX>(a, hsel ? vB : vA, p);
          s1 = fit_voxel<M, T, T, EMAX, true>(ys, a.xt, a.E, a.vo, p, r, it, fl);
        }
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
    store_voxel<P, T, EMAX, GATHER>(a, vA, p0_, r2.lo, true, st[0], iters[0]);
    if (both) store_voxel<P, T, EMAX, GATHER>(a, vB, p1_, r2.hi, true, st[1], iters[1]);
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
// Two voxels per lane + TMA staging (fp32 or raw 16-bit samples, converted on the way to registers):
// persistent warps, each with its own ring of [E][64-voxel] sample tiles in shared memory.  The Tensor Memory Accelerator fills a stage (one cp.async.bulk.tensor.2d over
// the 2-D map of the planar samples, box = 64 voxels x E echoes, completion on the stage's mbarrier) while
// the warp is fitting earlier tiles, so the HBM latency that the plain kernel exposes at the top of every
// CTA is hidden behind arithmetic.  Lanes read their two voxels of every echo as one conflict-free 8-byte
// shared load.  No block-level synchronisation inside the loop.
// Resident CTAs per SM the compiler is asked to make room for (register cap) and depth of each warp's tile ring.
// Measured on the benchmark volume (8 echoes; ms per 56.6 M voxels): 5 CTAs x 4 stages 0.525, 6 x 4 0.543,
// 7 x 3 0.505 (72 registers, no spills), 8 x 3 0.579 (64 registers, spills).  Above 8 echoes the samples alone
// take up to 32 registers: 4 CTAs.
#ifndef DFIT_M2_MIN_CTAS
#define DFIT_M2_MIN_CTAS 7
#endif
#ifndef DFIT_M2_STAGES
#define DFIT_M2_STAGES 3
#endif
constexpr int m2_min_ctas(int E) { return E <= 8 ? DFIT_M2_MIN_CTAS : 4; }
constexpr int kM2Warps = 4;
constexpr int kM2Tile = 64;
constexpr int kLmListMin = 16;  // LM-bound voxels among the 32 of a deferred batch from which they go to the LM tail's list
constexpr int kDeferCap = 96;  // per-warp queue of deferred voxels: at most 31 left over + 64 from one tile
constexpr int m2_stages(int E) { return E <= 8 ? DFIT_M2_STAGES : 2; }  // <= 24 KB (32 KB above 8 echoes) of tiles per CTA

// One voxel per lane over a warp's queue of deferred voxels (indices into the launch's voxel range): the
// generic Newton loop of the fast path first, the LM from the caller's initial guess where that declines.
// `lane < cnt` lanes work; the samples are re-read from global memory (the tile they came from is usually
// still in L2; deferred voxels are rare on tissue and LM-bound on noise).
template <class M, int EMAX, bool GATHER>
__device__ __forceinline__ void fit_deferred(const KernelArgs<float, EMAX>& a, const unsigned* __restrict__ queue, int cnt,
                                             int lane, unsigned& n_fit, unsigned& n_fail, unsigned& n_nf, unsigned& n_oob,
                                             unsigned& it_sum, unsigned& it_max) {
  typedef float T;
  constexpr int P = 2;
  const bool active = lane < cnt;
  int64_t v = 0;
  T y[EMAX], p[P], r2 = 0;
  int it = 0, st = -1;
  unsigned fl = 0;
  if (active) {
    v = (int64_t)queue[lane];
    load_samples<T, EMAX, true>(a, v, y);
    st = fit_voxel_fast<M, T, EMAX, true>(y, a.xt, a.vo, p, r2, it);
  }
  const unsigned m_lm = __ballot_sync(0xffffffffu, active && st < 0);  // (warp-uniform call sites: all 32 lanes are here)
  if (!GATHER && a.lm_list != nullptr && __popc(m_lm) >= kLmListMin) {
    // LM tail: what the Newton loop turned down as well goes onto the launch's list (one atomic per warp) and is fitted
    // by the LM-in-rounds kernel that follows this one -- the LM's pass count varies from 5 to 50 on such voxels, and run
    // here a warp would wait for its slowest lane.  Only when at least kLmListMin of the 32 are LM-bound (air, background):
    // on tissue they are one or two per batch, cheaper to fit right here behind the other warps' work than in a launch
    // of their own, and the list stays empty.  (A deterministic rule: where a voxel is fitted depends on the data only.)
    const bool to_lm = active && st < 0;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(a.lm_count, (unsigned)__popc(m_lm));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (to_lm) a.lm_list[base + __popc(m_lm & ((1u << lane) - 1u))] = (unsigned)v;
    if (!active || st < 0) return;
  } else if (active && st < 0) {
    load_p0<P, T, EMAX>(a, v, p);
    st = fit_voxel<M, T, T, EMAX, true>(y, a.xt, a.E, a.vo, p, r2, it, fl);
  }
  if (active) {
    store_voxel<P, T, EMAX, GATHER>(a, v, p, r2, true, st, it);
    n_fit += (unsigned)(st >= ST_CONV_F);
    n_fail += (unsigned)(st >= ST_MAXITER);
    n_nf += (unsigned)((fl & FLAG_NONFINITE) != 0);
    n_oob += (unsigned)((fl & FLAG_OOB) != 0);
    it_sum += (unsigned)it;
    it_max = (unsigned)it > it_max ? (unsigned)it : it_max;
  }
}

template <class M, int EMAX, bool GATHER, typename S>
__global__ void __launch_bounds__(kM2Warps * 32, m2_min_ctas(EMAX))
    fit_kernel_mono2_tma(const __grid_constant__ KernelArgs<float, EMAX> a, const __grid_constant__ CUtensorMap tmap) {
  typedef float T;
  constexpr int P = 2;
  constexpr int kStages = m2_stages(EMAX);
  constexpr unsigned kTileBytes = EMAX * kM2Tile * sizeof(S);  // S = float, or the raw 16-bit DICOM sample type
  __shared__ __align__(128) S tiles[kM2Warps][kStages][EMAX][kM2Tile];
  __shared__ __align__(8) uint64_t full[kM2Warps][kStages];
  __shared__ unsigned defer_q[kM2Warps][kDeferCap];
  // (the warp index through a shuffle: provably warp-uniform, so that the tile bookkeeping and the TMA operands
  // live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
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

  unsigned n_fast = 0, n_fit = 0, n_fail = 0, n_nf = 0, n_oob = 0, it_sum = 0, it_max = 0, n_deferred = 0;
  int n_def = 0;  // entries in this warp's queue (warp-uniform)
  // fp32 maps (raw parameters or the fp32 form of the fused epilogue) and nothing else to write: two vector stores per lane
  const bool plain = a.out_dtype == DT_F32 && a.popt != nullptr && a.status == nullptr && a.niter == nullptr;
  const int pw = a.sel >= 0 ? 1 : P;  // floats per voxel in popt (one selected parameter, or all)
  const int64_t kPoptTile = (int64_t)kM2Tile * pw * sizeof(float);
  constexpr int64_t kR2Tile = (int64_t)kM2Tile * sizeof(float);
  char* popt_lane = reinterpret_cast<char*>(a.popt) + lane * (2 * pw * (int)sizeof(float)) + (int64_t)warp_global * kPoptTile;
  char* r2_lane = reinterpret_cast<char*>(a.r2) + lane * (2 * sizeof(float)) + (int64_t)warp_global * kR2Tile;
  int stage = 0;
  unsigned phase = 0;
  [[maybe_unused]] int k = 0;
  for (int t = warp_global; t < n_tiles; t += warp_stride, ++k, popt_lane += warp_stride * kPoptTile, r2_lane += warp_stride * kR2Tile) {
    mbar_wait(&full[warp][stage], phase);
    pair2<T> Y[EMAX], pa, pb, r2;
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      const typename Vec2<S>::type v = *reinterpret_cast<const typename Vec2<S>::type*>(&tiles[warp][stage][e][2 * lane]);
      Y[e] = p2_make<T>((T)v.x, (T)v.y);
    }
    const int v0 = t * kM2Tile + 2 * lane;
    bool ok[2];
    fit_voxel_fast2s<M, T, EMAX, pair2<T>[EMAX]>(Y, a.xt, a.vo, pa, pb, r2, ok);
    // The stage was drained long ago (every sample has been used): refill it with the tile kStages trips ahead.
    // (Issued here rather than right after the shared loads so that the proxy fence finds nothing to wait for.)
    __syncwarp();
    if (lane == 0) {
      const int tn = t + kStages * warp_stride;
      if (tn < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[warp][stage], kTileBytes);
        tma_load_2d(&tiles[warp][stage][0][0], &tmap, tn * kM2Tile, 0, &full[warp][stage]);
      }
    }
    if (++stage == kStages) {
      stage = 0;
      phase ^= 1u;
    }
    bool defA = !ok[0], defB = !ok[1];
    if (t == n_tiles - 1) {  // the last tile may be ragged: voxels past the end are zero-filled and must go nowhere
      const bool validA = v0 < n_vox, validB = v0 + 1 < n_vox;
      ok[0] = ok[0] && validA;
      ok[1] = ok[1] && validB;
      defA = defA && validA;
      defB = defB && validB;
    }
    if constexpr (GATHER) {
      // Fused all-gather: the lane's two rows are adjacent in every rank's map.  With one parameter + r2 per row (the
      // T2 map: 8-byte rows) they are ONE 16-byte store per rank -- a warp writes 512 contiguous bytes -- or one
      // multicast store that the NVSwitch replicates; with all three columns, three 8-byte stores.  Rows of deferred
      // voxels are written when their queue is run.  (The launcher admits this kernel for these two row formats.)
      pair2<float> ga = pa, gb = pb;
      if (a.po.enabled) {
        ga = post_pair_f32(a.po, 0, pa, r2);
        gb = post_pair_f32(a.po, 1, pb, r2);
      }
      const int64_t row = a.g.row0 + v0;
      if (a.g.ncols == 2) {
        const pair2<float> gp = (a.g.cols & 1u) ? ga : gb;
        if (ok[0] && ok[1]) {
          const float w[4] = {gp.lo, r2.lo, gp.hi, r2.hi};
          gather_store<4>(a.g, row * 2, w);
        } else {
          const float w0[2] = {gp.lo, r2.lo}, w1[2] = {gp.hi, r2.hi};
          if (ok[0]) gather_store<2>(a.g, row * 2, w0);
          if (ok[1]) gather_store<2>(a.g, row * 2 + 2, w1);
        }
      } else {
        const float w0[2] = {ga.lo, gb.lo}, w1[2] = {r2.lo, ga.hi}, w2[2] = {gb.hi, r2.hi};
        if (ok[0] && ok[1]) {
          gather_store<2>(a.g, row * 3, w0);
          gather_store<2>(a.g, row * 3 + 2, w1);
          gather_store<2>(a.g, row * 3 + 4, w2);
        } else {
          const float s[6] = {ga.lo, gb.lo, r2.lo, ga.hi, gb.hi, r2.hi};
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float w[1] = {s[i]};
            if (i < 3 ? ok[0] : ok[1]) gather_store<1>(a.g, row * 3 + i, w);
          }
        }
      }
    }
    if (plain) {
      float4 q = make_float4(pa.lo, pb.lo, pa.hi, pb.hi);
      if (a.po.enabled) {  // fused _process_params + rounding (fitting.py:109-146, 734-737), fp32 form
        const pair2<float> qa = post_pair_f32(a.po, 0, pa, r2), qb = post_pair_f32(a.po, 1, pb, r2);
        q = make_float4(qa.lo, qb.lo, qa.hi, qb.hi);
      }
      if (a.sel >= 0) {  // one selected parameter (the time-constant map): popt is [N]
        const float2 qs = a.sel == 0 ? make_float2(q.x, q.z) : make_float2(q.y, q.w);
        float* pp = reinterpret_cast<float*>(popt_lane);
        float* pr = reinterpret_cast<float*>(r2_lane);
        if (ok[0] && ok[1]) {
          __stcs(reinterpret_cast<float2*>(pp), qs);
          __stcs(reinterpret_cast<float2*>(pr), make_float2(r2.lo, r2.hi));
        } else {
          if (ok[0]) { __stcs(pp, qs.x); __stcs(pr, r2.lo); }
          if (ok[1]) { __stcs(pp + 1, qs.y); __stcs(pr + 1, r2.hi); }
        }
      } else if (ok[0] && ok[1]) {
        __stcs(reinterpret_cast<float4*>(popt_lane), q);
        __stcs(reinterpret_cast<float2*>(r2_lane), make_float2(r2.lo, r2.hi));
      } else {
        float2* pp = reinterpret_cast<float2*>(popt_lane);
        float* pr = reinterpret_cast<float*>(r2_lane);
        if (ok[0]) { __stcs(pp, make_float2(q.x, q.y)); __stcs(pr, r2.lo); }
        if (ok[1]) { __stcs(pp + 1, make_float2(q.z, q.w)); __stcs(pr + 1, r2.hi); }
      }
    } else if (!(GATHER && a.popt == nullptr)) {  // (with the fused gather the maps may be the only output)
      const T p0_[P] = {pa.lo, pb.lo}, p1_[P] = {pa.hi, pb.hi};
      if (ok[0]) store_voxel<P, T, EMAX, false>(a, v0, p0_, r2.lo, true, ST_CONV_F, kFast2Passes);
      if (ok[1]) store_voxel<P, T, EMAX, false>(a, v0 + 1, p1_, r2.hi, true, ST_CONV_F, kFast2Passes);
    }
    n_fast += (unsigned)ok[0] + (unsigned)ok[1];
    // voxels the straight-line attempt turned down join the warp's queue; full warps of them are fitted at once
    const unsigned mA = __ballot_sync(0xffffffffu, defA), mB = __ballot_sync(0xffffffffu, defB);
    if ((mA | mB) != 0u) {
      const unsigned below = (1u << lane) - 1u;
      const int nA = __popc(mA);
      if (defA) defer_q[warp][n_def + __popc(mA & below)] = (unsigned)v0;
      if (defB) defer_q[warp][n_def + nA + __popc(mB & below)] = (unsigned)(v0 + 1);
      n_def += nA + __popc(mB);
      if (lane == 0) n_deferred += (unsigned)(nA + __popc(mB));
      __syncwarp();
      if (n_def >= 32) {
        do {
          n_def -= 32;
          fit_deferred<M, EMAX, GATHER>(a, &defer_q[warp][n_def], 32, lane, n_fit, n_fail, n_nf, n_oob, it_sum, it_max);
        } while (n_def >= 32);
        __syncwarp();
      }
    }
  }
  if (n_def > 0) fit_deferred<M, EMAX, GATHER>(a, &defer_q[warp][0], n_def, lane, n_fit, n_fail, n_nf, n_oob, it_sum, it_max);
  // statistics: per-thread accumulators -> one reduction per warp at the end of the kernel
  {
    const unsigned fm = 0xffffffffu;
    unsigned long long* const stat_slot = a.counters + (size_t)(warp_global & (kStatSlots - 1)) * CNT_COUNT;
    const unsigned c_fast = __reduce_add_sync(fm, n_fast);  // straight-line voxels: kFast2Passes passes each
    const unsigned c_fit = __reduce_add_sync(fm, n_fit) + c_fast, c_fail = __reduce_add_sync(fm, n_fail);
    const unsigned c_nf = __reduce_add_sync(fm, n_nf), c_oob = __reduce_add_sync(fm, n_oob);
    const unsigned c_it = __reduce_add_sync(fm, it_sum);
    unsigned c_max = __reduce_max_sync(fm, it_max);
    if (c_fast != 0u && c_max < (unsigned)kFast2Passes) c_max = (unsigned)kFast2Passes;
    if (lane == 0) {
      const unsigned long long its = (unsigned long long)c_it + (unsigned long long)kFast2Passes * c_fast;
      if (c_fit) atomicAdd(stat_slot + CNT_FITTED, (unsigned long long)c_fit);
      if (c_fail) atomicAdd(stat_slot + CNT_FAILED, (unsigned long long)c_fail);
      if (c_nf) atomicAdd(stat_slot + CNT_NONFINITE, (unsigned long long)c_nf);
      if (c_oob) atomicAdd(stat_slot + CNT_OOB, (unsigned long long)c_oob);
      if (its) atomicAdd(stat_slot + CNT_ITERS, its);
      if (c_max) atomicMax(stat_slot + CNT_MAXITER, (unsigned long long)c_max);
      if (n_deferred) atomicAdd(stat_slot + CNT_DEFERRED, (unsigned long long)n_deferred);
    }
  }
}

#endif  // __CUDACC__

}  // namespace dfit
