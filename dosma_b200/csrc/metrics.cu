// Region statistics of a fitted map on the GPU -- the step after the fit in every DOSMA pipeline
// (SURVEY.md section 8 row f4): `QuantitativeValue.to_metrics`, dosma/core/quant_vals.py:145-229.
//
// For every region (a label of a label mask, "any positive label", or the whole volume) over the
// VALID voxels (finite and inside `bounds`, :182-190): count, mean (np.nanmean), population standard
// deviation (np.nanstd) and median (np.nanmedian), all in float64 like numpy.
//   pass 1  count + sum                    -> mean
//   pass 2  sum of squared deviations      -> std   (two-pass: no cancellation)
//   pass 3+ exact median by radix selection on the order-preserving 64-bit key of each value, 8 bits
//           per pass, for the two middle ranks of every region at once (even counts average them)
// Every pass is one streaming read of map + labels (HBM-bound); histograms live in shared memory.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "dfit_internal.h"

namespace dfit {

constexpr int kMaxRegions = 16;
constexpr int kMaxQueries = 2 * kMaxRegions;
constexpr int kRadixBits = 8;
constexpr int kBins = 1 << kRadixBits;

struct MetricsArgs {
  const void* map;
  const void* labels;  // null: no label mask
  int64_t n;
  int map_dtype, labels_dtype;
  int n_regions;
  int region_label[kMaxRegions];  // > 0: that label; -1: any positive label; -2: every voxel
  int has_bounds, closed_left, closed_right;
  double lb, ub;
};

__device__ __forceinline__ bool voxel_valid(const MetricsArgs& a, double v) {
  bool ok = isfinite(v);
  if (a.has_bounds) {
    ok = ok && (a.closed_left ? v >= a.lb : v > a.lb) && (a.closed_right ? v <= a.ub : v < a.ub);
  }
  return ok;
}

__device__ __forceinline__ int load_label(const MetricsArgs& a, int64_t i) {
  if (!a.labels) return 0;
  switch (a.labels_dtype) {
    case DT_U8: return (int)reinterpret_cast<const unsigned char*>(a.labels)[i];
    case DT_I16: return (int)reinterpret_cast<const short*>(a.labels)[i];
    case DT_U16: return (int)reinterpret_cast<const unsigned short*>(a.labels)[i];
    case DT_I32: return reinterpret_cast<const int*>(a.labels)[i];
    case DT_F32: return (int)reinterpret_cast<const float*>(a.labels)[i];
    default: return (int)reinterpret_cast<const double*>(a.labels)[i];
  }
}

__device__ __forceinline__ bool in_region(const MetricsArgs& a, int r, int label) {
  const int want = a.region_label[r];
  return want == -2 || (want == -1 ? label > 0 : label == want);
}

// order-preserving map double -> uint64
__device__ __forceinline__ unsigned long long ordered_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static inline double key_to_double(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double v;
  memcpy(&v, &b, sizeof(v));
  return v;
}

// pass 1 (MODE 0): count, sum.  pass 2 (MODE 1): sum (v - mean)^2.
template <int MODE>
__global__ void __launch_bounds__(256) metrics_moments_kernel(const __grid_constant__ MetricsArgs a, const double* mean,
                                                              double* out /*[regions][2]*/) {
  __shared__ double s_acc[kMaxRegions][2];
  for (int i = threadIdx.x; i < kMaxRegions * 2; i += blockDim.x) (&s_acc[0][0])[i] = 0.0;
  __syncthreads();
  double cnt[kMaxRegions], sum[kMaxRegions];
#pragma unroll
  for (int r = 0; r < kMaxRegions; ++r) cnt[r] = sum[r] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = load_as<double>(a.map, a.map_dtype, i);
    if (!voxel_valid(a, v)) continue;
    const int label = load_label(a, i);
#pragma unroll
    for (int r = 0; r < kMaxRegions; ++r) {
      if (r < a.n_regions && in_region(a, r, label)) {
        if (MODE == 0) {
          cnt[r] += 1.0;
          sum[r] += v;
        } else {
          const double d = v - mean[r];
          sum[r] += d * d;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kMaxRegions; ++r) {
    if (r < a.n_regions) {
      double c = cnt[r], s = sum[r];
      for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      if ((threadIdx.x & 31) == 0) {
        if (MODE == 0 && c != 0.0) atomicAdd(&s_acc[r][0], c);
        if (s != 0.0) atomicAdd(&s_acc[r][1], s);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < a.n_regions) {
    if (MODE == 0 && s_acc[threadIdx.x][0] != 0.0) atomicAdd(out + 2 * threadIdx.x, s_acc[threadIdx.x][0]);
    if (s_acc[threadIdx.x][1] != 0.0) atomicAdd(out + 2 * threadIdx.x + 1, s_acc[threadIdx.x][1]);
  }
}

struct SelectArgs {
  int n_queries;
  int region[kMaxQueries];
  unsigned long long prefix[kMaxQueries];  // key bits decided so far (high bits)
  int shift;                               // the next kRadixBits bits are key >> shift
};

__global__ void __launch_bounds__(256) metrics_hist_kernel(const __grid_constant__ MetricsArgs a,
                                                           const __grid_constant__ SelectArgs s,
                                                           unsigned* hist /*[queries][kBins]*/) {
  __shared__ unsigned s_hist[kMaxQueries][kBins];
  for (int i = threadIdx.x; i < kMaxQueries * kBins; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int hi_shift = s.shift + kRadixBits;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = load_as<double>(a.map, a.map_dtype, i);
    if (!voxel_valid(a, v)) continue;
    const int label = load_label(a, i);
    const unsigned long long key = ordered_key(v);
    const unsigned long long hi = hi_shift >= 64 ? 0ull : key >> hi_shift;
    const unsigned bin = (unsigned)(key >> s.shift) & (kBins - 1);
    for (int q = 0; q < s.n_queries; ++q) {
      if (in_region(a, s.region[q], label) && hi == s.prefix[q]) atomicAdd(&s_hist[q][bin], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < s.n_queries * kBins; i += blockDim.x) {
    const unsigned c = (&s_hist[0][0])[i];
    if (c) atomicAdd(hist + i, c);
  }
}

}  // namespace dfit

using namespace dfit;

extern "C" int dfit_region_metrics_host(dfit_handle* h, int64_t n_vox, const void* map, int map_dtype, const void* labels,
                                        int labels_dtype, int n_regions, const int32_t* region_labels, int has_bounds,
                                        double lb, double ub, int closed_left, int closed_right, double* out) {
  if (!h) return fail(DFIT_ERR_BAD_ARG, "handle is NULL");
  if (n_vox < 0 || !map || !out || !region_labels) return fail(DFIT_ERR_BAD_ARG, "bad argument");
  if (n_regions < 1 || n_regions > kMaxRegions)
    return fail(DFIT_ERR_UNSUPPORTED, "n_regions=%d outside [1, %d]", n_regions, kMaxRegions);
  if (map_dtype != DFIT_F32 && map_dtype != DFIT_F64) return fail(DFIT_ERR_BAD_ARG, "map must be f32 or f64");
  if (labels && (labels_dtype < DFIT_F32 || labels_dtype > DFIT_U8)) return fail(DFIT_ERR_BAD_ARG, "bad labels_dtype");
  CUDA_TRY(cudaSetDevice(h->device));
  Slot& sl = h->slots[0];
  cudaStream_t st = sl.stream;
  const size_t msz = dtype_size(map_dtype), lsz = labels ? dtype_size(labels_dtype) : 0;
  int rc;
  if ((rc = ensure(sl.y, (size_t)n_vox * msz + 16)) != DFIT_OK) return rc;
  if (labels && (rc = ensure(sl.mask, (size_t)n_vox * lsz + 16)) != DFIT_OK) return rc;
  const size_t scratch_bytes = 1024 + kMaxQueries * kBins * sizeof(unsigned);
  if ((rc = ensure(h->scratch, scratch_bytes)) != DFIT_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(sl.y.p, map, (size_t)n_vox * msz, cudaMemcpyHostToDevice, st));
  if (labels) CUDA_TRY(cudaMemcpyAsync(sl.mask.p, labels, (size_t)n_vox * lsz, cudaMemcpyHostToDevice, st));

  MetricsArgs a;
  a.map = sl.y.p;
  a.labels = labels ? sl.mask.p : nullptr;
  a.n = n_vox;
  a.map_dtype = map_dtype;
  a.labels_dtype = labels_dtype;
  a.n_regions = n_regions;
  for (int r = 0; r < kMaxRegions; ++r) a.region_label[r] = r < n_regions ? region_labels[r] : 0;
  a.has_bounds = has_bounds;
  a.closed_left = closed_left;
  a.closed_right = closed_right;
  a.lb = lb;
  a.ub = ub;

  double* d_mom = reinterpret_cast<double*>(h->scratch.p);        // [regions][2]
  double* d_mean = d_mom + 2 * kMaxRegions;                       // [regions]
  unsigned* d_hist = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(h->scratch.p) + 1024);
  const int blocks = h->sm_count * 8;
  double mom[kMaxRegions][2], mean[kMaxRegions], var[kMaxRegions][2];

  CUDA_TRY(cudaMemsetAsync(d_mom, 0, sizeof(mom), st));
  metrics_moments_kernel<0><<<blocks, 256, 0, st>>>(a, nullptr, d_mom);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(mom, d_mom, sizeof(mom), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int r = 0; r < kMaxRegions; ++r) mean[r] = (r < n_regions && mom[r][0] > 0) ? mom[r][1] / mom[r][0] : NAN;
  CUDA_TRY(cudaMemcpyAsync(d_mean, mean, sizeof(mean), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(d_mom, 0, sizeof(mom), st));
  metrics_moments_kernel<1><<<blocks, 256, 0, st>>>(a, d_mean, d_mom);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(var, d_mom, sizeof(var), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));

  // exact median: two order statistics per region (the two middle ranks; equal for odd counts)
  SelectArgs s;
  std::memset(&s, 0, sizeof(s));
  long long rank[kMaxQueries];
  int nq = 0;
  int qmap[kMaxRegions][2];
  for (int r = 0; r < n_regions; ++r) {
    const long long c = (long long)mom[r][0];
    qmap[r][0] = qmap[r][1] = -1;
    if (c <= 0) continue;
    const long long lo = (c - 1) / 2, hi = c / 2;
    qmap[r][0] = nq;
    s.region[nq] = r;
    rank[nq++] = lo;
    if (hi != lo) {
      qmap[r][1] = nq;
      s.region[nq] = r;
      rank[nq++] = hi;
    } else {
      qmap[r][1] = qmap[r][0];
    }
  }
  s.n_queries = nq;
  std::vector<unsigned> hist((size_t)kMaxQueries * kBins);
  for (int shift = 64 - kRadixBits; nq > 0 && shift >= 0; shift -= kRadixBits) {
    s.shift = shift;
    CUDA_TRY(cudaMemsetAsync(d_hist, 0, (size_t)nq * kBins * sizeof(unsigned), st));
    metrics_hist_kernel<<<blocks, 256, 0, st>>>(a, s, d_hist);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(hist.data(), d_hist, (size_t)nq * kBins * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int q = 0; q < nq; ++q) {
      long long left = rank[q];
      int b = 0;
      for (; b < kBins - 1; ++b) {
        const long long c = hist[(size_t)q * kBins + b];
        if (left < c) break;
        left -= c;
      }
      rank[q] = left;
      s.prefix[q] = (s.prefix[q] << kRadixBits) | (unsigned long long)b;
    }
  }
  for (int r = 0; r < n_regions; ++r) {
    const double c = mom[r][0];
    out[4 * r + 0] = c;
    out[4 * r + 1] = c > 0 ? mean[r] : NAN;
    out[4 * r + 2] = c > 0 ? std::sqrt(var[r][1] / c) : NAN;
    // the midpoint of the two middle values is taken in the map's own type, like np.nanmedian does
    const double lo = key_to_double(s.prefix[qmap[r][0]]), hi = key_to_double(s.prefix[qmap[r][1]]);
    const double med = map_dtype == DFIT_F32 ? (double)(0.5f * ((float)lo + (float)hi)) : 0.5 * (lo + hi);
    out[4 * r + 3] = c > 0 ? med : NAN;
  }
  return DFIT_OK;
}
