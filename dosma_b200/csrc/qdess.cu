// qDESS analytic T2 map -- the element-wise companion of the fit path.
//
// Restates the arithmetic of dosma/scan_sequences/mri/qdess.py:225-252 per voxel:
//     ratio = nan_to_num(S2 / S1)                                   (:228-229)
//     t2    = nan_to_num(-2000 (TR - TE) / (log(|ratio| / k) + c1))  (:232-234)
//     t2 outside [lower, upper] -> NaN (:237-239); NaN -> fill (:240-245); around(decimals) (:247-248)
//     optional fat / fluid suppression masks against global maxima (:250-255)
// k and c1 are scalars computed on the host from the sequence parameters (:204-223).
// 8 B read + 4 B written per voxel: a genuinely HBM-bound kernel (float4-vectorised, MUFU log/rcp).
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "dfit_internal.h"

namespace dfit {

struct QdessArgs {
  const void* e1;
  const void* e2;
  void* out;
  int64_t n;
  int in_dtype, out_dtype;
  double k, inv_k, c1, scale;  // scale = -2000 (TR - TE)
  double round_scale;          // 10^decimals
  int has_bounds;
  double lb, ub;
  int has_fill;
  double fill;
  int decimals;
  int suppress_fat, suppress_fluid;
  double beta;
  const double* maxima;  // [0] max(S1), [1] max(S1 - beta S2)  (device)
  int wide_masks;        // suppression masks compared in float64 (any sample type but float32, see qdess_voxel)
};

template <typename T>
__device__ __forceinline__ T nan_to_num_t(T v) {
  const T big = sizeof(T) == 4 ? (T)FLT_MAX : (T)DBL_MAX;
  if (v != v) return (T)0;
  if (v > big) return big;
  if (v < -big) return -big;
  return v;
}

// Exact path (T = double): the reference's own arithmetic -- its `mask = ones(...)` promotes every
// volume to float64 (qdess.py:226-228) -- so results are bit-comparable with numpy.
// The suppression masks (:250-255) are compared in the arithmetic numpy uses for them: float32 echoes stay float32
// (`0.15 * np.float32` is a float32 under numpy >= 2 scalar promotion, and so is `echo_1 - beta * echo_2`), every
// other sample type (int16 DICOM pixels, float64) is promoted to float64.
__device__ __forceinline__ double qdess_voxel(const QdessArgs& a, double s1, double s2, double max1, double maxnf) {
  double ratio = nan_to_num_t<double>(s2 / s1);
  double v = nan_to_num_t<double>(a.scale / (log(fabs(ratio) / a.k) + a.c1));
  if (a.has_bounds && (v < a.lb || v > a.ub)) v = NAN;
  if (a.has_fill && v != v) v = a.fill;
  if (a.decimals >= 0) v = rint(v * a.round_scale) / a.round_scale;
  if (a.wide_masks) {
    if (a.suppress_fat) v = v * (double)(s1 > 0.15 * max1);
    // (numpy rounds beta * S2 before subtracting: no fused multiply-add here -- on integer samples many voxels sit
    // EXACTLY on the threshold and the last bit decides)
    if (a.suppress_fluid) v = v * (double)(__dsub_rn(s1, __dmul_rn(a.beta, s2)) > 0.1 * maxnf);
  } else {
    if (a.suppress_fat) v = v * (double)((float)s1 > 0.15f * (float)max1);
    if (a.suppress_fluid) v = v * (double)(__fsub_rn((float)s1, __fmul_rn((float)a.beta, (float)s2)) > 0.1f * (float)maxnf);
  }
  return v;
}

// Fast path (T = float): fp32 with MUFU reciprocal / log2, ~2e-6 relative before rounding; only the
// final rounding divide is IEEE so that rounded values are the nearest float to the decimal.
__device__ __forceinline__ float qdess_voxel(const QdessArgs& a, float s1, float s2, double max1d, double maxnfd) {
  const float max1 = (float)max1d, maxnf = (float)maxnfd;
  float ratio = nan_to_num_t<float>(__fdividef(s2, s1));
  if (s1 == 0.f) ratio = s2 == 0.f ? 0.f : (s2 > 0.f ? FLT_MAX : -FLT_MAX);  // __fdividef(x, 0) is not IEEE
  float v = nan_to_num_t<float>(__fdividef((float)a.scale, __logf(fabsf(ratio) * (float)a.inv_k) + (float)a.c1));
  if (a.has_bounds && (v < (float)a.lb || v > (float)a.ub)) v = NAN;
  if (a.has_fill && v != v) v = (float)a.fill;
  if (a.decimals >= 0) v = rintf(v * (float)a.round_scale) / (float)a.round_scale;
  if (a.suppress_fat) v = v * (float)(s1 > 0.15f * max1);
  if (a.suppress_fluid) v = v * (float)(__fsub_rn(s1, __fmul_rn((float)a.beta, s2)) > 0.1f * maxnf);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) qdess_kernel(const __grid_constant__ QdessArgs a) {
  const double max1 = a.maxima ? a.maxima[0] : 0.0, maxnf = a.maxima ? a.maxima[1] : 0.0;
  const int64_t n4 = a.n / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec_in = a.in_dtype == DT_F32 &&
                      ((reinterpret_cast<uintptr_t>(a.e1) | reinterpret_cast<uintptr_t>(a.e2) |
                        reinterpret_cast<uintptr_t>(a.out)) & 15) == 0;
  if (vec_in) {  // 4 voxels per thread: 2 x LDG.128 in, STG.128 (fp32 out) or 2 x STG.128 (fp64 out)
    const float4* p1 = reinterpret_cast<const float4*>(a.e1);
    const float4* p2 = reinterpret_cast<const float4*>(a.e2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 u = __ldcs(p1 + i), w = __ldcs(p2 + i);
      const T r0 = qdess_voxel(a, (T)u.x, (T)w.x, max1, maxnf), r1 = qdess_voxel(a, (T)u.y, (T)w.y, max1, maxnf);
      const T r2 = qdess_voxel(a, (T)u.z, (T)w.z, max1, maxnf), r3 = qdess_voxel(a, (T)u.w, (T)w.w, max1, maxnf);
      if (a.out_dtype == DT_F32) {
        __stcs(reinterpret_cast<float4*>(a.out) + i, make_float4((float)r0, (float)r1, (float)r2, (float)r3));
      } else {
        __stcs(reinterpret_cast<double2*>(a.out) + 2 * i, make_double2((double)r0, (double)r1));
        __stcs(reinterpret_cast<double2*>(a.out) + 2 * i + 1, make_double2((double)r2, (double)r3));
      }
    }
  }
  for (int64_t v = (vec_in ? n4 * 4 : 0) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < a.n; v += stride) {
    const T s1 = load_as<T>(a.e1, a.in_dtype, v), s2 = load_as<T>(a.e2, a.in_dtype, v);
    const T r = qdess_voxel(a, s1, s2, max1, maxnf);
    if (a.out_dtype == DT_F32) reinterpret_cast<float*>(a.out)[v] = (float)r;
    else reinterpret_cast<double*>(a.out)[v] = (double)r;
  }
}

// Global maxima for the suppression masks: max(S1) and max(S1 - beta S2), reduced as order-preserving integers
// (in float32 arithmetic for float32 echoes, float64 otherwise -- see qdess_voxel).
__device__ __forceinline__ long long double_to_ordered(double f) {
  const long long i = __double_as_longlong(f);
  return i >= 0 ? i : i ^ 0x7fffffffffffffffll;
}
__device__ __forceinline__ double ordered_to_double(long long i) { return __longlong_as_double(i >= 0 ? i : i ^ 0x7fffffffffffffffll); }

__global__ void __launch_bounds__(256) qdess_max_kernel(const void* e1, const void* e2, int dtype, int64_t n, double beta,
                                                        int wide, long long* ordered /*[2]*/) {
  double m1 = -DBL_MAX, m2 = -DBL_MAX;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    if (wide) {
      const double s1 = load_as<double>(e1, dtype, v), s2 = load_as<double>(e2, dtype, v);
      m1 = fmax(m1, s1);  // (fmax ignores NaN like np.max would not; volumes are finite by construction)
      m2 = fmax(m2, __dsub_rn(s1, __dmul_rn(beta, s2)));  // (product rounded first, like numpy: see qdess_voxel)
    } else {
      const float s1 = load_as<float>(e1, dtype, v), s2 = load_as<float>(e2, dtype, v);
      m1 = fmax(m1, (double)s1);
      m2 = fmax(m2, (double)__fsub_rn(s1, __fmul_rn((float)beta, s2)));
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(ordered + 0, double_to_ordered(m1));
    atomicMax(ordered + 1, double_to_ordered(m2));
  }
}

__global__ void qdess_max_finish(long long* ordered, double* maxima) {
  maxima[0] = ordered_to_double(ordered[0]);
  maxima[1] = ordered_to_double(ordered[1]);
}

static inline long long double_to_ordered_host(double f) {
  long long i;
  memcpy(&i, &f, sizeof(i));
  return i >= 0 ? i : i ^ 0x7fffffffffffffffll;
}

cudaError_t launch_qdess(const QdessArgs& a, int compute_f64, int sm_count, long long* ordered, double* maxima,
                         cudaStream_t stream) {
  QdessArgs args = a;
  args.wide_masks = (compute_f64 && a.in_dtype != DT_F32) ? 1 : 0;
  if (a.suppress_fat || a.suppress_fluid) {
    const long long init[2] = {double_to_ordered_host(-DBL_MAX), double_to_ordered_host(-DBL_MAX)};
    cudaError_t e = cudaMemcpyAsync(ordered, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    qdess_max_kernel<<<sm_count * 8, 256, 0, stream>>>(a.e1, a.e2, a.in_dtype, a.n, a.beta, args.wide_masks, ordered);
    qdess_max_finish<<<1, 1, 0, stream>>>(ordered, maxima);
    args.maxima = maxima;
  } else {
    args.maxima = nullptr;
  }
  const int64_t work = (a.n + 3) / 4;
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (compute_f64) qdess_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(args);
  else qdess_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace dfit

using namespace dfit;

extern "C" {

int dfit_default_qdess_opts(dfit_qdess_opts* o) {
  if (!o) return fail(DFIT_ERR_BAD_ARG, "opts is NULL");
  memset(o, 0, sizeof(*o));
  o->struct_size = (int32_t)sizeof(dfit_qdess_opts);
  o->k = 1.0;
  o->has_bounds = 1;  // qdess.py:119 nan_bounds = (0, 100)
  o->lb = 0.0;
  o->ub = 100.0;
  o->has_nan_fill = 1;  // qdess.py:120 nan_to_num = 0.0
  o->nan_fill = 0.0;
  o->decimals = 1;  // qdess.py:121
  o->beta = 1.2;    // qdess.py:110
  o->compute_dtype = -1;
  return DFIT_OK;
}

static int qdess_args(const dfit_qdess_opts* o, int64_t n, int in_dtype, int out_dtype, QdessArgs& a, int& f64) {
  if (!o || o->struct_size != (int32_t)sizeof(dfit_qdess_opts))
    return fail(DFIT_ERR_BAD_ARG, "bad dfit_qdess_opts (use dfit_default_qdess_opts)");
  if (n < 0) return fail(DFIT_ERR_BAD_ARG, "n_vox < 0");
  if (in_dtype < DFIT_F32 || in_dtype > DFIT_U8) return fail(DFIT_ERR_BAD_ARG, "bad in_dtype");
  if (out_dtype != DFIT_F32 && out_dtype != DFIT_F64) return fail(DFIT_ERR_BAD_ARG, "bad out_dtype");
  a.n = n;
  a.in_dtype = in_dtype;
  a.out_dtype = out_dtype;
  a.k = o->k;
  a.inv_k = 1.0 / o->k;
  a.round_scale = 1.0;
  for (int d = 0; d < o->decimals; ++d) a.round_scale *= 10.0;
  a.c1 = o->c1;
  a.scale = -2000.0 * o->tr_minus_te;
  a.has_bounds = o->has_bounds;
  a.lb = o->lb;
  a.ub = o->ub;
  a.has_fill = o->has_nan_fill;
  a.fill = o->nan_fill;
  a.decimals = o->decimals;
  a.suppress_fat = o->suppress_fat;
  a.suppress_fluid = o->suppress_fluid;
  a.beta = o->beta;
  a.maxima = nullptr;
  a.wide_masks = 0;
  // default: the reference's float64 arithmetic (exact parity); DFIT_F32 selects the HBM-bound fast path
  f64 = o->compute_dtype != DFIT_F32;
  return DFIT_OK;
}

int dfit_qdess_t2_device(dfit_handle* h, const dfit_qdess_opts* opts, int64_t n_vox, const void* echo1, const void* echo2,
                         int in_dtype, void* t2, int out_dtype, void* stream) {
  if (!h) return fail(DFIT_ERR_BAD_ARG, "handle is NULL");
  QdessArgs a;
  int f64 = 0;
  int rc = qdess_args(opts, n_vox, in_dtype, out_dtype, a, f64);
  if (rc != DFIT_OK) return rc;
  if (n_vox == 0) return DFIT_OK;
  if (!echo1 || !echo2 || !t2) return fail(DFIT_ERR_BAD_ARG, "NULL buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  if ((rc = ensure(h->scratch, 64)) != DFIT_OK) return rc;
  a.e1 = echo1;
  a.e2 = echo2;
  a.out = t2;
  CUDA_TRY(launch_qdess(a, f64, h->sm_count, reinterpret_cast<long long*>(h->scratch.p),
                        reinterpret_cast<double*>(h->scratch.p) + 4, (cudaStream_t)stream));
  return DFIT_OK;
}

int dfit_qdess_t2_host(dfit_handle* h, const dfit_qdess_opts* opts, int64_t n_vox, const void* echo1, const void* echo2,
                       int in_dtype, void* t2, int out_dtype) {
  if (!h) return fail(DFIT_ERR_BAD_ARG, "handle is NULL");
  QdessArgs a;
  int f64 = 0;
  int rc = qdess_args(opts, n_vox, in_dtype, out_dtype, a, f64);
  if (rc != DFIT_OK) return rc;
  if (n_vox == 0) return DFIT_OK;
  if (!echo1 || !echo2 || !t2) return fail(DFIT_ERR_BAD_ARG, "NULL buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t isz = dtype_size(in_dtype), osz = out_dtype == DFIT_F32 ? 4 : 8;
  Slot& sl = h->slots[0];
  if ((rc = ensure(sl.y, 2 * (size_t)n_vox * isz + 32)) != DFIT_OK) return rc;
  if ((rc = ensure(sl.popt, (size_t)n_vox * osz)) != DFIT_OK) return rc;
  if ((rc = ensure(h->scratch, 64)) != DFIT_OK) return rc;
  cudaStream_t st = sl.stream;
  char* d1 = static_cast<char*>(sl.y.p);
  char* d2 = d1 + (((size_t)n_vox * isz + 15) / 16) * 16;
  CUDA_TRY(cudaMemcpyAsync(d1, echo1, (size_t)n_vox * isz, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d2, echo2, (size_t)n_vox * isz, cudaMemcpyHostToDevice, st));
  a.e1 = d1;
  a.e2 = d2;
  a.out = sl.popt.p;
  CUDA_TRY(launch_qdess(a, f64, h->sm_count, reinterpret_cast<long long*>(h->scratch.p),
                        reinterpret_cast<double*>(h->scratch.p) + 4, st));
  CUDA_TRY(cudaMemcpyAsync(t2, sl.popt.p, (size_t)n_vox * osz, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DFIT_OK;
}

}  // extern "C"
