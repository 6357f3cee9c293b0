// C-ABI of libdfit.so (include/dfit.h): handles, streams, chunked host<->device pipeline, dispatch.
// Device code only -- there is deliberately no host implementation of the fit in this library.
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <pthread.h>
#include <sched.h>

#include <cctype>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <thread>
#include <vector>

#include "dfit_internal.h"

using namespace dfit;

namespace dfit {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const char* last_error() { return g_err; }

int ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return DFIT_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + (bytes >> 3) + 256;
  CUDA_TRY(cudaMalloc(&b.p, want));
  b.cap = want;
  return DFIT_OK;
}

int ensure_host(HostBuf& b, size_t bytes, const cpu_set_t* cpus) {
  if (bytes <= b.cap) return DFIT_OK;
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + (bytes >> 3) + 256;
  if (cpus) {
    // allocated (and first touched) by a thread on the GPU's NUMA node, so that the pages live next to the PCIe root
    cudaError_t err = cudaSuccess;
    int device = 0;
    cudaGetDevice(&device);
    void* p = nullptr;
    std::thread t([&] {
      pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), cpus);
      cudaSetDevice(device);
      err = cudaHostAlloc(&p, want, cudaHostAllocDefault);
      if (err == cudaSuccess) std::memset(p, 0, want);
    });
    t.join();
    CUDA_TRY(err);
    b.p = p;
  } else {
    CUDA_TRY(cudaHostAlloc(&b.p, want, cudaHostAllocDefault));
  }
  b.cap = want;
  return DFIT_OK;
}

}  // namespace dfit

namespace {

// Copy `count` segments {dst, src, bytes} with a few host threads (1 MiB blocks, static partition).  A single
// core moves ~10 GB/s; the pageable side of the host entry point needs several to keep up with PCIe.
struct CopySeg {
  void* dst;
  const void* src;
  size_t bytes;
};

// Worker threads of the host-side copies run on the CPUs next to the GPU (set per call by the entry points).
thread_local const cpu_set_t* g_worker_cpus = nullptr;

inline void bind_worker(const cpu_set_t* cpus) {
  if (cpus) pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), cpus);
}

// A few persistent host threads for the staging copies (a chunk is filled / drained in ~1 ms: spawning 15 threads
// per copy would cost as much again).  run(nt, cpus, work) executes work(0) on the caller and work(1..nt-1) on the
// pool's threads, bound to `cpus`, and returns when all are done.  One job at a time (concurrent callers queue up).
class HostPool {
 public:
  static HostPool& get() {
    static HostPool* p = new HostPool();  // (never destroyed: its threads may outlive static destruction order)
    return *p;
  }
  int size() const { return (int)threads_.size() + 1; }
  void run(int nt, const cpu_set_t* cpus, const std::function<void(int)>& work) {
    if (nt <= 1) {
      work(0);
      return;
    }
    std::lock_guard<std::mutex> job(job_mu_);
    {
      std::lock_guard<std::mutex> lk(mu_);
      work_ = &work;
      cpus_ = cpus;
      nt_ = nt;
      pending_ = nt - 1;
      ++epoch_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    work_ = nullptr;
  }

 private:
  HostPool() {
    unsigned hw = std::thread::hardware_concurrency();
    const int n = (int)(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
    for (int t = 1; t < n; ++t) threads_.emplace_back([this, t] { loop(t); });
    for (auto& th : threads_) th.detach();
  }
  void loop(int t) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* work = nullptr;
      const cpu_set_t* cpus = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (t >= nt_) continue;  // not part of this job
        work = work_;
        cpus = cpus_;
      }
      bind_worker(cpus);
      (*work)(t);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_, job_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* work_ = nullptr;
  const cpu_set_t* cpus_ = nullptr;
  int nt_ = 0, pending_ = 0;
  unsigned long long epoch_ = 0;
};

// memcpy with streaming (non-temporal) stores: the destination -- a staging block about to be DMA'd, or a result array
// the caller reads later -- is not wanted in the cache, and a regular store would first READ every destination line
// (read for ownership): a third of the memory traffic of these copies, which are bandwidth-bound.
inline void stream_copy(void* dst, const void* src, size_t bytes) {
#if defined(__SSE2__)
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
  if (bytes < 256 + head) {
    std::memcpy(d, s, bytes);
    return;
  }
  std::memcpy(d, s, head);
  d += head;
  s += head;
  bytes -= head;
  const size_t blocks = bytes / 64;
  for (size_t i = 0; i < blocks; ++i, d += 64, s += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
  }
  std::memcpy(d, s, bytes - blocks * 64);
  _mm_sfence();
#else
  std::memcpy(dst, src, bytes);
#endif
}

void par_copy(const CopySeg* segs, int count) {
  const cpu_set_t* cpus = g_worker_cpus;
  constexpr size_t kBlk = (size_t)1 << 20;
  size_t total_blocks = 0;
  for (int i = 0; i < count; ++i) total_blocks += (segs[i].bytes + kBlk - 1) / kBlk;
  int nt = HostPool::get().size();
  if ((size_t)nt > total_blocks) nt = (int)total_blocks;
  auto work = [&](int t) {
    size_t b = 0;  // global block counter; thread t takes blocks with b % nt == t
    for (int i = 0; i < count; ++i) {
      const size_t nb = (segs[i].bytes + kBlk - 1) / kBlk;
      for (size_t k = 0; k < nb; ++k, ++b) {
        if ((int)(b % (size_t)nt) != t) continue;
        const size_t off = k * kBlk, len = segs[i].bytes - off < kBlk ? segs[i].bytes - off : kBlk;
        stream_copy((char*)segs[i].dst + off, (const char*)segs[i].src + off, len);
      }
    }
  };
  if (total_blocks == 0) return;
  HostPool::get().run(nt, cpus, work);
}

// Generic static partition of [0, n) over a few host threads.
template <class F>
void par_for(int64_t n, int64_t grain, F fn) {
  const cpu_set_t* cpus = g_worker_cpus;
  int nt = HostPool::get().size();
  const int64_t nb = (n + grain - 1) / grain;
  if (nt > nb) nt = (int)nb;
  if (n <= 0) return;
  auto work = [&](int t) {
    const int64_t b0 = nb * t / nt, b1 = nb * (t + 1) / nt;
    const int64_t lo = b0 * grain, hi = b1 * grain < n ? b1 * grain : n;
    if (lo < hi) fn(lo, hi);
  };
  HostPool::get().run(nt < 1 ? 1 : nt, cpus, work);
}

// float32 staging [n, ncols] -> the caller's float64 map.  scale[c] > 0: column c was rounded to a decimal grid on the
// device (v = fl32(m / scale), m an integer below 2^22): it is put back onto the float64 grid with numpy.around's own
// formula, rint(v scale) / scale, which recovers m exactly; other columns (and NaN / inf) are widened as they are.
void par_widen(double* dst, const float* src, int64_t n, int ncols, const double* scale) {
  // flat element range [lo, hi) of the (n, ncols) map; element k belongs to column k % ncols
  par_for(n * ncols, (int64_t)1 << 18, [&](int64_t lo, int64_t hi) {
    auto one = [&](int64_t k) {
      const double v = (double)src[k], s = scale[ncols == 1 ? 0 : k % ncols];
      return s > 0 ? std::nearbyint(v * s) / s : v;
    };
#if defined(__SSE2__)
    // streaming stores, two doubles at a time (see stream_copy); the pair (k, k + 1) starts on a 16-byte boundary
    int64_t k = lo;
    while (k < hi && (reinterpret_cast<uintptr_t>(dst + k) & 15)) {
      dst[k] = one(k);
      ++k;
    }
    bool any_scale = false;
    for (int c = 0; c < ncols; ++c) any_scale = any_scale || scale[c] > 0;
    if (!any_scale) {
      for (; k + 4 <= hi; k += 4) {
        const __m128 f = _mm_loadu_ps(src + k);
        _mm_stream_pd(dst + k, _mm_cvtps_pd(f));
        _mm_stream_pd(dst + k + 2, _mm_cvtps_pd(_mm_movehl_ps(f, f)));
      }
    }
    for (; k + 2 <= hi; k += 2) _mm_stream_pd(dst + k, _mm_set_pd(one(k + 1), one(k)));
    for (; k < hi; ++k) dst[k] = one(k);
    _mm_sfence();
#else
    for (int64_t k = lo; k < hi; ++k) dst[k] = one(k);
#endif
  });
}

int model_nparams(int model) {
  switch (model) {
    case DFIT_MODEL_MONOEXP: return 2;
    case DFIT_MODEL_BIEXP: return 4;
    case DFIT_MODEL_LINEAR: return 1;
    default: return -1;
  }
}

}  // namespace

namespace {

int validate(const dfit_opts* o, int n_echo, int64_t n_vox, const double* x, int y_dtype, int p0_dtype, int out_dtype,
             bool have_p0v) {
  if (!o) return fail(DFIT_ERR_BAD_ARG, "opts is NULL");
  if (o->struct_size != (int32_t)sizeof(dfit_opts))
    return fail(DFIT_ERR_BAD_ARG, "opts->struct_size=%d, expected %d (use dfit_default_opts)", o->struct_size,
                (int)sizeof(dfit_opts));
  const int P = model_nparams(o->model);
  if (P < 0) return fail(DFIT_ERR_BAD_ARG, "unknown model %d", o->model);
  if (n_echo < 1 || n_echo > DFIT_MAX_ECHOES)
    return fail(DFIT_ERR_UNSUPPORTED, "n_echo=%d outside [1, %d]", n_echo, DFIT_MAX_ECHOES);
  if (n_echo < P) return fail(DFIT_ERR_BAD_ARG, "n_echo=%d < number of parameters %d", n_echo, P);
  if (n_vox < 0) return fail(DFIT_ERR_BAD_ARG, "n_vox < 0");
  if (!x) return fail(DFIT_ERR_BAD_ARG, "x is NULL");
  if (y_dtype < DFIT_F32 || y_dtype > DFIT_U8) return fail(DFIT_ERR_BAD_ARG, "bad y_dtype %d", y_dtype);
  if (out_dtype != DFIT_F32 && out_dtype != DFIT_F64) return fail(DFIT_ERR_BAD_ARG, "bad out_dtype %d", out_dtype);
  if (have_p0v && p0_dtype != DFIT_F32 && p0_dtype != DFIT_F64) return fail(DFIT_ERR_BAD_ARG, "bad p0_dtype");
  if (o->compute_dtype != DFIT_F32 && o->compute_dtype != DFIT_F64)
    return fail(DFIT_ERR_BAD_ARG, "bad compute_dtype %d", o->compute_dtype);
  if (o->init_mode == DFIT_INIT_LOGLINEAR && o->model != DFIT_MODEL_MONOEXP)
    return fail(DFIT_ERR_UNSUPPORTED, "log-linear initialisation is defined for the mono-exponential model only");
  if (o->init_mode != DFIT_INIT_GIVEN && o->init_mode != DFIT_INIT_LOGLINEAR)
    return fail(DFIT_ERR_BAD_ARG, "bad init_mode %d", o->init_mode);
  for (int i = 0; i < P; ++i)
    if (std::isnan(o->p0[i]) && !have_p0v && o->init_mode == DFIT_INIT_GIVEN)
      return fail(DFIT_ERR_BAD_ARG, "p0[%d] is NaN (per-voxel) but p0_voxel is NULL", i);
  if (o->maxfev < 1) return fail(DFIT_ERR_BAD_ARG, "maxfev < 1");
  if (o->out_param >= P) return fail(DFIT_ERR_BAD_ARG, "out_param=%d but the model has %d parameters", o->out_param, P);
  return DFIT_OK;
}

void make_desc(const dfit_opts* o, int n_echo, int64_t n_vox, const double* x, LaunchDesc& d) {
  const int P = model_nparams(o->model);
  const bool f32 = o->compute_dtype == DFIT_F32;
  const double eps = f32 ? 1.1920929e-7 : 2.220446049250313e-16;
  d.model = o->model;
  d.compute_dtype = o->compute_dtype;
  d.n_echo = n_echo;
  d.n_vox = n_vox;
  d.x = x;
  d.p0_voxel_bits = 0;
  for (int i = 0; i < 4; ++i) {
    d.p0s[i] = i < P ? o->p0[i] : 0.0;
    if (i < P && std::isnan(o->p0[i])) {
      d.p0_voxel_bits |= 1u << i;
      d.p0s[i] = 1.0;
    }
  }
  // Engine tolerances: the reference's ftol bounds MINPACK's *last* relative cost decrease, which
  // leaves its iterate up to ~1e-4 (relative) away from the minimiser on noisy data.  The engine
  // iterates ftol_scale (two decades by default) further so that it lands within rtol 1e-4 of wherever
  // MINPACK stopped.
  const double scale = o->ftol_scale > 0 ? o->ftol_scale : 1e-2;
  double ftol = o->ftol * scale;
  const double ftol_floor = f32 ? 1e-8 : 1e-14;
  d.ftol = ftol < ftol_floor ? ftol_floor : ftol;
  d.xtol = o->xtol > 0 ? o->xtol : (f32 ? 1e-6 : 1e-10);
  d.lambda0 = o->lambda0 > 0 ? o->lambda0 : 1e-3;
  d.floor_rel = (8 * eps) * (8 * eps);
  d.r2_eps = o->r2_eps;
  d.y_lo = o->y_lo;
  d.y_hi = o->y_hi;
  d.maxfev = o->maxfev;
  d.init_mode = o->init_mode;
  d.init_linear = o->init_linear < 0 ? (o->model == DFIT_MODEL_BIEXP ? 0 : 1) : o->init_linear;
  // The fast path spends at most kMonoFastPasses passes; a budget too small for that means the caller is
  // probing maxfev behaviour (fitting.py:761), which only the LM reproduces.
  d.fast_path = (o->model == DFIT_MODEL_MONOEXP && o->fast_path != 0 && o->maxfev >= 3 * (kMonoFastPasses + 1))
                    ? (o->fast_path == 2 ? 2 : 1)
                    : 0;
  d.po.enabled = o->post_enabled;
  for (int i = 0; i < 4; ++i) {
    d.po.ufunc[i] = o->ufunc[i];
    d.po.lb[i] = o->lb[i];
    d.po.ub[i] = o->ub[i];
    d.po.decimals[i] = o->decimals[i];
  }
  d.po.has_r2_thresh = o->has_r2_threshold;
  d.po.r2_thresh = o->r2_threshold;
  d.po.has_fill = o->has_nan_fill;
  d.po.fill = o->nan_fill;
  set_post_scales(d.po);  // (derived fp32 thresholds / plans: after every field they are derived from)
  d.mask_fill = o->has_nan_fill ? o->nan_fill : std::numeric_limits<double>::quiet_NaN();
  d.use_tma = o->use_tma;
  d.sel = o->out_param < 0 ? -1 : o->out_param;
  d.tmap = nullptr;
  d.tmap2 = nullptr;
  d.sm_count = 148;
  d.index = nullptr;
  d.index_count = nullptr;
  d.lm_list = nullptr;
  d.lm_head = nullptr;
  d.lm_parity = nullptr;
  d.g = GatherArgs{};
}

// LM tail of the dense two-voxel TMA kernel (see KernelArgs::lm_list): device scratch for the list of voxels that go
// to the LM and its two alternating counters.  DFIT_LM_TAIL=0 keeps the LM inside the kernel (A/B runs).
int lm_tail_prepare(DevBuf& buf, int& parity, int64_t n_vox, cudaStream_t st, LaunchDesc& d) {
  if (const char* e = std::getenv("DFIT_LM_TAIL")) {
    if (e[0] == '0') return DFIT_OK;
  }
  // (small launches keep the LM inside the kernel: a second launch costs them a few microseconds -- 10 % of the 64 x 64 x 16
  // volume's 0.04 ms -- and at that size a divergent warp costs nothing)
  if (d.tmap2 == nullptr || d.g.world != 0 || n_vox >= ((int64_t)1 << 32) || n_vox < ((int64_t)1 << 20)) return DFIT_OK;
  const size_t need = 16 + (size_t)n_vox * sizeof(unsigned);
  if (need > buf.cap) {
    const int rc = ensure(buf, need);
    if (rc != DFIT_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(buf.p, 0, 16, st));
    parity = 0;
  }
  d.lm_head = reinterpret_cast<unsigned*>(buf.p);
  d.lm_list = d.lm_head + 4;
  d.lm_parity = &parity;
  return DFIT_OK;
}

// May the result maps of this fit cross PCIe as float32 and be widened on the host?  Yes for fp32 arithmetic into
// float64 maps, unless a rounded parameter could exceed what a float32 carries exactly on its decimal grid
// (|value| 10^d < 2^22 needs finite bounds).  scale[i] = 10^decimals[i] for rounded parameters, else 0.
bool widen_plan(const dfit_opts* o, int out_dtype, double (&scale)[DFIT_MAX_PARAMS]) {
  for (int i = 0; i < DFIT_MAX_PARAMS; ++i) scale[i] = 0.0;
  if (out_dtype != DFIT_F64 || o->compute_dtype != DFIT_F32) return false;
  if (const char* e = std::getenv("DFIT_HOST_WIDEN")) {
    if (e[0] == '0') return false;
  }
  if (!o->post_enabled) return true;
  const int P = model_nparams(o->model);
  for (int i = 0; i < P; ++i) {
    if (o->decimals[i] < 0) continue;
    if (o->decimals[i] > 10) return false;
    double s = 1.0;
    for (int k = 0; k < o->decimals[i]; ++k) s *= 10.0;
    const double m = std::fmax(std::fabs(o->lb[i]), std::fabs(o->ub[i]));
    if (!(m * s < 4.0e6)) return false;
    if (o->has_nan_fill && !(std::fabs(o->nan_fill) * s < 4.0e6)) return false;
    scale[i] = s;
  }
  return true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  }
  return fn;
}

// 2-D tensor map over planar fp32 samples: dim0 = voxels (contiguous), dim1 = echoes (pitch ld).
// Box = box_vox voxels x E echoes; out-of-range voxels of the last tile are zero-filled.
bool make_sample_tmap(CUtensorMap* map, const void* y, int n_echo, int64_t n_vox, int64_t ld, int box_vox = kTmaTile,
                      int y_dtype = DFIT_F32) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const size_t esz = dtype_size(y_dtype);
  if ((reinterpret_cast<uintptr_t>(y) & 15) != 0 || (ld * esz) % 16 != 0) return false;
  cuuint64_t dims[2] = {(cuuint64_t)n_vox, (cuuint64_t)n_echo};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_vox, (cuuint32_t)n_echo};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = y_dtype == DFIT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
  return fn(map, dt, 2, const_cast<void*>(y), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The TMA-staged kernel is used when asked for (use_tma = 1) and the samples qualify.
bool tma_eligible(const LaunchDesc& d) {
  return d.use_tma == 1 && d.compute_dtype == DFIT_F32 && d.y_dtype == DFIT_F32 && d.layout == DFIT_PLANAR &&
         d.n_echo <= 16 && d.n_vox < (int64_t)1 << 31;
}

// The two-voxels-per-lane mono-exponential kernel takes its tiles through TMA unless told not to
// (use_tma = 0); launch_one checks the rest of its conditions.
bool tma2_eligible(const LaunchDesc& d) {
  return d.use_tma != 0 && d.model == DFIT_MODEL_MONOEXP && d.fast_path == 1 && d.compute_dtype == DFIT_F32 &&
         (d.y_dtype == DFIT_F32 || d.y_dtype == DFIT_I16 || d.y_dtype == DFIT_U16) && d.layout == DFIT_PLANAR &&
         d.mask == nullptr &&
         d.n_echo >= 3 && d.n_echo <= 16 && d.n_vox < ((int64_t)1 << 31) - 2 * kM2Tile;  // 32-bit tile arithmetic
}

cudaError_t dispatch(const LaunchDesc& d) {
  const bool f32 = d.compute_dtype == DFIT_F32;
  switch (d.model) {
    case DFIT_MODEL_MONOEXP: return f32 ? launch_mono_f32(d) : launch_mono_f64(d);
    case DFIT_MODEL_BIEXP: return f32 ? launch_biexp_f32(d) : launch_biexp_f64(d);
    default: return f32 ? launch_linear_f32(d) : launch_linear_f64(d);
  }
}

bool is_pinned_or_device(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace

extern "C" {

int dfit_version(void) { return DFIT_VERSION; }

int dfit_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* dfit_strerror(int code) {
  switch (code) {
    case DFIT_OK: return "ok";
    case DFIT_ERR_BAD_ARG: return "bad argument";
    case DFIT_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU implementation)";
    case DFIT_ERR_CUDA: return "CUDA error";
    case DFIT_ERR_OOM: return "out of device memory";
    case DFIT_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

const char* dfit_last_error(void) { return dfit::last_error(); }

int dfit_model_nparams(int model) { return model_nparams(model); }

int dfit_default_opts(dfit_opts* o, int model) {
  if (!o) return fail(DFIT_ERR_BAD_ARG, "opts is NULL");
  if (model_nparams(model) < 0) return fail(DFIT_ERR_BAD_ARG, "unknown model %d", model);
  std::memset(o, 0, sizeof(*o));
  o->struct_size = (int32_t)sizeof(dfit_opts);
  o->model = model;
  o->compute_dtype = DFIT_F32;
  o->init_mode = DFIT_INIT_GIVEN;
  o->init_linear = -1;
  o->maxfev = 100;      // fitting.py:761
  o->ftol = 1e-5;       // fitting.py:762
  o->ftol_scale = 1e-2;
  o->xtol = 0;
  o->lambda0 = 0;
  o->r2_eps = 1e-8;     // fitting.py:763
  o->y_lo = -std::numeric_limits<double>::infinity();
  o->y_hi = std::numeric_limits<double>::infinity();
  for (int i = 0; i < DFIT_MAX_PARAMS; ++i) {
    o->p0[i] = 1.0;  // SciPy's default when p0 is None (fitting.py:820-826)
    o->ufunc[i] = DFIT_UFUNC_NONE;
    o->lb[i] = -std::numeric_limits<double>::infinity();
    o->ub[i] = std::numeric_limits<double>::infinity();
    o->decimals[i] = -1;
  }
  o->fast_path = -1;
  o->use_tma = -1;
  o->out_param = -1;
  return DFIT_OK;
}

int dfit_create(int device, dfit_handle** out) {
  if (!out) return fail(DFIT_ERR_BAD_ARG, "out is NULL");
  *out = nullptr;
  int n = dfit_device_count();
  if (n <= 0) return fail(DFIT_ERR_NO_DEVICE, "no CUDA device visible; libdfit has no CPU path");
  if (device < 0 || device >= n) return fail(DFIT_ERR_BAD_ARG, "device %d out of range [0, %d)", device, n);
  CUDA_TRY(cudaSetDevice(device));
  dfit_handle* h = new (std::nothrow) dfit_handle();
  if (!h) return fail(DFIT_ERR_OOM, "host allocation failed");
  h->device = device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    delete h;
    return fail(DFIT_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  }
  h->sm_count = prop.multiProcessorCount;
  {
    const char* off = std::getenv("DFIT_HOST_AFFINITY");
    char bus[32] = "";
    if (!(off && off[0] == '0') && cudaDeviceGetPCIBusId(bus, sizeof(bus), device) == cudaSuccess) {
      for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
      char path[128];
      std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
      if (FILE* f = std::fopen(path, "r")) {
        char line[4096];
        if (std::fgets(line, sizeof(line), f)) {
          CPU_ZERO(&h->local_cpus);
          int n_set = 0;
          for (char* tok = std::strtok(line, ",\n"); tok; tok = std::strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            const int k = std::sscanf(tok, "%d-%d", &a, &b);
            if (k == 1) b = a;
            if (k >= 1)
              for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
                CPU_SET(c, &h->local_cpus);
                ++n_set;
              }
          }
          // only CPUs this process may run on; none left (cgroup elsewhere): no binding
          cpu_set_t allowed;
          if (n_set > 0 && sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            CPU_AND(&h->local_cpus, &h->local_cpus, &allowed);
            h->have_local_cpus = CPU_COUNT(&h->local_cpus) > 0;
          }
        }
        std::fclose(f);
      }
    }
    cudaGetLastError();
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (int s = 0; s < kSlots; ++s) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->slots[s].stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->slots[s].ev_in, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->slots[s].ev_out, cudaEventDisableTiming));
  }
  CUDA_TRY(cudaMalloc(&h->counters, kStatSlots * CNT_COUNT * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(h->counters, 0, kStatSlots * CNT_COUNT * sizeof(unsigned long long)));
  CUDA_TRY(cudaEventCreate(&h->ev_start));
  CUDA_TRY(cudaEventCreate(&h->ev_stop));
  *out = h;
  return DFIT_OK;
}

int dfit_destroy(dfit_handle* h) {
  if (!h) return DFIT_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int s = 0; s < kSlots; ++s) {
    Slot& sl = h->slots[s];
    DevBuf* bufs[] = {&sl.y, &sl.mask, &sl.p0, &sl.popt, &sl.r2, &sl.status, &sl.niter, &sl.index, &sl.lm};
    for (DevBuf* b : bufs)
      if (b->p) cudaFree(b->p);
    HostBuf* hbufs[] = {&sl.hin, &sl.hpopt, &sl.hr2};
    for (HostBuf* b : hbufs)
      if (b->p) cudaFreeHost(b->p);
    if (sl.ev_in) cudaEventDestroy(sl.ev_in);
    if (sl.ev_out) cudaEventDestroy(sl.ev_out);
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->counters) cudaFree(h->counters);
  if (h->scratch.p) cudaFree(h->scratch.p);
  if (h->index_buf.p) cudaFree(h->index_buf.p);
  if (h->lm_buf.p) cudaFree(h->lm_buf.p);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->ev_stop) cudaEventDestroy(h->ev_stop);
  delete h;
  return DFIT_OK;
}

int dfit_fit_device(dfit_handle* h, const dfit_opts* opts, int n_echo, int64_t n_vox, const double* x, const void* y,
                    int y_dtype, int y_layout, int64_t ld, const uint8_t* mask, const void* p0_voxel, int p0_dtype,
                    void* popt, void* r2, int out_dtype, uint8_t* status, uint8_t* niter, void* stream) {
  if (!h) return fail(DFIT_ERR_BAD_ARG, "handle is NULL");
  int rc = validate(opts, n_echo, n_vox, x, y_dtype, p0_dtype, out_dtype, p0_voxel != nullptr);
  if (rc != DFIT_OK) return rc;
  if (y_layout != DFIT_PLANAR && y_layout != DFIT_ECHO_FASTEST) return fail(DFIT_ERR_BAD_ARG, "bad layout");
  const bool gathering = h->g.world > 0;
  if (n_vox > 0 && !y) return fail(DFIT_ERR_BAD_ARG, "y must not be NULL");
  if (n_vox > 0 && (!popt != !r2)) return fail(DFIT_ERR_BAD_ARG, "popt and r2 must both be given or both be NULL");
  if (n_vox > 0 && !popt && !gathering) return fail(DFIT_ERR_BAD_ARG, "popt/r2 may only be NULL with dfit_set_gather");
  if (gathering && h->g.row0 + n_vox > h->gather_rows) return fail(DFIT_ERR_BAD_ARG, "row0 + n_vox exceeds the rows of the gather maps");
  if (gathering && opts->compute_dtype != DFIT_F32)
    return fail(DFIT_ERR_UNSUPPORTED, "the fused all-gather map is fp32: use compute_dtype = DFIT_F32 with dfit_set_gather");
  if (gathering && h->g.split_list && (!mask || p0_voxel))
    return fail(DFIT_ERR_BAD_ARG, "split_list gathers need a mask (of the whole volume) and a scalar initial guess");
  if (gathering && (h->g.cols >> model_nparams(opts->model)) != 0u)
    return fail(DFIT_ERR_BAD_ARG, "gather param_mask names a parameter the model does not have");
  // (split-list gathers: y holds only this rank's span of the volume, its pitch is the caller's business)
  if (!(gathering && h->g.split_list) && ld < (y_layout == DFIT_PLANAR ? n_vox : (int64_t)n_echo))
    return fail(DFIT_ERR_BAD_ARG, "ld too small");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the legacy default stream (what torch calls its default stream)
  LaunchDesc d;
  make_desc(opts, n_echo, n_vox, x, d);
  d.y = y;
  d.y_dtype = y_dtype;
  d.layout = y_layout;
  d.ld = ld;
  d.mask = mask;
  d.p0v = p0_voxel;
  d.p0_dtype = p0_dtype;
  if (!p0_voxel) d.p0_voxel_bits = 0;
  d.popt = popt;
  d.r2 = r2;
  d.out_dtype = out_dtype;
  d.status = status;
  d.niter = niter;
  d.counters = h->counters;
  d.stream = st;
  if (gathering) {
    d.g = h->g;
    if (d.g.cols == 0u) d.g.cols = (1u << model_nparams(opts->model)) - 1u;  // all parameters
    d.g.ncols = __builtin_popcount(d.g.cols) + 1;
  }
  CUDA_TRY(cudaMemsetAsync(h->counters, 0, kStatSlots * CNT_COUNT * sizeof(unsigned long long), st));
  CUDA_TRY(cudaEventRecord(h->ev_start, st));
  h->last_launches = 0;
  CUtensorMap tmap;
  d.sm_count = h->sm_count;
  if (mask != nullptr && n_vox > 0) {
    if (n_vox >= ((int64_t)1 << 32)) return fail(DFIT_ERR_UNSUPPORTED, "masked fits are limited to 2^32 voxels per call");
    if ((rc = ensure(h->index_buf, (size_t)n_vox * sizeof(unsigned) + 16)) != DFIT_OK) return rc;
    d.index_count = reinterpret_cast<unsigned*>(h->index_buf.p);
    d.index = d.index_count + 4;
  }
  CUtensorMap tmap2;
  if (n_vox > 0 && tma2_eligible(d) && make_sample_tmap(&tmap2, y, n_echo, n_vox, ld, kM2Tile, y_dtype)) d.tmap2 = &tmap2;
  if (n_vox > 0 && tma_eligible(d)) {
    if (!make_sample_tmap(&tmap, y, n_echo, n_vox, ld))
      return fail(DFIT_ERR_UNSUPPORTED, "use_tma=1 but the samples do not qualify (16-byte aligned base and pitch)");
    d.tmap = &tmap;
  } else if (opts->use_tma == 1 && d.tmap2 == nullptr) {
    return fail(DFIT_ERR_UNSUPPORTED,
                "use_tma=1 needs planar fp32 (mono-exponential fast path: also 16-bit) samples with a 16-byte aligned "
                "base and pitch, fp32 arithmetic and <= 16 echoes");
  }
  if (n_vox > 0) {
    if ((rc = lm_tail_prepare(h->lm_buf, h->lm_parity, n_vox, st, d)) != DFIT_OK) return rc;
    const int par0 = h->lm_parity;
    CUDA_TRY(dispatch(d));
    h->last_launches = (mask != nullptr && d.tmap == nullptr ? 2 : 1) + (h->lm_parity != par0 ? 1 : 0);
  }
  CUDA_TRY(cudaEventRecord(h->ev_stop, st));
  h->ev_valid = true;
  h->ev_stream = st;
  h->last_n = n_vox;
  h->last_total_ms = -1.f;
  h->host_kernel_ms = -1.f;
  return DFIT_OK;
}

static int fit_host_impl(dfit_handle* h, const dfit_opts* opts, int n_echo, int64_t n_vox, const double* x,
                         const void* const* y_planes, int y_dtype, const uint8_t* mask, const void* p0_voxel, int p0_dtype,
                         void* popt, void* r2, int out_dtype, uint8_t* status, uint8_t* niter);

int dfit_fit_host(dfit_handle* h, const dfit_opts* opts, int n_echo, int64_t n_vox, const double* x,
                  const void* const* y_planes, int y_dtype, const uint8_t* mask, const void* p0_voxel, int p0_dtype,
                  void* popt, void* r2, int out_dtype, uint8_t* status, uint8_t* niter) {
  if (!h) return fail(DFIT_ERR_BAD_ARG, "handle is NULL");
  const int rc = fit_host_impl(h, opts, n_echo, n_vox, x, y_planes, y_dtype, mask, p0_voxel, p0_dtype, popt, r2, out_dtype,
                               status, niter);
  if (rc != DFIT_OK) {
    // An error in the middle of the chunk loop leaves earlier chunks in flight, with asynchronous copies still
    // writing into the caller's buffers: nothing may be pending on this handle when the error is returned.
    // (The error detail of the failure itself is kept.)
    char detail[sizeof(dfit::g_err)];
    std::memcpy(detail, dfit::g_err, sizeof(detail));
    if (cudaSetDevice(h->device) == cudaSuccess)
      for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(h->slots[s].stream);
    cudaGetLastError();
    std::memcpy(dfit::g_err, detail, sizeof(detail));
    h->ev_valid = false;
  }
  return rc;
}

static int fit_host_impl(dfit_handle* h, const dfit_opts* opts, int n_echo, int64_t n_vox, const double* x,
                         const void* const* y_planes, int y_dtype, const uint8_t* mask, const void* p0_voxel, int p0_dtype,
                         void* popt, void* r2, int out_dtype, uint8_t* status, uint8_t* niter) {
  int rc = validate(opts, n_echo, n_vox, x, y_dtype, p0_dtype, out_dtype, p0_voxel != nullptr);
  if (rc != DFIT_OK) return rc;
  if (n_vox > 0 && (!y_planes || !popt || !r2)) return fail(DFIT_ERR_BAD_ARG, "y_planes/popt/r2 must not be NULL");
  for (int e = 0; e < n_echo && n_vox > 0; ++e)
    if (!y_planes[e]) return fail(DFIT_ERR_BAD_ARG, "y_planes[%d] is NULL", e);
  CUDA_TRY(cudaSetDevice(h->device));
  const auto t0 = std::chrono::steady_clock::now();
  const cpu_set_t* cpus = h->have_local_cpus ? &h->local_cpus : nullptr;
  g_worker_cpus = cpus;
  const int P = model_nparams(opts->model);
  const int PW = opts->out_param >= 0 ? 1 : P;  // parameters written per voxel
  // fp32 arithmetic with float64 result maps (the reference's dtype, fitting.py:870): the maps cross PCIe as float32 --
  // half the device-to-host bytes -- and are widened by the host threads that drain the staging blocks.  Rounded
  // parameters come back onto their decimal grid in float64 there (rint(v 10^d) / 10^d, numpy.around's own formula).
  double wscale[DFIT_MAX_PARAMS];
  const bool widen = widen_plan(opts, out_dtype, wscale);
  const int dev_out = widen ? DFIT_F32 : out_dtype;
  const size_t ysz = dtype_size(y_dtype), osz = dev_out == DFIT_F32 ? 4 : 8, psz = p0_dtype == DFIT_F32 ? 4 : 8;
  const size_t hosz = out_dtype == DFIT_F32 ? 4 : 8;  // element size of the caller's maps

  // Chunks of voxels flow through kSlots stream slots: H2D(chunk i+1) overlaps fit(chunk i) and
  // D2H(chunk i-1).  Chunk size keeps every copy large enough to run PCIe at full rate.
  int64_t chunk = 1 << 22;  // measured on the 384^3 x 8 workload: 2^20 40.5 ms, 2^21 36.9, 2^22 36.3, 2^23 36.5
  if (n_vox < chunk * 2) chunk = (n_vox + 1) / 2;
  chunk = (chunk + 127) / 128 * 128;  // 512 B row alignment for vector/TMA access
  if (chunk <= 0) chunk = 128;

  CUDA_TRY(cudaMemsetAsync(h->counters, 0, kStatSlots * CNT_COUNT * sizeof(unsigned long long), h->slots[0].stream));
  CUDA_TRY(cudaStreamSynchronize(h->slots[0].stream));
  h->last_launches = 0;

  LaunchDesc d;
  make_desc(opts, n_echo, n_vox, x, d);
  if (!p0_voxel) d.p0_voxel_bits = 0;
  d.y_dtype = y_dtype;
  d.layout = DFIT_PLANAR;
  d.ld = chunk;
  d.p0_dtype = p0_dtype;
  d.out_dtype = dev_out;
  d.counters = h->counters;

  // the E planes of one contiguous (E, N) host array are equidistant: detect it once
  ptrdiff_t plane_pitch = 0;
  if (n_echo >= 2 && n_vox > 0) {
    plane_pitch = (const char*)y_planes[1] - (const char*)y_planes[0];
    for (int e = 2; e < n_echo && plane_pitch > 0; ++e)
      if ((const char*)y_planes[e] - (const char*)y_planes[e - 1] != plane_pitch) plane_pitch = 0;
    // (cudaMemcpy2D pitches are limited to cudaDeviceProp::memPitch = 2^31 - 1 bytes)
    if (plane_pitch < (ptrdiff_t)((size_t)n_vox * ysz) || plane_pitch > (ptrdiff_t)0x7fffffff) plane_pitch = 0;
  }
  // Pageable caller buffers (numpy arrays ...) go through page-locked staging, copied by host threads:
  // in:  planes -> slot.hin  (host threads)  -> device (one DMA per chunk)
  // out: device -> slot.hpopt / hr2 (DMA)    -> caller's arrays (host threads), one chunk behind the GPU
  bool in_pageable = false;
  for (int e = 0; e < n_echo && n_vox > 0; ++e) in_pageable = in_pageable || !is_pinned_or_device(y_planes[e]);
  // results go through the page-locked staging blocks when the caller's arrays are pageable or have to be widened
  const bool out_pageable = n_vox > 0 && (widen || !is_pinned_or_device(popt) || !is_pinned_or_device(r2));
  struct Pending {
    int64_t v0, n;
  } pending[kSlots];
  auto drain = [&](int64_t j) -> int {  // chunk j's results: staging -> caller's arrays
    Slot& sj = h->slots[j % kSlots];
    CUDA_TRY(cudaEventSynchronize(sj.ev_out));
    const Pending& pj = pending[j % kSlots];
    if (widen) {
      double sc[DFIT_MAX_PARAMS];
      for (int i = 0; i < PW; ++i) sc[i] = wscale[opts->out_param >= 0 ? opts->out_param : i];
      par_widen((double*)popt + (size_t)pj.v0 * PW, (const float*)sj.hpopt.p, pj.n, PW, sc);
      const double none[1] = {0.0};
      par_widen((double*)r2 + (size_t)pj.v0, (const float*)sj.hr2.p, pj.n, 1, none);
      return DFIT_OK;
    }
    const CopySeg segs[2] = {{(char*)popt + (size_t)pj.v0 * PW * hosz, sj.hpopt.p, (size_t)pj.n * PW * hosz},
                             {(char*)r2 + (size_t)pj.v0 * hosz, sj.hr2.p, (size_t)pj.n * hosz}};
    par_copy(segs, 2);
    return DFIT_OK;
  };
  int64_t idx = 0;
  for (int64_t v0 = 0; v0 < n_vox; v0 += chunk, ++idx) {
    const int64_t n = n_vox - v0 < chunk ? n_vox - v0 : chunk;
    Slot& sl = h->slots[idx % kSlots];
    if ((rc = ensure(sl.y, (size_t)n_echo * chunk * ysz)) != DFIT_OK) return rc;
    if ((rc = ensure(sl.popt, (size_t)chunk * PW * osz)) != DFIT_OK) return rc;
    if ((rc = ensure(sl.r2, (size_t)chunk * osz)) != DFIT_OK) return rc;
    if (mask && (rc = ensure(sl.mask, (size_t)chunk)) != DFIT_OK) return rc;
    if (mask && (rc = ensure(sl.index, (size_t)chunk * sizeof(unsigned) + 16)) != DFIT_OK) return rc;
    if (p0_voxel && (rc = ensure(sl.p0, (size_t)chunk * P * psz)) != DFIT_OK) return rc;
    if (status && (rc = ensure(sl.status, (size_t)chunk)) != DFIT_OK) return rc;
    if (niter && (rc = ensure(sl.niter, (size_t)chunk)) != DFIT_OK) return rc;
    cudaStream_t st = sl.stream;
    if (in_pageable) {
      if ((rc = ensure_host(sl.hin, (size_t)n_echo * chunk * ysz, cpus)) != DFIT_OK) return rc;
      if (idx >= kSlots) CUDA_TRY(cudaEventSynchronize(sl.ev_in));  // the DMA that last read this staging block is done
      CopySeg segs[DFIT_MAX_ECHOES];
      for (int e = 0; e < n_echo; ++e)
        segs[e] = CopySeg{(char*)sl.hin.p + (size_t)e * chunk * ysz, (const char*)y_planes[e] + (size_t)v0 * ysz, (size_t)n * ysz};
      par_copy(segs, n_echo);
      if (n == chunk) {
        CUDA_TRY(cudaMemcpyAsync(sl.y.p, sl.hin.p, (size_t)n_echo * chunk * ysz, cudaMemcpyHostToDevice, st));
      } else {
        CUDA_TRY(cudaMemcpy2DAsync(sl.y.p, (size_t)chunk * ysz, sl.hin.p, (size_t)chunk * ysz, (size_t)n * ysz, (size_t)n_echo,
                                   cudaMemcpyHostToDevice, st));
      }
      CUDA_TRY(cudaEventRecord(sl.ev_in, st));
    } else if (plane_pitch > 0) {  // equidistant planes (one (E, N) array): one strided copy per chunk instead of E
      CUDA_TRY(cudaMemcpy2DAsync(sl.y.p, (size_t)chunk * ysz, (const char*)y_planes[0] + (size_t)v0 * ysz, (size_t)plane_pitch,
                                 (size_t)n * ysz, (size_t)n_echo, cudaMemcpyHostToDevice, st));
    } else {
      for (int e = 0; e < n_echo; ++e)
        CUDA_TRY(cudaMemcpyAsync((char*)sl.y.p + (size_t)e * chunk * ysz, (const char*)y_planes[e] + (size_t)v0 * ysz,
                                 (size_t)n * ysz, cudaMemcpyHostToDevice, st));
    }
    if (mask) CUDA_TRY(cudaMemcpyAsync(sl.mask.p, mask + v0, (size_t)n, cudaMemcpyHostToDevice, st));
    if (p0_voxel)
      CUDA_TRY(cudaMemcpyAsync(sl.p0.p, (const char*)p0_voxel + (size_t)v0 * P * psz, (size_t)n * P * psz,
                               cudaMemcpyHostToDevice, st));
    d.n_vox = n;
    d.y = sl.y.p;
    d.mask = mask ? (const uint8_t*)sl.mask.p : nullptr;
    d.index_count = mask ? reinterpret_cast<unsigned*>(sl.index.p) : nullptr;
    d.index = mask ? d.index_count + 4 : nullptr;
    d.p0v = p0_voxel ? sl.p0.p : nullptr;
    d.popt = sl.popt.p;
    d.r2 = sl.r2.p;
    d.status = status ? (uint8_t*)sl.status.p : nullptr;
    d.niter = niter ? (uint8_t*)sl.niter.p : nullptr;
    d.stream = st;
    CUtensorMap tmap;
    d.sm_count = h->sm_count;
    d.tmap = nullptr;
    if (tma_eligible(d) && make_sample_tmap(&tmap, d.y, n_echo, n, chunk)) d.tmap = &tmap;
    CUtensorMap tmap2;
    d.tmap2 = nullptr;
    if (tma2_eligible(d) && make_sample_tmap(&tmap2, d.y, n_echo, n, chunk, kM2Tile, y_dtype)) d.tmap2 = &tmap2;
    d.lm_list = nullptr;
    if ((rc = lm_tail_prepare(sl.lm, sl.lm_parity, chunk, st, d)) != DFIT_OK) return rc;
    const int par0 = sl.lm_parity;
    CUDA_TRY(dispatch(d));
    h->last_launches += (mask && d.tmap == nullptr ? 2 : 1) + (sl.lm_parity != par0 ? 1 : 0);
    if (out_pageable) {
      if ((rc = ensure_host(sl.hpopt, (size_t)chunk * PW * osz, cpus)) != DFIT_OK) return rc;
      if ((rc = ensure_host(sl.hr2, (size_t)chunk * osz, cpus)) != DFIT_OK) return rc;
      CUDA_TRY(cudaMemcpyAsync(sl.hpopt.p, sl.popt.p, (size_t)n * PW * osz, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(sl.hr2.p, sl.r2.p, (size_t)n * osz, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaEventRecord(sl.ev_out, st));
      pending[idx % kSlots] = Pending{v0, n};
      // the staging blocks of a slot are emptied before the slot is used again: drain the oldest chunk in flight
      if (idx >= kSlots - 1 && (rc = drain(idx - (kSlots - 1))) != DFIT_OK) return rc;
    } else {
      CUDA_TRY(cudaMemcpyAsync((char*)popt + (size_t)v0 * PW * osz, sl.popt.p, (size_t)n * PW * osz, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync((char*)r2 + (size_t)v0 * osz, sl.r2.p, (size_t)n * osz, cudaMemcpyDeviceToHost, st));
    }
    if (status) CUDA_TRY(cudaMemcpyAsync(status + v0, sl.status.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (niter) CUDA_TRY(cudaMemcpyAsync(niter + v0, sl.niter.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  if (out_pageable)  // the chunks still in flight
    for (int64_t j = idx >= kSlots - 1 ? idx - (kSlots - 1) : 0; j < idx; ++j)
      if ((rc = drain(j)) != DFIT_OK) return rc;
  for (int s = 0; s < kSlots; ++s) CUDA_TRY(cudaStreamSynchronize(h->slots[s].stream));
  const auto t1 = std::chrono::steady_clock::now();
  h->last_total_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
  h->ev_valid = false;
  h->last_n = n_vox;
  return DFIT_OK;
}

int dfit_set_gather_ex(dfit_handle* h, const dfit_gather_desc* gd) {
  if (!h) return fail(DFIT_ERR_BAD_ARG, "handle is NULL");
  if (!gd || gd->world == 0) {
    h->g = GatherArgs{};
    return DFIT_OK;
  }
  if (gd->struct_size != (int32_t)sizeof(dfit_gather_desc))
    return fail(DFIT_ERR_BAD_ARG, "dfit_gather_desc.struct_size=%d, expected %d", gd->struct_size, (int)sizeof(dfit_gather_desc));
  if (gd->world < 1 || gd->world > kMaxPeers || gd->rank < 0 || gd->rank >= gd->world || !gd->maps || gd->rows < 0 || gd->row0 < 0)
    return fail(DFIT_ERR_BAD_ARG, "bad gather specification (world=%d rank=%d, at most %d peers)", gd->world, gd->rank, kMaxPeers);
  CUDA_TRY(cudaSetDevice(h->device));
  GatherArgs g{};
  for (int r = 0; r < gd->world; ++r) {
    if (!gd->maps[r]) return fail(DFIT_ERR_BAD_ARG, "maps[%d] is NULL", r);
    cudaPointerAttributes at;
    CUDA_TRY(cudaPointerGetAttributes(&at, gd->maps[r]));
    if (at.type != cudaMemoryTypeDevice) return fail(DFIT_ERR_BAD_ARG, "maps[%d] is not device memory", r);
    if (at.device != h->device) {  // a peer's map: this device must be allowed to store into it
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, h->device, at.device));
      if (!can) return fail(DFIT_ERR_UNSUPPORTED, "device %d cannot access peer device %d", h->device, at.device);
    }
    g.maps[r] = static_cast<float*>(gd->maps[r]);
  }
  g.mc = static_cast<float*>(gd->multicast);
  g.world = gd->world;
  g.self = gd->rank;
  g.cols = gd->param_mask;
  g.ncols = 0;  // (set per launch: depends on the model)
  g.row0 = gd->row0;
  g.split_list = gd->split_list ? 1 : 0;
  g.fit_lo = gd->fit_lo;
  g.fit_hi = gd->fit_hi;
  g.y_voxel0 = gd->y_voxel0;
  if (g.split_list && !(g.fit_lo >= 0 && g.fit_hi >= g.fit_lo && g.y_voxel0 >= 0 && g.y_voxel0 <= g.fit_lo))
    return fail(DFIT_ERR_BAD_ARG, "split_list needs 0 <= y_voxel0 <= fit_lo <= fit_hi");
  if (const char* e = std::getenv("DFIT_MC_WEAK")) g.mc_weak = e[0] == '1';
  h->g = g;
  h->gather_rows = gd->rows;
  return DFIT_OK;
}

int dfit_set_gather(dfit_handle* h, int world, int rank, void* const* maps, int64_t rows_per_rank) {
  if (world == 0) return dfit_set_gather_ex(h, nullptr);
  dfit_gather_desc gd;
  std::memset(&gd, 0, sizeof(gd));
  gd.struct_size = (int32_t)sizeof(gd);
  gd.world = world;
  gd.rank = rank;
  gd.maps = maps;
  gd.rows = (int64_t)world * rows_per_rank;
  gd.row0 = (int64_t)rank * rows_per_rank;
  return dfit_set_gather_ex(h, &gd);
}

int dfit_ipc_alloc(dfit_handle* h, size_t bytes, void** dev_ptr, unsigned char* handle_out) {
  if (!h || !dev_ptr || !handle_out || bytes == 0) return fail(DFIT_ERR_BAD_ARG, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == DFIT_IPC_HANDLE_BYTES, "IPC handle size");
  CUDA_TRY(cudaSetDevice(h->device));
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  CUDA_TRY(cudaMemset(p, 0, bytes));
  // cudaMemset of device memory may return before the fill has run: it must not land on top of a peer's
  // stores once the handle has been published
  CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(DFIT_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  std::memcpy(handle_out, &hd, sizeof(hd));
  *dev_ptr = p;
  return DFIT_OK;
}

int dfit_ipc_open(dfit_handle* h, const unsigned char* handle, void** dev_ptr) {
  if (!h || !handle || !dev_ptr) return fail(DFIT_ERR_BAD_ARG, "bad argument");
  // Opened with THIS rank's device current so that the mapping is made for the device whose kernels
  // will store through it (lazy peer access over NVLink).
  CUDA_TRY(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle, sizeof(hd));
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return DFIT_OK;
}

int dfit_ipc_close(dfit_handle* h, void* dev_ptr) {
  if (!h || !dev_ptr) return fail(DFIT_ERR_BAD_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return DFIT_OK;
}

int dfit_ipc_free(dfit_handle* h, void* dev_ptr) {
  if (!h || !dev_ptr) return fail(DFIT_ERR_BAD_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaFree(dev_ptr));
  return DFIT_OK;
}

int dfit_get_stats(dfit_handle* h, dfit_stats* out) {
  if (!h || !out) return fail(DFIT_ERR_BAD_ARG, "NULL argument");
  CUDA_TRY(cudaSetDevice(h->device));
  std::memset(out, 0, sizeof(*out));
  float kms = -1.f;
  if (h->ev_valid) {
    CUDA_TRY(cudaEventSynchronize(h->ev_stop));
    CUDA_TRY(cudaEventElapsedTime(&kms, h->ev_start, h->ev_stop));
  } else {
    for (int s = 0; s < kSlots; ++s) CUDA_TRY(cudaStreamSynchronize(h->slots[s].stream));
  }
  static thread_local unsigned long long slots[kStatSlots * CNT_COUNT];
  CUDA_TRY(cudaMemcpy(slots, h->counters, sizeof(slots), cudaMemcpyDeviceToHost));
  unsigned long long c[CNT_COUNT] = {0};
  for (int s = 0; s < kStatSlots; ++s)
    for (int k = 0; k < CNT_COUNT; ++k) {
      const unsigned long long v = slots[s * CNT_COUNT + k];
      if (k == CNT_MAXITER) c[k] = v > c[k] ? v : c[k];
      else c[k] += v;
    }
  out->n_voxels = h->last_n;
  out->n_fitted = (int64_t)c[CNT_FITTED];
  out->n_failed = (int64_t)c[CNT_FAILED];
  out->n_nonfinite = (int64_t)c[CNT_NONFINITE];
  out->n_oob = (int64_t)c[CNT_OOB];
  out->sum_iters = (int64_t)c[CNT_ITERS];
  out->max_iters = (int32_t)c[CNT_MAXITER];
  out->n_deferred = (int64_t)c[CNT_DEFERRED];
  out->n_launches = h->last_launches;
  out->kernel_ms = kms;
  out->total_ms = h->last_total_ms;
  return DFIT_OK;
}

}  // extern "C"
