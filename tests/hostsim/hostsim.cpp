// TEST-ONLY harness: compiles the device solver headers (dosma_b200/csrc/lm_core.cuh, mono_fast.cuh) with g++ so the
// CPU test-suite (-m "not gpu") can exercise the exact per-voxel arithmetic the CUDA kernels run.
// It is NOT part of the product: dosma_b200/ never loads it and the shipped library (libdfit.so)
// contains no host implementation of the fit (no CPU fallback).
#include <cstdint>
#include <cstring>

#include "../../dosma_b200/csrc/lm_core.cuh"
#include "../../dosma_b200/csrc/mono_fast.cuh"

using namespace dfit;

// fit_kernel_lmq's way through the LM: lm_begin, then rounds of lm_iterate with a budget of evaluations, the solver
// state parked (copied away and back) between rounds.  0 = off: fit_voxel / lm_solve.
static int g_round_first = 0, g_round_next = 0, g_uni = 0;
extern "C" void hostsim_set_uniform_recurrence(int on) { g_uni = on; }
extern "C" void hostsim_set_rounds(int k_first, int k_next) {
  g_round_first = k_first;
  g_round_next = k_next;
}

template <class M, typename T, typename TA, int EMAX, bool EXACT>
static int fit_voxel_any(const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, const VoxelOpts<T>& vo, T (&p)[M::P], T& r2,
                         int& iters, unsigned& flags) {
  if (g_round_first <= 0) return fit_voxel<M, T, TA, EMAX, EXACT>(y, xt, E, vo, p, r2, iters, flags);
  constexpr int P = M::P;
  LmState<P, T, TA> s;
  s.F = 0;
  s.iters = 0;
  int st = voxel_prepare<M, T, EMAX, EXACT>(y, xt, E, vo, p, flags);
  // (fit_kernel_lmq's choice: exponentials by recurrence on uniformly spaced echoes, fp32, >= 4 echoes)
  constexpr bool CAN_UNI = M::HAS_REC && EXACT && EMAX >= 4 && sizeof(T) == 4 && sizeof(TA) == 4;
  const bool uni = CAN_UNI && g_uni && xt.uniform != 0;
  if (st == ST_PENDING)
    st = uni ? lm_begin<M, T, TA, EMAX, EXACT, CAN_UNI>(p, y, xt, E, vo.s, s) : lm_begin<M, T, TA, EMAX, EXACT>(p, y, xt, E, vo.s, s);
  bool resume = false;
  int budget = g_round_first;
  while (st == ST_PENDING) {
    st = uni ? lm_iterate<M, T, TA, EMAX, EXACT, CAN_UNI>(p, y, xt, E, vo.s, s, budget, resume)
             : lm_iterate<M, T, TA, EMAX, EXACT>(p, y, xt, E, vo.s, s, budget, resume);
    if (st == ST_PENDING) {  // park: the state and the parameters survive as plain bytes, nothing else does
      unsigned char park[sizeof(s) + sizeof(p)];
      memcpy(park, &s, sizeof(s));
      memcpy(park + sizeof(s), p, sizeof(p));
      memset(&s, 0xff, sizeof(s));
      memset(p, 0xff, sizeof(p));
      memcpy(&s, park, sizeof(s));
      memcpy(p, park + sizeof(s), sizeof(p));
      resume = true;
      budget = g_round_next;
    }
  }
  voxel_finish<M, T, EMAX, EXACT>(st, y, E, vo, (T)s.F, p, r2);
  iters = s.iters;
  return st;
}

template <class M, typename T, typename TA, int EMAX, bool EXACT>
static void run(int E, int64_t N, const double* x, const double* y, const double* p0, int64_t n_p0, int init_mode,
                int init_linear, int fast, double ftol, double xtol, double lambda0, double floor_rel, int max_iter,
                double r2_eps, double y_lo, double y_hi, double* popt, double* r2, int32_t* status, int32_t* iters) {
  constexpr int P = M::P;
  XTab<T, EMAX> xt;
  fill_xtab<T, EMAX>(xt, x, E);
  VoxelOpts<T> vo;
  vo.s.ftol = (T)ftol;
  vo.s.xtol = (T)xtol;
  vo.s.lambda0 = (T)lambda0;
  vo.s.floor_rel = (T)floor_rel;
  vo.s.maxfev = max_iter;
  vo.s.init_linear = init_linear;
  vo.y_lo = (T)y_lo;
  vo.y_hi = (T)y_hi;
  vo.r2_eps = (T)r2_eps;
  vo.init_mode = init_mode;
  vo.has_bounds = (y_lo > -1.7e308 || y_hi < 1.7e308) ? 1 : 0;
  vo.fast = fast;
  // fast == 2: the two-voxels-per-lane variant of the fast path on consecutive voxel pairs (what
  // fit_kernel_mono2 runs), general path for the voxels it declines
  if constexpr (M::MONO && EXACT && EMAX >= 3 && sizeof(T) == sizeof(TA)) {
    if (fast == 2 && !vo.has_bounds) {
      for (int64_t v = 0; v < N; v += 2) {
        const bool both = v + 1 < N;
        pair2<T> Y[EMAX], pa, pb, r2p;
        for (int e = 0; e < EMAX; ++e)
          Y[e] = p2_make<T>((T)y[(size_t)e * N + v], (T)y[(size_t)e * N + (both ? v + 1 : v)]);
        // what the dense GPU kernels do: the straight-line two-pass attempt for the pair; a voxel it turns
        // down goes through the one-voxel path (generic Newton loop, then the LM from p0)
        bool ok2[2];
        fit_voxel_fast2s<M, T, EMAX, pair2<T>[EMAX]>(Y, xt, vo, pa, pb, r2p, ok2);
        for (int hsel = 0; hsel < (both ? 2 : 1); ++hsel) {
          T p[P], r2v = hsel ? r2p.hi : r2p.lo;
          int st = ok2[hsel] ? (int)ST_CONV_F : -1, it = kFast2Passes;
          p[0] = hsel ? pa.hi : pa.lo;
          p[P - 1] = hsel ? pb.hi : pb.lo;
          if (st < 0) {
            T yy[EMAX];
            unsigned flags;
            for (int e = 0; e < EMAX; ++e) yy[e] = hsel ? Y[e].hi : Y[e].lo;
            st = fit_voxel_fast<M, T, EMAX, EXACT>(yy, xt, vo, p, r2v, it);
            if (st < 0) {
              const double* pv = p0 + (n_p0 > 1 ? (size_t)(v + hsel) * P : 0);
              for (int i = 0; i < P; ++i) p[i] = (T)pv[i];
              st = fit_voxel_any<M, T, TA, EMAX, EXACT>(yy, xt, E, vo, p, r2v, it, flags);
            }
          }
          for (int i = 0; i < P; ++i) popt[(size_t)(v + hsel) * P + i] = (double)p[i];
          r2[v + hsel] = (double)r2v;
          status[v + hsel] = st;
          iters[v + hsel] = it;
        }
      }
      return;
    }
  }
  for (int64_t v = 0; v < N; ++v) {
    T yy[EMAX];
    for (int e = 0; e < EMAX; ++e) yy[e] = e < E ? (T)y[(size_t)e * N + v] : (T)0;
    T p[P];
    const double* pv = p0 + (n_p0 > 1 ? (size_t)v * P : 0);
    for (int i = 0; i < P; ++i) p[i] = (T)pv[i];
    T r2v;
    int it;
    unsigned flags;
    int st = -1;
    flags = 0;
    if constexpr (sizeof(T) == sizeof(TA)) st = fit_voxel_fast<M, T, EMAX, EXACT>(yy, xt, vo, p, r2v, it);
    if (st < 0) {
      for (int i = 0; i < P; ++i) p[i] = (T)pv[i];
      st = fit_voxel_any<M, T, TA, EMAX, EXACT>(yy, xt, E, vo, p, r2v, it, flags);
    }
    for (int i = 0; i < P; ++i) popt[(size_t)v * P + i] = (double)p[i];
    r2[v] = (double)r2v;
    status[v] = st;
    iters[v] = it;
  }
}

extern "C" int hostsim_fit(int model, int dtype, int acc64, int E, int64_t N, const double* x, const double* y,
                           const double* p0, int64_t n_p0, int init_mode, int init_linear, int fast, double ftol, double xtol,
                           double lambda0, double floor_rel, int max_iter, double r2_eps, double y_lo, double y_hi,
                           double* popt, double* r2, int32_t* status, int32_t* iters) {
  if (E > 32) return -1;
#define ARGS E, N, x, y, p0, n_p0, init_mode, init_linear, fast, ftol, xtol, lambda0, floor_rel, max_iter, r2_eps, y_lo, y_hi, popt, r2, status, iters
#define RUN_E(M, T, TA)                                                                     \
  switch (E) {                                                                                \
    case 3: if (E == 3 && 3 >= M::P) { run<M, T, TA, 3, true>(ARGS); break; }                 \
    case 4: if (E == 4 && 4 >= M::P) { run<M, T, TA, 4, true>(ARGS); break; }                 \
    case 5: if (E == 5 && 5 >= M::P) { run<M, T, TA, 5, true>(ARGS); break; }                 \
    case 7: if (E == 7 && 7 >= M::P) { run<M, T, TA, 7, true>(ARGS); break; }                 \
    case 8: if (E == 8) { run<M, T, TA, 8, true>(ARGS); break; }                              \
    case 16: if (E == 16) { run<M, T, TA, 16, true>(ARGS); break; }                           \
    default: run<M, T, TA, 32, false>(ARGS);                                                  \
  }
#define DISPATCH(M)                                  \
  if (dtype == 0 && !acc64) { RUN_E(M, float, float) } \
  else if (dtype == 0) { run<M, float, double, 32, false>(ARGS); } \
  else { RUN_E(M, double, double) }
  switch (model) {
    case 0: DISPATCH(MonoExp); break;
    case 1: DISPATCH(BiExp); break;
    case 2: DISPATCH(Linear1); break;
    default: return -2;
  }
  return 0;
}

extern "C" double hostsim_post_param(int enabled, const int* ufunc, const double* lb, const double* ub, int has_thr,
                                     double thr, int has_fill, double fill, const int* decimals, int i, double v,
                                     double r2) {
  PostOpts po;
  po.enabled = enabled;
  for (int k = 0; k < 4; ++k) {
    po.ufunc[k] = ufunc[k];
    po.lb[k] = lb[k];
    po.ub[k] = ub[k];
    po.decimals[k] = decimals[k];
  }
  set_post_scales(po);
  po.has_r2_thresh = has_thr;
  po.r2_thresh = thr;
  po.has_fill = has_fill;
  po.fill = fill;
  return post_param(po, i, v, r2);
}

static int g_pairs = 0;

// fp32 epilogue (post_param_f32) next to the float64 one (post_param) on the same inputs: returns both
extern "C" void hostsim_post_param_f32(int enabled, const int* ufunc, const double* lb, const double* ub, int has_thr,
                                       double thr, int has_fill, double fill, const int* decimals, int i, int64_t n,
                                       const float* v, const float* r2, float* out_f32, float* out_f64path) {
  PostOpts po;
  po.enabled = enabled;
  for (int k = 0; k < 4; ++k) {
    po.ufunc[k] = ufunc[k];
    po.lb[k] = lb[k];
    po.ub[k] = ub[k];
    po.decimals[k] = decimals[k];
  }
  po.has_r2_thresh = has_thr;
  po.r2_thresh = thr;
  po.has_fill = has_fill;
  po.fill = fill;
  set_post_scales(po);
  // even n: through the packed two-voxel form (what the two-voxel kernels run), pairs (k, k + 1); the last of an odd n
  // and everything when `pairs` is off: the one-voxel form
  for (int64_t k = 0; k < n; ++k) out_f64path[k] = (float)post_param(po, i, (double)v[k], (double)r2[k]);
  int64_t k = 0;
  if (g_pairs)
    for (; k + 1 < n; k += 2) {
      const pair2<float> o = post_pair_f32(po, i, p2_make<float>(v[k], v[k + 1]), p2_make<float>(r2[k], r2[k + 1]));
      out_f32[k] = o.lo;
      out_f32[k + 1] = o.hi;
    }
  for (; k < n; ++k) out_f32[k] = post_param_f32(po, i, v[k], r2[k]);
}

extern "C" void hostsim_set_pairs(int on) { g_pairs = on; }

// echo-table classification the launcher relies on: bit 0 uniform spacing, bit 1 descending (backward Prony)
extern "C" int hostsim_xtab_flags(int dtype, int E, const double* x) {
  if (E < 1 || E > 16) return -1;
  if (dtype == 0) {
    XTab<float, 16> xt;
    fill_xtab<float, 16>(xt, x, E);
    return xt.uniform | (xt.backward << 1);
  }
  XTab<double, 16> xt;
  fill_xtab<double, 16>(xt, x, E);
  return xt.uniform | (xt.backward << 1);
}

// which plan set_post_scales chose for parameter i: bit 0 simple (comparisons only), bit 1 fastinv (fp32 1 / |v| form)
extern "C" int hostsim_post_plan(const int* ufunc, const double* lb, const double* ub, int has_fill, double fill,
                                 const int* decimals, int i) {
  PostOpts po;
  po.enabled = 1;
  for (int k = 0; k < 4; ++k) {
    po.ufunc[k] = ufunc[k];
    po.lb[k] = lb[k];
    po.ub[k] = ub[k];
    po.decimals[k] = decimals[k];
  }
  po.has_r2_thresh = 0;
  po.r2_thresh = 0;
  po.has_fill = has_fill;
  po.fill = fill;
  set_post_scales(po);
  return po.simple[i] | (po.fastinv[i] << 1);
}

// The fused epilogue over arrays (what store_voxel applies per voxel), for the test-only host engine:
// popt (n, P) in place, double arithmetic like the float64 maps of the device path.
extern "C" void hostsim_post_params(int64_t n, int P, double* popt, const double* r2, const int* ufunc, const double* lb,
                                    const double* ub, int has_thr, double thr, int has_fill, double fill,
                                    const int* decimals) {
  PostOpts po;
  po.enabled = 1;
  for (int k = 0; k < 4; ++k) {
    po.ufunc[k] = ufunc[k];
    po.lb[k] = lb[k];
    po.ub[k] = ub[k];
    po.decimals[k] = decimals[k];
  }
  po.has_r2_thresh = has_thr;
  po.r2_thresh = thr;
  po.has_fill = has_fill;
  po.fill = fill;
  set_post_scales(po);
  for (int64_t v = 0; v < n; ++v)
    for (int i = 0; i < P; ++i) popt[v * P + i] = post_param(po, i, popt[v * P + i], r2[v]);
}
