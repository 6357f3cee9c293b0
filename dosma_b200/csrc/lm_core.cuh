// Per-voxel Levenberg-Marquardt core: analytic Jacobians, J^T J / J^T r normal equations kept in
// registers, Marquardt-scaled Cholesky solve, gain-ratio damping update.
//
// Replaces the arithmetic the reference delegates to SciPy/MINPACK per voxel
// (dosma/core/fitting.py:1026-1073 -> scipy.optimize.curve_fit -> lmdif) with a solver designed
// for one-voxel-per-lane SIMT execution: no callbacks, no QR workspace, no forward differences.
// It converges to the same least-squares minimiser; see DESIGN.md "Parity definition".
//
// The header is written against a tiny portability shim (DFIT_HD, dfit::num<T>) so that the very
// same solver source can be compiled by g++ into the test-only `tests/hostsim` harness, which lets
// the CPU test-suite exercise the device arithmetic.  The product library contains device code only.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DFIT_HD __host__ __device__ __forceinline__
#else
#define DFIT_HD inline __attribute__((always_inline))
#endif

namespace dfit {

// Status codes written per voxel.  1..4 mirror MINPACK's successful `info` values so that the
// host layer can treat them exactly like SciPy's `ier in (1, 2, 3, 4)` (fitting.py:1069-1073).
enum Status : int {
  ST_SKIPPED = 0,    // masked out, all samples zero, or a sample outside y_bounds (fitting.py:1065-1067)
  ST_CONV_F = 1,     // relative cost reduction (actual and predicted) below ftol
  ST_CONV_X = 2,     // scaled step below xtol * scaled parameter norm
  ST_CONV_FX = 3,    // both
  ST_EXACT = 4,      // residual at the arithmetic floor / gradient orthogonal (exact fit)
  ST_MAXITER = 5,    // iteration budget exhausted -> NaN parameters, r2 = 0 (like MINPACK info 5)
  ST_NONFINITE = 6,  // NaN/Inf in the input samples or initial guess
  ST_NUMERIC = 7,    // model not finite at the initial guess
};

template <typename T>
struct num;

template <>
struct num<float> {
  typedef float type;
  static DFIT_HD float eps() { return 1.1920929e-7f; }
  static DFIT_HD float tiny() { return 1.0e-30f; }
  static DFIT_HD float huge() { return 3.0e38f; }
  // exp(b * x); xs = x * log2(e) is precomputed on the host so the device issues FMUL + MUFU.EX2.
  static DFIT_HD float expbx(float b, float /*x*/, float xs) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b * xs));
    return r;
#else
    return exp2f(b * xs);
#endif
  }
  static DFIT_HD float exp_(float v) { return expf(v); }
  static DFIT_HD float log_(float v) { return logf(v); }
  static DFIT_HD float sqrt_(float v) { return sqrtf(v); }
  static DFIT_HD float abs_(float v) { return fabsf(v); }
  static DFIT_HD float max_(float a, float b) { return fmaxf(a, b); }
  static DFIT_HD float min_(float a, float b) { return fminf(a, b); }
  static DFIT_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
  static DFIT_HD bool finite(float v) { return fabsf(v) <= 3.4028234e38f; }
};

template <>
struct num<double> {
  typedef double type;
  static DFIT_HD double eps() { return 2.220446049250313e-16; }
  static DFIT_HD double tiny() { return 1.0e-290; }
  static DFIT_HD double huge() { return 1.0e300; }
  static DFIT_HD double expbx(double b, double x, double /*xs*/) { return exp(b * x); }
  static DFIT_HD double exp_(double v) { return exp(v); }
  static DFIT_HD double log_(double v) { return log(v); }
  static DFIT_HD double sqrt_(double v) { return sqrt(v); }
  static DFIT_HD double abs_(double v) { return fabs(v); }
  static DFIT_HD double max_(double a, double b) { return fmax(a, b); }
  static DFIT_HD double min_(double a, double b) { return fmin(a, b); }
  static DFIT_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
  static DFIT_HD bool finite(double v) { return fabs(v) <= 1.7976931348623157e308; }
};

// ------------------------------------------------------------------------------------ models
// Each model provides the value f and the analytic Jacobian row J[P] for one sample, and LIN, the
// bit-mask of parameters the model is linear in (used for the variable-projection start).

// fitting.py:1016-1018  f = a * exp(b x)
struct MonoExp {
  static constexpr int P = 2;
  static constexpr unsigned LIN = 0x1u;
  template <typename T>
  static DFIT_HD void eval(const T (&p)[2], T x, T xs, T& f, T (&J)[2]) {
    T e = num<T>::expbx(p[1], x, xs);
    J[0] = e;
    f = p[0] * e;
    J[1] = x * f;
  }
};

// fitting.py:1021-1023  f = a1 exp(b1 x) + a2 exp(b2 x)
struct BiExp {
  static constexpr int P = 4;
  static constexpr unsigned LIN = 0x5u;
  template <typename T>
  static DFIT_HD void eval(const T (&p)[4], T x, T xs, T& f, T (&J)[4]) {
    T e1 = num<T>::expbx(p[1], x, xs);
    T e2 = num<T>::expbx(p[3], x, xs);
    T f1 = p[0] * e1, f2 = p[2] * e2;
    J[0] = e1;
    J[1] = x * f1;
    J[2] = e2;
    J[3] = x * f2;
    f = f1 + f2;
  }
};

// f = a x : the 1-parameter custom model of the reference's tests (tests/core/test_fitting.py:52-53)
struct Linear1 {
  static constexpr int P = 1;
  static constexpr unsigned LIN = 0x1u;
  template <typename T>
  static DFIT_HD void eval(const T (&p)[1], T x, T /*xs*/, T& f, T (&J)[1]) {
    J[0] = x;
    f = p[0] * x;
  }
};

// ------------------------------------------------------------------------------------ options
template <typename T>
struct SolverOpts {
  T ftol;       // relative cost-reduction tolerance (engine tolerance, tighter than SciPy's 1e-5)
  T xtol;       // relative scaled-step tolerance
  T lambda0;    // initial Marquardt damping (relative to the unit-diagonal scaled normal matrix)
  T floor_rel;  // F <= floor_rel * sum(y^2) counts as an exact fit
  int maxfev;   // evaluation budget counted like MINPACK: 1 per trial step + P per accepted step
                // (an accepted step is where lmdif re-differences its Jacobian); fitting.py:761
  int init_linear;  // 1: start from the linear-least-squares optimum of the linear parameters
};

// Packed lower-triangular index, i >= j.
DFIT_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// One pass over the echoes: residuals r = f - y, cost F = sum r^2, A = J^T J (packed), g = J^T r.
// TA is the accumulator type (float, or double for the ill-conditioned 4-parameter model).
template <class M, typename T, typename TA, int EMAX, bool EXACT>
DFIT_HD void eval_all(const T (&p)[M::P], const T (&y)[EMAX], const T* __restrict__ x,
                      const T* __restrict__ xs, int E, T (&r)[EMAX], TA& F, TA (&A)[M::P * (M::P + 1) / 2],
                      TA (&g)[M::P]) {
  constexpr int P = M::P;
  F = 0;
#pragma unroll
  for (int k = 0; k < P * (P + 1) / 2; ++k) A[k] = 0;
#pragma unroll
  for (int k = 0; k < P; ++k) g[k] = 0;
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (EXACT || e < E) {
      T f, J[P];
      M::template eval<T>(p, x[e], xs[e], f, J);
      T re = f - y[e];
      r[e] = re;
      F = num<TA>::fma_((TA)re, (TA)re, F);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        g[i] = num<TA>::fma_((TA)J[i], (TA)re, g[i]);
#pragma unroll
        for (int j = 0; j <= i; ++j) A[tri(i, j)] = num<TA>::fma_((TA)J[i], (TA)J[j], A[tri(i, j)]);
      }
    }
  }
}

// Solve (C + lam I) z = -gs for symmetric C (packed lower) by Cholesky, all in registers.
// Rows/columns whose bit is cleared in `active` are frozen (z_i = 0).
template <int P, typename TA>
DFIT_HD bool chol_solve(const TA (&C)[P * (P + 1) / 2], TA lam, const TA (&gs)[P], unsigned active, TA (&z)[P]) {
  TA L[P * (P + 1) / 2];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const bool aj = (active >> j) & 1u;
    TA s = aj ? C[tri(j, j)] + lam : (TA)1;
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[tri(j, k)] * L[tri(j, k)];
    if (!(s > num<TA>::eps() * (TA)4)) return false;
    TA d = num<TA>::sqrt_(s);
    L[tri(j, j)] = d;
    TA inv = (TA)1 / d;
#pragma unroll
    for (int i = j + 1; i < P; ++i) {
      const bool ai = (active >> i) & 1u;
      TA t = (aj && ai) ? C[tri(i, j)] : (TA)0;
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[tri(i, k)] * L[tri(j, k)];
      L[tri(i, j)] = t * inv;
    }
  }
  TA w[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    TA t = ((active >> i) & 1u) ? -gs[i] : (TA)0;
#pragma unroll
    for (int k = 0; k < i; ++k) t -= L[tri(i, k)] * w[k];
    w[i] = t / L[tri(i, i)];
  }
#pragma unroll
  for (int i = P - 1; i >= 0; --i) {
    TA t = w[i];
#pragma unroll
    for (int k = i + 1; k < P; ++k) t -= L[tri(k, i)] * z[k];
    z[i] = t / L[tri(i, i)];
  }
  return true;
}

// Log-linear (degree-1 polyfit of ln y) initial guess for the mono-exponential model:
// fitting.py:701-718.  Exact zeros are replaced by 1e-10 before the log (:713); a negative sample
// makes the log NaN, the reference's nan_to_num=0 then yields (slope, intercept) = (0, 0), i.e.
// p0 = (a = exp(0) = 1, b = 0) (:703-706, :717).  xc = x - mean(x), inv_sxx = 1 / sum(xc^2).
template <typename T, int EMAX, bool EXACT>
DFIT_HD void loglinear_init(const T (&y)[EMAX], const T* __restrict__ xc, T xbar, T inv_sxx, int E, T (&p)[2]) {
  T sl = 0, sxl = 0;
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (EXACT || e < E) {
      T v = y[e] == (T)0 ? (T)1e-10 : y[e];
      T l = num<T>::log_(v);  // NaN for v < 0
      sl += l;
      sxl = num<T>::fma_(xc[e], l, sxl);
    }
  }
  T slope = sxl * inv_sxx;
  T icpt = sl / (T)E - slope * xbar;
  T a = num<T>::exp_(icpt);
  if (num<T>::finite(slope) && num<T>::finite(a)) {
    p[0] = a;
    p[1] = slope;
  } else {
    p[0] = (T)1;
    p[1] = (T)0;
  }
}

// The solver.  On entry p holds the initial guess; on exit the accepted parameters.
// Returns a Status; F_out is the final sum of squared residuals, iters the trial steps taken.
template <class M, typename T, typename TA, int EMAX, bool EXACT>
DFIT_HD int lm_solve(T (&p)[M::P], const T (&y)[EMAX], const T* __restrict__ x, const T* __restrict__ xs, int E,
                     const SolverOpts<T>& o, T& F_out, int& iters) {
  constexpr int P = M::P;
  constexpr int NA = P * (P + 1) / 2;
  constexpr unsigned ALL = (1u << P) - 1u;
  T r[EMAX];
  TA F, A[NA], g[P];
  iters = 0;

  TA ysq = 0;
#pragma unroll
  for (int e = 0; e < EMAX; ++e)
    if (EXACT || e < E) ysq = num<TA>::fma_((TA)y[e], (TA)y[e], ysq);
  const TA floorF = (TA)o.floor_rel * ysq;

  if (o.init_linear) {
    // Variable-projection start: with the non-linear parameters fixed, the optimum of the linear
    // ones is a linear least-squares problem -- one Gauss-Newton step from zero, restricted to them.
    T q[P];
#pragma unroll
    for (int i = 0; i < P; ++i) q[i] = ((M::LIN >> i) & 1u) ? (T)0 : p[i];
    eval_all<M, T, TA, EMAX, EXACT>(q, y, x, xs, E, r, F, A, g);
    TA D[P], C[NA], gs[P], z[P];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      D[i] = num<TA>::sqrt_(A[tri(i, i)]);
      if (!(D[i] > 0)) D[i] = 1;
      D[i] = (TA)1 / D[i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      gs[i] = g[i] * D[i];
#pragma unroll
      for (int j = 0; j <= i; ++j) C[tri(i, j)] = A[tri(i, j)] * D[i] * D[j];
    }
    ok = chol_solve<P, TA>(C, (TA)0, gs, M::LIN, z);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      q[i] = ((M::LIN >> i) & 1u) ? (T)(z[i] * D[i]) : p[i];
      ok = ok && num<T>::finite(q[i]);
    }
    if (ok) {
#pragma unroll
      for (int i = 0; i < P; ++i) p[i] = q[i];
    }
  }

  eval_all<M, T, TA, EMAX, EXACT>(p, y, x, xs, E, r, F, A, g);
  if (!num<TA>::finite(F)) {
    F_out = (T)F;
    return ST_NUMERIC;
  }

  TA D[P];
#pragma unroll
  for (int i = 0; i < P; ++i) D[i] = 0;
  TA lam = (TA)o.lambda0, nu = 2;
  int status = ST_MAXITER;
  const TA ftol = (TA)o.ftol, xtol2 = (TA)o.xtol * (TA)o.xtol;

  int fev = 1 + P;
  while (fev < o.maxfev) {
    if (F <= floorF) {
      status = ST_EXACT;
      break;
    }
    ++iters;
    ++fev;
    // Marquardt scaling with MINPACK's running-maximum rule: D_i = max(D_i, ||J_i||), 1 if zero.
    TA Di[P], C[NA], gs[P], z[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      D[i] = num<TA>::max_(D[i], num<TA>::sqrt_(A[tri(i, i)]));
      Di[i] = (TA)1 / (D[i] > 0 ? D[i] : (TA)1);
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      gs[i] = g[i] * Di[i];
#pragma unroll
      for (int j = 0; j <= i; ++j) C[tri(i, j)] = A[tri(i, j)] * Di[i] * Di[j];
    }
    if (!chol_solve<P, TA>(C, lam, gs, ALL, z)) {
      lam = num<TA>::max_(lam * (TA)10, (TA)1e-3);
      continue;
    }
    T pn[P];
#pragma unroll
    for (int i = 0; i < P; ++i) pn[i] = p[i] + (T)(z[i] * Di[i]);

    T rn[EMAX];
    TA Fn, An[NA], gn[P];
    eval_all<M, T, TA, EMAX, EXACT>(pn, y, x, xs, E, rn, Fn, An, gn);

    // predicted reduction of the linearised model: z^T C z + 2 lam z^T z  (all terms >= 0)
    TA zz = 0, zCz = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      zz = num<TA>::fma_(z[i], z[i], zz);
      zCz = num<TA>::fma_(C[tri(i, i)] * z[i], z[i], zCz);
#pragma unroll
      for (int j = 0; j < i; ++j) zCz = num<TA>::fma_((TA)2 * C[tri(i, j)] * z[i], z[j], zCz);
    }
    const TA pred = zCz + (TA)2 * lam * zz;
    // actual reduction F - Fn, accumulated as sum (r - rn)(r + rn) so that it stays accurate when
    // the two costs agree to more digits than T resolves.
    TA act = 0;
#pragma unroll
    for (int e = 0; e < EMAX; ++e)
      if (EXACT || e < E) act = num<TA>::fma_((TA)(r[e] - rn[e]), (TA)(r[e] + rn[e]), act);

    const bool good = num<TA>::finite(Fn);
    // `act` is a difference of model evaluations each carrying ~eps*|y| of rounding, so it cannot
    // resolve reductions below tau ~ eps*sqrt(sum y^2 * F).  Below that level the gain ratio is
    // noise: trust the (accurately computed) predicted reduction instead of rejecting at random.
    const TA tau = (TA)8 * (TA)num<T>::eps() * num<TA>::sqrt_(ysq * F);
    const bool reliable = pred > tau;
    const TA rho = (good && pred > 0) ? (reliable ? act / pred : (TA)1) : (TA)-1;
    const bool accept = good && (reliable ? rho > (TA)1e-4 : act > -tau);
    const TA relact = reliable ? num<TA>::abs_(act) / F : (TA)0, relpred = pred / F;
    if (accept) {
#pragma unroll
      for (int i = 0; i < P; ++i) {
        p[i] = pn[i];
        g[i] = gn[i];
      }
#pragma unroll
      for (int k = 0; k < NA; ++k) A[k] = An[k];
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (EXACT || e < E) r[e] = rn[e];
      F = Fn;
      const TA t = (TA)2 * rho - (TA)1;
      lam *= num<TA>::max_((TA)(1.0 / 3.0), (TA)1 - t * t * t);
      lam = num<TA>::max_(lam, (TA)1e-9);
      nu = 2;
      fev += P;
    } else {
      lam *= nu;
      nu *= 2;
    }
    TA pnorm2 = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) pnorm2 = num<TA>::fma_(D[i] * (TA)p[i], D[i] * (TA)p[i], pnorm2);
    const bool conv_f = good && relact <= ftol && relpred <= ftol && rho <= (TA)2;
    const bool conv_x = accept && zz <= xtol2 * pnorm2;
    if (conv_f || conv_x) {
      status = conv_f ? (conv_x ? ST_CONV_FX : ST_CONV_F) : ST_CONV_X;
      break;
    }
  }
  F_out = (T)F;
  return status;
}

// ------------------------------------------------------------------------------------ one voxel
// Echo-time table shared by all voxels of a launch (lives in the kernel parameter / constant bank).
template <typename T, int EMAX>
struct XTab {
  T x[EMAX];   // echo / spin-lock times
  T xs[EMAX];  // x * log2(e), so exp(b x) = ex2(b * xs)
  T xc[EMAX];  // x - mean(x), for the log-linear initial guess
  T xbar, inv_sxx;
};

enum InitMode : int { INIT_GIVEN = 0, INIT_LOGLINEAR = 1 };
enum VoxelFlags : unsigned { FLAG_NONFINITE = 1u, FLAG_OOB = 2u };

template <typename T>
struct VoxelOpts {
  SolverOpts<T> s;
  T y_lo, y_hi;   // y_bounds (fitting.py:1065): any sample outside -> voxel skipped
  T r2_eps;       // fitting.py:763
  int init_mode;  // InitMode
};

// Everything the reference does for one voxel (`_curve_fit`, fitting.py:1026-1073), on samples that
// are already in registers.  p: in = initial guess, out = fitted parameters (NaN on skip/failure).
template <class M, typename T, typename TA, int EMAX, bool EXACT>
DFIT_HD int fit_voxel(const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, const VoxelOpts<T>& vo, T (&p)[M::P],
                      T& r2, int& iters, unsigned& flags) {
  constexpr int P = M::P;
  bool all_zero = true, oob = false, nonfinite = false;
  T ysum = 0;
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (EXACT || e < E) {
      const T v = y[e];
      all_zero = all_zero && (v == (T)0);
      oob = oob || (v < vo.y_lo) || (v > vo.y_hi);
      nonfinite = nonfinite || !num<T>::finite(v);
      ysum += v;
    }
  }
  iters = 0;
  flags = 0;
  int status;
  T F = 0;
  if (oob || all_zero) {
    status = ST_SKIPPED;
    if (oob) flags |= FLAG_OOB;
  } else {
    if (vo.init_mode == INIT_LOGLINEAR && P == 2) {
      T q[2];
      loglinear_init<T, EMAX, EXACT>(y, xt.xc, xt.xbar, xt.inv_sxx, E, q);
      p[0] = q[0];
      p[P - 1] = q[1];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) nonfinite = nonfinite || !num<T>::finite(p[i]);
    if (nonfinite) {
      status = ST_NONFINITE;
      flags |= FLAG_NONFINITE;
    } else {
      status = lm_solve<M, T, TA, EMAX, EXACT>(p, y, xt.x, xt.xs, E, vo.s, F, iters);
    }
  }
  if (status >= ST_CONV_F && status <= ST_EXACT) {
    const T mean = ysum / (T)E;
    T ss_tot = 0;
#pragma unroll
    for (int e = 0; e < EMAX; ++e)
      if (EXACT || e < E) ss_tot = num<T>::fma_(y[e] - mean, y[e] - mean, ss_tot);
    r2 = (T)1 - F / (ss_tot + vo.r2_eps);  // fitting.py:1032-1035
  } else {
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = (T)NAN;
    r2 = (T)0;  // fitting.py:1067, 1072
  }
  return status;
}

// ------------------------------------------------------------------------------------ epilogue
// Fused restatement of `_process_params` + mask fill + rounding (fitting.py:109-146, 205-215,
// 734-737), applied per voxel in double so that rounding matches numpy's float64 `around`.
enum Ufunc : int { UF_NONE = 0, UF_INV_ABS = 1, UF_NEG_INV = 2, UF_ABS = 3, UF_INV = 4 };

struct PostOpts {
  int enabled;        // 0: raw parameters are written
  int ufunc[4];       // per parameter (fitting.py:123-128)
  double lb[4], ub[4];  // inclusive bounds -> NaN outside (fitting.py:130-138)
  int has_r2_thresh;  // parameters of voxels with r2 < thresh -> NaN (fitting.py:140-141)
  double r2_thresh;
  int has_fill;       // nan_to_num (fitting.py:143-144): NaN -> fill, +-inf -> +-DBL_MAX
  double fill;
  int decimals[4];    // np.around per parameter, < 0: none (fitting.py:736-737)
};

DFIT_HD double apply_ufunc(int id, double v) {
  switch (id) {
    case UF_INV_ABS: return 1.0 / fabs(v);
    case UF_NEG_INV: return -1.0 / v;
    case UF_ABS: return fabs(v);
    case UF_INV: return 1.0 / v;
    default: return v;
  }
}

DFIT_HD double pow10i(int d) {
  double s = 1.0;
  for (int k = 0; k < d; ++k) s *= 10.0;
  return s;
}

DFIT_HD double post_param(const PostOpts& po, int i, double v, double r2) {
  if (!po.enabled) return v;
  v = apply_ufunc(po.ufunc[i], v);
  if (v < po.lb[i] || v > po.ub[i]) v = NAN;
  if (po.has_r2_thresh && r2 < po.r2_thresh) v = NAN;
  if (po.has_fill) {
    if (v != v) v = po.fill;
    else if (v > 1.7976931348623157e308) v = 1.7976931348623157e308;
    else if (v < -1.7976931348623157e308) v = -1.7976931348623157e308;
  }
  if (po.decimals[i] >= 0) {
    const double s = pow10i(po.decimals[i]);
    v = rint(v * s) / s;  // numpy.around: round-half-even of the scaled value
  }
  return v;
}

}  // namespace dfit
