// Per-voxel Levenberg-Marquardt core: analytic Jacobians, J^T J / J^T r normal equations kept in
// registers, Marquardt-scaled Cholesky solve, gain-ratio damping update.
//
// Replaces the arithmetic the reference delegates to SciPy/MINPACK per voxel
// (dosma/core/fitting.py:1026-1073 -> scipy.optimize.curve_fit -> lmdif) with a solver designed
// for one-voxel-per-lane SIMT execution: no callbacks, no QR workspace, no forward differences.
// It converges to the same least-squares minimiser; see DESIGN.md "Parity definition".
// This is the GENERAL solver (every model; per-voxel fallback of the mono-exponential fast path in
// mono_fast.cuh), together with the shared pieces: arithmetic shim, packed pairs, models, echo table,
// the one-voxel driver fit_voxel and the fused epilogue.
//
// The header is written against a tiny portability shim (DFIT_HD, dfit::num<T>) so that the very
// same solver source can be compiled by g++ into the test-only `tests/hostsim` harness, which lets
// the CPU test-suite exercise the device arithmetic.  The product library contains device code only.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

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
  // Approximate reciprocal / rsqrt (MUFU, ~1 ulp): used only inside the solver (scaling, damping,
  // step computation), where an inexact step is corrected by the next iteration.  The model
  // evaluation and r2 use exact arithmetic.
  static DFIT_HD float rcp_(float v) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return 1.0f / v;
#endif
  }
  static DFIT_HD float rsqrt_(float v) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return 1.0f / sqrtf(v);
#endif
  }
  static DFIT_HD float exp_(float v) { return expf(v); }
  static DFIT_HD float log_(float v) { return logf(v); }
  // natural log for the log-linear INITIAL GUESS only (MUFU.LG2 + FMUL on the device)
  static DFIT_HD float log_fast(float v) {
#if defined(__CUDA_ARCH__)
    return __logf(v);
#else
    return logf(v);
#endif
  }
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
  static DFIT_HD double rcp_(double v) { return 1.0 / v; }
  static DFIT_HD double rsqrt_(double v) { return 1.0 / sqrt(v); }
  static DFIT_HD double exp_(double v) { return exp(v); }
  static DFIT_HD double log_(double v) { return log(v); }
  static DFIT_HD double log_fast(double v) { return log(v); }
  static DFIT_HD double sqrt_(double v) { return sqrt(v); }
  static DFIT_HD double abs_(double v) { return fabs(v); }
  static DFIT_HD double max_(double a, double b) { return fmax(a, b); }
  static DFIT_HD double min_(double a, double b) { return fmin(a, b); }
  static DFIT_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
  static DFIT_HD bool finite(double v) { return fabs(v) <= 1.7976931348623157e308; }
};

// ------------------------------------------------------------------------------------ pair2
// Two samples processed per instruction.  For float on sm_100a the arithmetic maps onto Blackwell's
// packed FP32 instructions (PTX fma/mul/add.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2): the FMA pipe does
// the same 128 FMA/clk/SM either way (profiles/microbench/pipe_rates.cu), but a packed instruction
// takes ONE issue slot for two FMAs, and this kernel is issue-bound.  Elsewhere (double, host) the
// pair is two scalar operations.
template <typename T>
struct pair2 {
  T lo, hi;
};

template <typename T>
DFIT_HD pair2<T> p2_make(T lo, T hi) {
  pair2<T> r;
  r.lo = lo;
  r.hi = hi;
  return r;
}
template <typename T>
DFIT_HD pair2<T> p2_bcast(T v) {
  return p2_make<T>(v, v);
}
template <typename T>
DFIT_HD pair2<T> p2_fma(pair2<T> a, pair2<T> b, pair2<T> c) {
  return p2_make<T>(num<T>::fma_(a.lo, b.lo, c.lo), num<T>::fma_(a.hi, b.hi, c.hi));
}
template <typename T>
DFIT_HD pair2<T> p2_mul(pair2<T> a, pair2<T> b) {
  return p2_make<T>(a.lo * b.lo, a.hi * b.hi);
}
template <typename T>
DFIT_HD pair2<T> p2_add(pair2<T> a, pair2<T> b) {
  return p2_make<T>(a.lo + b.lo, a.hi + b.hi);
}
#if defined(__CUDA_ARCH__)
template <>
__device__ __forceinline__ pair2<float> p2_fma<float>(pair2<float> a, pair2<float> b, pair2<float> c) {
  pair2<float> d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.lo), "=f"(d.hi)
      : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi), "f"(c.lo), "f"(c.hi));
  return d;
}
template <>
__device__ __forceinline__ pair2<float> p2_mul<float>(pair2<float> a, pair2<float> b) {
  pair2<float> d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.lo), "=f"(d.hi)
      : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi));
  return d;
}
template <>
__device__ __forceinline__ pair2<float> p2_add<float>(pair2<float> a, pair2<float> b) {
  pair2<float> d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.lo), "=f"(d.hi)
      : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi));
  return d;
}
#endif
// Packed add / multiply that are guaranteed not to be contracted with neighbouring operations (the device versions
// above are inline asm already; the host versions go through volatile temporaries).
DFIT_HD pair2<float> p2_add_rn(pair2<float> a, pair2<float> b) {
#if defined(__CUDA_ARCH__)
  return p2_add<float>(a, b);
#else
  volatile float lo = a.lo + b.lo, hi = a.hi + b.hi;
  return p2_make<float>(lo, hi);
#endif
}
DFIT_HD pair2<float> p2_mul_rn(pair2<float> a, pair2<float> b) {
#if defined(__CUDA_ARCH__)
  return p2_mul<float>(a, b);
#else
  volatile float lo = a.lo * b.lo, hi = a.hi * b.hi;
  return p2_make<float>(lo, hi);
#endif
}
template <typename T>
DFIT_HD pair2<T> p2_expbx(T b, pair2<T> x, pair2<T> xs) {
  return p2_make<T>(num<T>::expbx(b, x.lo, xs.lo), num<T>::expbx(b, x.hi, xs.hi));
}
#if defined(__CUDA_ARCH__)
template <>
__device__ __forceinline__ pair2<float> p2_expbx<float>(float b, pair2<float> /*x*/, pair2<float> xs) {
  const pair2<float> t = p2_mul<float>(p2_bcast<float>(b), xs);  // one FMUL2 for both arguments
  pair2<float> e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.lo) : "f"(t.lo));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.hi) : "f"(t.hi));
  return e;
}
#endif

// ------------------------------------------------------------------------------------ echo table
// Echo-time table shared by all voxels of a launch (lives in the kernel parameter / constant bank).
template <typename T, int EMAX>
struct XTab {
  T x[EMAX];   // echo / spin-lock times
  T xs[EMAX];  // x * log2(e), so exp(b x) = ex2(b * xs)
  T xc[EMAX];  // x - mean(x), for the log-linear initial guess
  T xbar, inv_sxx;
  // Uniform spacing x_k = x0 + k dx (multi-echo spin echo, cones, ...): exp(b x_k) = e0 q^k with
  // q = exp(b dx), so the mono-exponential model is a POLYNOMIAL in q (mono_uniform_newton).
  int uniform;   // 1: spacing is uniform to within the arithmetic's resolution (host decides)
  int backward;  // 1: dx < 0 (descending echo times): the Prony start is taken backwards
  T x0, x0s;     // first echo time and x0 * log2(e)
  T inv_dx;      // 1 / dx
  T dx2, dxs2;   // 2 dx and 2 dx * log2(e): exp(b x_k+2) = exp(b x_k) * exp(b * dx2)
  T q_lo, q_hi;  // admissible range of q: q^(2E-2) must stay finite
  T xx[EMAX];    // x^2, for the second derivatives of the general (non-uniform) fast path
  T inv_xmax;    // 1 / max |x|: the largest step in b the general fast path takes at once
  T span2;       // (max x - min x)^2
};

// Host-side fill of the echo table (shared by the C-ABI layer and the test-only host build).
template <typename T, int EMAX>
inline void fill_xtab(XTab<T, EMAX>& xt, const double* x, int n_echo) {
  double xbar = 0, sxx = 0;
  for (int e = 0; e < n_echo; ++e) xbar += x[e];
  xbar /= n_echo;
  for (int e = 0; e < EMAX; ++e) {
    const double xe = e < n_echo ? x[e] : 0.0;
    xt.x[e] = (T)xe;
    xt.xs[e] = (T)(xe * 1.4426950408889634074);
    xt.xc[e] = (T)(e < n_echo ? xe - xbar : 0.0);
    if (e < n_echo) sxx += (xe - xbar) * (xe - xbar);
  }
  xt.xbar = (T)xbar;
  xt.inv_sxx = (T)(sxx > 0 ? 1.0 / sxx : 0.0);
  const double dx = n_echo >= 2 ? (x[n_echo - 1] - x[0]) / (n_echo - 1) : 0.0;
  bool uni = n_echo >= 3 && dx != 0.0 && dx == dx;
  double xmax = 0;
  for (int e = 0; e < n_echo; ++e) xmax = fmax(xmax, fabs(x[e]));
  // "uniform" must not change the model: deviations from the grid have to be below what the
  // arithmetic type resolves in b * x (one ulp of the largest echo time)
  const double tol = (double)num<T>::eps() * xmax;
  for (int e = 0; e < n_echo && uni; ++e) uni = fabs(x[e] - (x[0] + e * dx)) <= tol;
  xt.uniform = uni ? 1 : 0;
  xt.backward = dx < 0 ? 1 : 0;
  xt.x0 = (T)(n_echo ? x[0] : 0.0);
  xt.x0s = (T)(n_echo ? x[0] * 1.4426950408889634074 : 0.0);
  xt.inv_dx = (T)(uni ? 1.0 / dx : 0.0);
  xt.dx2 = (T)(2.0 * dx);
  xt.dxs2 = (T)(2.0 * dx * 1.4426950408889634074);
  const double decades = sizeof(T) == 4 ? 30.0 : 280.0;
  xt.q_hi = (T)pow(10.0, decades / (2.0 * (n_echo > 1 ? n_echo : 2) - 2.0));
  xt.q_lo = (T)(sizeof(T) == 4 ? 1e-30 : 1e-280);
  for (int e = 0; e < EMAX; ++e) xt.xx[e] = (T)(e < n_echo ? x[e] * x[e] : 0.0);
  xt.inv_xmax = (T)(xmax > 0 ? 1.0 / xmax : 0.0);
  double xlo = n_echo ? x[0] : 0.0, xhi = xlo;
  for (int e = 1; e < n_echo; ++e) {
    xlo = fmin(xlo, x[e]);
    xhi = fmax(xhi, x[e]);
  }
  xt.span2 = (T)((xhi - xlo) * (xhi - xlo));
}

// ------------------------------------------------------------------------------------ models
// Each model provides, for one sample, the value f and an UNSCALED Jacobian row Jh[P]; the true
// Jacobian is J_i = cs_i * Jh_i with per-parameter column scales cs (colscale()) that do not depend
// on the sample.  Accumulating Jh^T Jh and applying cs once per pass saves a multiply per sample and
// column (d/db of a*exp(b x) is a * [x exp(b x)]).  LIN is the bit-mask of parameters the model is
// linear in (used for the variable-projection start).

// fitting.py:1016-1018  f = a * exp(b x)
struct MonoExp {
  static constexpr int P = 2;
  static DFIT_HD constexpr bool dup(int, int) { return false; }  // (see BiExp)
  static constexpr unsigned LIN = 0x1u;
  static constexpr bool MONO = true;  // eligible for the variable-projection Newton fast path
  template <typename T>
  static DFIT_HD void eval(const T (&p)[2], T x, T xs, T& f, T (&Jh)[2]) {
    T e = num<T>::expbx(p[1], x, xs);
    Jh[0] = e;
    Jh[1] = x * e;
    f = p[0] * e;
  }
  template <typename T>
  static DFIT_HD void eval2(const T (&p)[2], pair2<T> x, pair2<T> xs, pair2<T> yneg, pair2<T>& r, pair2<T> (&Jh)[2]) {
    const pair2<T> e = p2_expbx<T>(p[1], x, xs);
    Jh[0] = e;
    Jh[1] = p2_mul<T>(x, e);
    r = p2_fma<T>(p2_bcast<T>(p[0]), e, yneg);
  }
  template <typename T>
  static DFIT_HD void colscale(const T (&p)[2], T (&cs)[2]) {
    cs[0] = (T)1;
    cs[1] = p[0];
  }
  // Uniformly spaced echoes: exp(b x_k+2) = exp(b x_k) * exp(2 b dx), so after the first echo pair every further pair
  // costs ONE packed multiply instead of a packed multiply and two MUFU.EX2 -- the evaluation loop of eval_all is
  // bound by the MUFU pipe otherwise (4 per SM sub-partition and clock).  The rounding of the multiplier acts like a
  // relative perturbation of b by ~3e-7; see eval_all.
  static constexpr bool HAS_REC = true;
  template <typename T>
  struct Rec {
    pair2<T> e, m;
  };
  template <typename T, int EMAX>
  static DFIT_HD void rec_init(const T (&p)[2], const XTab<T, EMAX>& xt, Rec<T>& rc) {
    rc.e = p2_expbx<T>(p[1], p2_make<T>(xt.x[0], xt.x[EMAX > 1 ? 1 : 0]), p2_make<T>(xt.xs[0], xt.xs[EMAX > 1 ? 1 : 0]));
    rc.m = p2_bcast<T>(num<T>::expbx(p[1], xt.dx2, xt.dxs2));
  }
  template <typename T>
  static DFIT_HD void eval2u(const T (&p)[2], Rec<T>& rc, pair2<T> x, pair2<T> yneg, pair2<T>& r, pair2<T> (&Jh)[2]) {
    Jh[0] = rc.e;
    Jh[1] = p2_mul<T>(x, rc.e);
    r = p2_fma<T>(p2_bcast<T>(p[0]), rc.e, yneg);
    rc.e = p2_mul<T>(rc.e, rc.m);
  }
  template <typename T>
  static DFIT_HD void eval1u(const T (&p)[2], const Rec<T>& rc, T x, T& f, T (&Jh)[2]) {  // the odd last echo
    Jh[0] = rc.e.lo;
    Jh[1] = x * rc.e.lo;
    f = p[0] * rc.e.lo;
  }
};

// fitting.py:1021-1023  f = a1 exp(b1 x) + a2 exp(b2 x)
struct BiExp {
  static constexpr int P = 4;
  // Jh = [e1, x e1, e2, x e2]: the products Jh2 Jh1 = e2 (x e1) and Jh3 Jh0 = (x e2) e1 are the same number, so eval_all
  // accumulates entry (3, 0) of Jh^T Jh only and copies it into (2, 1)
  static DFIT_HD constexpr bool dup(int i, int j) { return i == 2 && j == 1; }
  static constexpr int DUP_DST = 2 * 3 / 2 + 1, DUP_SRC = 3 * 4 / 2 + 0;
  static constexpr unsigned LIN = 0x5u;
  static constexpr bool MONO = false;
  template <typename T>
  static DFIT_HD void eval(const T (&p)[4], T x, T xs, T& f, T (&Jh)[4]) {
    T e1 = num<T>::expbx(p[1], x, xs);
    T e2 = num<T>::expbx(p[3], x, xs);
    Jh[0] = e1;
    Jh[1] = x * e1;
    Jh[2] = e2;
    Jh[3] = x * e2;
    f = num<T>::fma_(p[0], e1, p[2] * e2);
  }
  template <typename T>
  static DFIT_HD void eval2(const T (&p)[4], pair2<T> x, pair2<T> xs, pair2<T> yneg, pair2<T>& r, pair2<T> (&Jh)[4]) {
    const pair2<T> e1 = p2_expbx<T>(p[1], x, xs);
    const pair2<T> e2 = p2_expbx<T>(p[3], x, xs);
    Jh[0] = e1;
    Jh[1] = p2_mul<T>(x, e1);
    Jh[2] = e2;
    Jh[3] = p2_mul<T>(x, e2);
    r = p2_fma<T>(p2_bcast<T>(p[0]), e1, p2_fma<T>(p2_bcast<T>(p[2]), e2, yneg));
  }
  template <typename T>
  static DFIT_HD void colscale(const T (&p)[4], T (&cs)[4]) {
    cs[0] = (T)1;
    cs[1] = p[0];
    cs[2] = (T)1;
    cs[3] = p[2];
  }
  static constexpr bool HAS_REC = true;  // (see MonoExp)
  template <typename T>
  struct Rec {
    pair2<T> e1, e2, m1, m2;
  };
  template <typename T, int EMAX>
  static DFIT_HD void rec_init(const T (&p)[4], const XTab<T, EMAX>& xt, Rec<T>& rc) {
    const pair2<T> x01 = p2_make<T>(xt.x[0], xt.x[EMAX > 1 ? 1 : 0]), xs01 = p2_make<T>(xt.xs[0], xt.xs[EMAX > 1 ? 1 : 0]);
    rc.e1 = p2_expbx<T>(p[1], x01, xs01);
    rc.e2 = p2_expbx<T>(p[3], x01, xs01);
    rc.m1 = p2_bcast<T>(num<T>::expbx(p[1], xt.dx2, xt.dxs2));
    rc.m2 = p2_bcast<T>(num<T>::expbx(p[3], xt.dx2, xt.dxs2));
  }
  template <typename T>
  static DFIT_HD void eval2u(const T (&p)[4], Rec<T>& rc, pair2<T> x, pair2<T> yneg, pair2<T>& r, pair2<T> (&Jh)[4]) {
    Jh[0] = rc.e1;
    Jh[1] = p2_mul<T>(x, rc.e1);
    Jh[2] = rc.e2;
    Jh[3] = p2_mul<T>(x, rc.e2);
    r = p2_fma<T>(p2_bcast<T>(p[0]), rc.e1, p2_fma<T>(p2_bcast<T>(p[2]), rc.e2, yneg));
    rc.e1 = p2_mul<T>(rc.e1, rc.m1);
    rc.e2 = p2_mul<T>(rc.e2, rc.m2);
  }
  template <typename T>
  static DFIT_HD void eval1u(const T (&p)[4], const Rec<T>& rc, T x, T& f, T (&Jh)[4]) {
    Jh[0] = rc.e1.lo;
    Jh[1] = x * rc.e1.lo;
    Jh[2] = rc.e2.lo;
    Jh[3] = x * rc.e2.lo;
    f = num<T>::fma_(p[0], rc.e1.lo, p[2] * rc.e2.lo);
  }
};

// f = a x : the 1-parameter custom model of the reference's tests (tests/core/test_fitting.py:52-53)
struct Linear1 {
  static constexpr int P = 1;
  static DFIT_HD constexpr bool dup(int, int) { return false; }
  static constexpr unsigned LIN = 0x1u;
  static constexpr bool MONO = false;
  template <typename T>
  static DFIT_HD void eval(const T (&p)[1], T x, T /*xs*/, T& f, T (&Jh)[1]) {
    Jh[0] = x;
    f = p[0] * x;
  }
  template <typename T>
  static DFIT_HD void eval2(const T (&p)[1], pair2<T> x, pair2<T> /*xs*/, pair2<T> yneg, pair2<T>& r, pair2<T> (&Jh)[1]) {
    Jh[0] = x;
    r = p2_fma<T>(p2_bcast<T>(p[0]), x, yneg);
  }
  template <typename T>
  static DFIT_HD void colscale(const T (&)[1], T (&cs)[1]) {
    cs[0] = (T)1;
  }
  static constexpr bool HAS_REC = false;  // no exponentials
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

template <class M, typename T, bool ON>
struct RecOf {
  typedef int type;
};
template <class M, typename T>
struct RecOf<M, T, true> {
  typedef typename M::template Rec<T> type;
};

// One pass over the echoes: cost F = sum (f - y)^2, A = J^T J (packed lower), g = J^T r.
// With an exact echo count and matching accumulator type the echoes are processed two at a time
// (pair2 -> packed FP32 instructions on sm_100a); partial sums of even and odd echoes are kept in
// the two halves and added at the end.
// MASK selects the parameters whose rows/columns are accumulated (all of them for an LM pass, only
// the linear ones for the projection pass, where the rest of the system is never read).
// UNI (models with exponentials, uniformly spaced echoes -- the caller checks xt.uniform): the exponentials come from
// the model's two-echo recurrence (M::Rec) instead of one MUFU.EX2 per echo and exponential.
template <class M, typename T, typename TA, int EMAX, bool EXACT, unsigned MASK = 0xffu, bool UNI = false>
DFIT_HD void eval_all(const T (&p)[M::P], const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, TA& F,
                      TA (&A)[M::P * (M::P + 1) / 2], TA (&g)[M::P]) {
  constexpr int P = M::P;
  constexpr int NA = P * (P + 1) / 2;
  const T* __restrict__ x = xt.x;
  const T* __restrict__ xs = xt.xs;
#define DFIT_ON(i) (((MASK) >> (i)) & 1u)
  constexpr bool PAIRED = EXACT && sizeof(T) == sizeof(TA) && EMAX >= 2;
  if constexpr (PAIRED) {
    pair2<T> F2 = p2_bcast<T>((T)0), A2[NA], g2[P];
#pragma unroll
    for (int k = 0; k < NA; ++k) A2[k] = p2_bcast<T>((T)0);
#pragma unroll
    for (int k = 0; k < P; ++k) g2[k] = p2_bcast<T>((T)0);
    [[maybe_unused]] typename RecOf<M, T, UNI && M::HAS_REC>::type rc;
    if constexpr (UNI && M::HAS_REC) M::template rec_init<T, EMAX>(p, xt, rc);
#pragma unroll
    for (int e = 0; e + 1 < EMAX; e += 2) {
      pair2<T> r, J[P];
      if constexpr (UNI && M::HAS_REC)
        M::template eval2u<T>(p, rc, p2_make<T>(x[e], x[e + 1]), p2_make<T>(-y[e], -y[e + 1]), r, J);
      else
        M::template eval2<T>(p, p2_make<T>(x[e], x[e + 1]), p2_make<T>(xs[e], xs[e + 1]), p2_make<T>(-y[e], -y[e + 1]), r,
                             J);
      F2 = p2_fma<T>(r, r, F2);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (DFIT_ON(i)) g2[i] = p2_fma<T>(J[i], r, g2[i]);
#pragma unroll
        for (int j = 0; j <= i; ++j)
          if (DFIT_ON(i) && DFIT_ON(j) && !M::dup(i, j)) A2[tri(i, j)] = p2_fma<T>(J[i], J[j], A2[tri(i, j)]);
      }
    }
    F = (TA)(F2.lo + F2.hi);
#pragma unroll
    for (int k = 0; k < NA; ++k) A[k] = (TA)(A2[k].lo + A2[k].hi);
#pragma unroll
    for (int k = 0; k < P; ++k) g[k] = (TA)(g2[k].lo + g2[k].hi);
    if constexpr (EMAX & 1) {
      T f, J[P];
      if constexpr (UNI && M::HAS_REC) M::template eval1u<T>(p, rc, x[EMAX - 1], f, J);
      else M::template eval<T>(p, x[EMAX - 1], xs[EMAX - 1], f, J);
      const T re = f - y[EMAX - 1];
      F = num<TA>::fma_((TA)re, (TA)re, F);
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (DFIT_ON(i)) g[i] = num<TA>::fma_((TA)J[i], (TA)re, g[i]);
#pragma unroll
        for (int j = 0; j <= i; ++j)
          if (DFIT_ON(i) && DFIT_ON(j) && !M::dup(i, j)) A[tri(i, j)] = num<TA>::fma_((TA)J[i], (TA)J[j], A[tri(i, j)]);
      }
    }
  } else {
    F = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k) A[k] = 0;
#pragma unroll
    for (int k = 0; k < P; ++k) g[k] = 0;
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (EXACT || e < E) {
        T f, J[P];
        M::template eval<T>(p, x[e], xs[e], f, J);
        const T re = f - y[e];
        F = num<TA>::fma_((TA)re, (TA)re, F);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          if (DFIT_ON(i)) g[i] = num<TA>::fma_((TA)J[i], (TA)re, g[i]);
#pragma unroll
          for (int j = 0; j <= i; ++j)
            if (DFIT_ON(i) && DFIT_ON(j) && !M::dup(i, j)) A[tri(i, j)] = num<TA>::fma_((TA)J[i], (TA)J[j], A[tri(i, j)]);
        }
      }
    }
  }
  if constexpr (M::dup(2, 1)) A[M::DUP_DST] = A[M::DUP_SRC];
  T cs[P];
  M::template colscale<T>(p, cs);
#pragma unroll
  for (int i = 0; i < P; ++i) {
    if (DFIT_ON(i)) g[i] *= (TA)cs[i];
#pragma unroll
    for (int j = 0; j <= i; ++j)
      if (DFIT_ON(i) && DFIT_ON(j)) A[tri(i, j)] *= (TA)cs[i] * (TA)cs[j];
  }
#undef DFIT_ON
}

// Solve (A + lam diag(dd)) z = -g for symmetric A (packed lower), all in registers: the Marquardt-scaled system
// (D^-1 A D^-1 + lam I)(D z) = -D^-1 g with D = sqrt(dd), solved WITHOUT forming the scaled matrix -- no rsqrt of the
// scales, no P (P + 1) multiplies to apply them (a Cholesky factorisation is invariant under diagonal scaling up to
// rounding; the positive-definiteness thresholds are relative to dd for the same reason).  Rows/columns whose bit is
// cleared in `active` are frozen (z_i = 0).  P <= 2 uses the closed form, larger systems an unrolled Cholesky.
// Reciprocals are approximate (see num<>::rcp_).
template <int P, typename TA>
DFIT_HD bool chol_solve(const TA (&A)[P * (P + 1) / 2], TA lam, const TA (&dd)[P], const TA (&g)[P], unsigned active,
                        TA (&z)[P]) {
  const TA tiny = num<TA>::eps() * (TA)4;
  if constexpr (P == 1) {
    const TA d = num<TA>::fma_(lam, dd[0], A[0]);
    z[0] = (active & 1u) ? -g[0] * num<TA>::rcp_(d) : (TA)0;
    return d > tiny * dd[0] || !(active & 1u);
  } else if constexpr (P == 2) {
    const bool a0 = active & 1u, a1 = (active >> 1) & 1u;
    const TA s0 = a0 ? dd[0] : (TA)1, s1 = a1 ? dd[1] : (TA)1;
    const TA d0 = a0 ? num<TA>::fma_(lam, dd[0], A[0]) : (TA)1, d1 = a1 ? num<TA>::fma_(lam, dd[1], A[2]) : (TA)1;
    const TA c = (a0 && a1) ? A[1] : (TA)0;
    const TA g0 = a0 ? g[0] : (TA)0, g1 = a1 ? g[1] : (TA)0;
    const TA det = d0 * d1 - c * c;
    const TA inv = num<TA>::rcp_(det);
    z[0] = (c * g1 - d1 * g0) * inv;
    z[1] = (c * g0 - d0 * g1) * inv;
    return det > tiny * d0 * d1 && d0 > tiny * s0 && d1 > tiny * s1;
  } else {
    TA L[P * (P + 1) / 2], Dinv[P];
    bool pd = true;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const bool aj = (active >> j) & 1u;
      TA s = aj ? num<TA>::fma_(lam, dd[j], A[tri(j, j)]) : (TA)1;
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[tri(j, k)] * L[tri(j, k)];
      // (no early exit: a pivot that is not positive poisons what follows, and the caller discards z when told so --
      // straight-line code instead of a branch per column)
      pd = pd && (s > tiny * (aj ? dd[j] : (TA)1));
      const TA inv = num<TA>::rsqrt_(s);
      Dinv[j] = inv;
#pragma unroll
      for (int i = j + 1; i < P; ++i) {
        const bool ai = (active >> i) & 1u;
        TA t = (aj && ai) ? A[tri(i, j)] : (TA)0;
#pragma unroll
        for (int k = 0; k < j; ++k) t -= L[tri(i, k)] * L[tri(j, k)];
        L[tri(i, j)] = t * inv;
      }
    }
    TA w[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      TA t = ((active >> i) & 1u) ? -g[i] : (TA)0;
#pragma unroll
      for (int k = 0; k < i; ++k) t -= L[tri(i, k)] * w[k];
      w[i] = t * Dinv[i];
    }
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {
      TA t = w[i];
#pragma unroll
      for (int k = i + 1; k < P; ++k) t -= L[tri(k, i)] * z[k];
      z[i] = t * Dinv[i];
    }
    return pd;
  }
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
      T l = num<T>::log_fast(v);  // NaN for v < 0
      sl += l;
      sxl = num<T>::fma_(xc[e], l, sxl);
    }
  }
  T slope = sxl * inv_sxx;
  T icpt = sl / (T)E - slope * xbar;
  T a = num<T>::exp_(icpt);
  if (num<T>::finite(slope) && num<T>::finite(a) && num<T>::finite(sl)) {
    p[0] = a;
    p[1] = slope;
  } else {
    p[0] = (T)1;
    p[1] = (T)0;
  }
}

// Marquardt-scaled step from the current normal equations: (A + lam diag(D2)) dp = -g with D2 the running maxima of
// diag(A) (MINPACK's diag rule).  Returns false if the (restricted) system is not positive definite.  Outputs the trial
// point pt, the squared scaled step |D dp|^2, the squared scaled norm of pt and the predicted reduction of the
// linearised model, dp^T A dp + 2 lam |D dp|^2 = lam |D dp|^2 - dp^T g (dp solves the system above; P multiply-adds
// instead of the P (P + 1) / 2 terms of the quadratic form, and both forms cancel to the same extent: by the
// condition number of the scaled matrix).
template <int P, typename T, typename TA>
DFIT_HD bool lm_step(const T (&p)[P], const TA (&A)[P * (P + 1) / 2], const TA (&g)[P], TA (&D2)[P], TA lam,
                     unsigned active, T (&pt)[P], TA& zz, TA& pnorm2, TA& pred) {
  TA dd[P], z[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    D2[i] = num<TA>::max_(D2[i], A[tri(i, i)]);  // running maximum: MINPACK's diag rule
    dd[i] = D2[i] > 0 ? D2[i] : (TA)1;
  }
  if (!chol_solve<P, TA>(A, lam, dd, g, active, z)) return false;
  zz = 0;
  pnorm2 = 0;
  TA zg = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    pt[i] = p[i] + (T)z[i];
    zz = num<TA>::fma_(dd[i] * z[i], z[i], zz);
    zg = num<TA>::fma_(z[i], g[i], zg);
    pnorm2 = num<TA>::fma_(D2[i] * (TA)pt[i], (TA)pt[i], pnorm2);
  }
  pred = lam * zz - zg;
  return true;
}

// The solver, in two pieces so that a kernel can run it in ROUNDS (fit_kernel_lmq): lm_begin evaluates the start
// point, lm_iterate runs the Levenberg-Marquardt trips and can stop -- right before a model evaluation, with the step
// to evaluate already computed -- once it has spent `budget` evaluations; called again with resume = true it carries on
// exactly where it stopped.  Everything the iteration carries from one trip to the next lives in LmState, so a
// suspended fit can be parked in shared memory and picked up by any lane; per voxel the arithmetic is the same, in
// the same order, however the trips are cut into rounds.  lm_solve is the two run back to back.
//
// Every pass (eval_all) yields the cost for the gain ratio AND the normal equations for the next
// step.  Two passes are saved relative to a textbook LM:
//   * the variable-projection start is a first step restricted to the linear parameters with zero
//     damping, taken from the linear parameters set to zero (its trial pass IS the first pass);
//   * a step whose predicted reduction is already below ftol*F is taken without evaluating it.
constexpr int ST_PENDING = -1;  // lm_iterate: the budget of this round is spent, the fit is not finished

template <int P, typename T, typename TA>
struct LmState {
  static constexpr int NA = P * (P + 1) / 2;
  TA F, A[NA], g[P], D2[P];  // cost, normal equations and running column scales at the accepted point
  TA lam, nu, ysq;           // damping, its growth factor, sum y^2
  T pt[P];                   // the trial point the next evaluation is wanted at ...
  TA zz, pnorm2, pred;       // ... its squared scaled step, squared scaled norm and predicted reduction
  int fev, iters;            // MINPACK-style budget spent, passes over the echoes spent
};

// Start: returns ST_PENDING (iterate from p) or a final status.
template <class M, typename T, typename TA, int EMAX, bool EXACT, bool UNI = false>
DFIT_HD int lm_begin(T (&p)[M::P], const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, const SolverOpts<T>& o,
                     LmState<M::P, T, TA>& s) {
  constexpr int P = M::P;
  constexpr int NA = P * (P + 1) / 2;
  s.iters = 0;
  s.fev = 0;
  TA ysq = 0;
#pragma unroll
  for (int e = 0; e < EMAX; ++e)
    if (EXACT || e < E) ysq = num<TA>::fma_((TA)y[e], (TA)y[e], ysq);
  s.ysq = ysq;
  s.F = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) s.D2[i] = 0;
  s.zz = s.pnorm2 = s.pred = 0;
  s.lam = (TA)o.lambda0;
  s.nu = 2;

  // Projection: with the linear parameters at zero the residual is -y and only the linear block of
  // the normal equations is needed (a cheap pass); its unregularised solution is the least-squares
  // optimum of the linear parameters for the given non-linear ones.  The projected point is then
  // evaluated in full; if that is not an improvement over sum y^2 (degenerate linear sub-problem),
  // or projection is off, the fit starts from p0 exactly as given.
  T plin[P], pt[P];
  bool projected = false;
#pragma unroll
  for (int i = 0; i < P; ++i) plin[i] = pt[i] = p[i];
  if (o.init_linear != 0) {
#pragma unroll
    for (int i = 0; i < P; ++i)
      if ((M::LIN >> i) & 1u) pt[i] = (T)0;
    TA F0, A0[NA], g0[P], D2p[P];
#pragma unroll
    for (int k = 0; k < NA; ++k) A0[k] = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      g0[i] = 0;
      D2p[i] = 0;
    }
    eval_all<M, T, TA, EMAX, EXACT, M::LIN, UNI>(pt, y, xt, E, F0, A0, g0);
    ++s.iters;
    s.fev += 1;
    T pp[P];
    projected = lm_step<P, T, TA>(pt, A0, g0, D2p, (TA)0, M::LIN, pp, s.zz, s.pnorm2, s.pred);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      projected = projected && num<T>::finite(pp[i]);
      pt[i] = pp[i];
    }
  }
  for (int trip = 0; trip < 2; ++trip) {
    if (!projected) {
#pragma unroll
      for (int i = 0; i < P; ++i) pt[i] = plin[i];
    }
    eval_all<M, T, TA, EMAX, EXACT, 0xffu, UNI>(pt, y, xt, E, s.F, s.A, s.g);
    ++s.iters;
    s.fev += 1 + P;
    const bool good = num<TA>::finite(s.F);
    if (projected && !(good && s.F <= ysq)) {
      projected = false;
      continue;
    }
    if (!good) return ST_NUMERIC;
    break;
  }
#pragma unroll
  for (int i = 0; i < P; ++i) p[i] = pt[i];
  return ST_PENDING;
}

// Levenberg-Marquardt trips from the state lm_begin (or an earlier, suspended lm_iterate: resume = true) left.
// Returns the final status, or ST_PENDING after `budget` evaluations with the fit still going.
template <class M, typename T, typename TA, int EMAX, bool EXACT, bool UNI = false>
DFIT_HD int lm_iterate(T (&p)[M::P], const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, const SolverOpts<T>& o,
                       LmState<M::P, T, TA>& s, int budget, bool resume) {
  constexpr int P = M::P;
  constexpr int NA = P * (P + 1) / 2;
  constexpr unsigned ALL = (1u << P) - 1u;
  const TA floorF = (TA)o.floor_rel * s.ysq;
  const TA ftol = (TA)o.ftol, xtol2 = (TA)o.xtol * (TA)o.xtol;
  const TA eps16 = (TA)16 * (TA)num<T>::eps();
  int status = ST_MAXITER;
  for (;;) {
    if (!resume) {
      if (!(s.fev < o.maxfev) | (s.F <= floorF)) break;  // (an exact fit reads ST_EXACT after the loop either way)
      bool solved = lm_step<P, T, TA>(p, s.A, s.g, s.D2, s.lam, ALL, s.pt, s.zz, s.pnorm2, s.pred);
      for (int tries = 0; tries < 12 && !solved; ++tries) {
        s.lam = num<TA>::max_(s.lam * (TA)10, (TA)1e-3);
        solved = lm_step<P, T, TA>(p, s.A, s.g, s.D2, s.lam, ALL, s.pt, s.zz, s.pnorm2, s.pred);
      }
      if (!solved) break;
      if (s.pred <= ftol * s.F && s.lam <= (TA)1) {
        // the step is below the tolerance before it is even evaluated: take it and stop
#pragma unroll
        for (int i = 0; i < P; ++i) p[i] = s.pt[i];
        status = s.zz <= xtol2 * s.pnorm2 ? ST_CONV_FX : ST_CONV_F;
        break;
      }
      if (budget <= 0) return ST_PENDING;  // suspended with the step computed: the next call evaluates it first
    }
    resume = false;
    --budget;
    TA Fn, An[NA], gn[P];
    eval_all<M, T, TA, EMAX, EXACT, 0xffu, UNI>(s.pt, y, xt, E, Fn, An, gn);
    ++s.iters;
    ++s.fev;
    const bool good = num<TA>::finite(Fn);
    const TA act = s.F - Fn;
    // F and Fn are sums of squares of model evaluations that each carry ~eps*|y| of rounding, so
    // their difference cannot resolve reductions below tau ~ eps*sqrt(sum y^2 * F).  Below that
    // level the gain ratio is noise: trust the (accurately computed) predicted reduction instead
    // of rejecting at random.  (Compared squared to avoid the square root.)
    // (Written with selects and non-short-circuit logic: on the device this is straight-line code; lanes of a warp
    // disagree about accept / reject all the time, and every branch here was a divergence point.)
    const TA tau2 = eps16 * eps16 * s.ysq * s.F;
    const bool reliable = s.pred * s.pred > tau2;
    const TA lhs = reliable ? act : act * num<TA>::abs_(act);
    const TA rhs = reliable ? (TA)1e-4 * s.pred : -tau2;
    const bool accept = good & (lhs > rhs);
    const TA rho = reliable ? act * num<TA>::rcp_(s.pred) : (TA)1;
    const TA ftolF = ftol * s.F;
    const bool small_f = good & (s.pred <= ftolF) & (!reliable | ((num<TA>::abs_(act) <= ftolF) & (act <= (TA)2 * s.pred)));
    const bool conv_x = accept & (s.zz <= xtol2 * s.pnorm2);
    {
      const TA t = (TA)2 * rho - (TA)1;
      const TA lam_acc = num<TA>::max_(s.lam * num<TA>::max_((TA)(1.0 / 3.0), (TA)1 - t * t * t), (TA)1e-9);
      const TA lam_rej = s.lam * s.nu;
      s.lam = accept ? lam_acc : lam_rej;
      s.nu = accept ? (TA)2 : s.nu * (TA)2;
      s.fev += accept ? P : 0;  // MINPACK re-differences its Jacobian after an accepted step: P more evaluations of its budget
      s.F = accept ? Fn : s.F;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        p[i] = accept ? s.pt[i] : p[i];
        s.g[i] = accept ? gn[i] : s.g[i];
      }
#pragma unroll
      for (int k = 0; k < NA; ++k) s.A[k] = accept ? An[k] : s.A[k];
    }
    if (small_f | conv_x) {
      status = small_f ? (conv_x ? ST_CONV_FX : ST_CONV_F) : ST_CONV_X;
      break;
    }
  }
  if (status == ST_MAXITER && s.F <= floorF) status = ST_EXACT;
  return status;
}

// On entry p holds the initial guess; on exit the accepted parameters.  Returns a Status; F_out is the sum of
// squared residuals at the returned point, iters the number of passes over the echoes (model + Jacobian
// evaluations) spent.
template <class M, typename T, typename TA, int EMAX, bool EXACT>
DFIT_HD int lm_solve(T (&p)[M::P], const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, const SolverOpts<T>& o, T& F_out,
                     int& iters) {
  LmState<M::P, T, TA> s;
  int status = lm_begin<M, T, TA, EMAX, EXACT>(p, y, xt, E, o, s);
  if (status == ST_PENDING) status = lm_iterate<M, T, TA, EMAX, EXACT>(p, y, xt, E, o, s, 0x7fffffff, false);
  F_out = (T)s.F;
  iters = s.iters;
  return status;
}

// ------------------------------------------------------------------------------------ one voxel
enum InitMode : int { INIT_GIVEN = 0, INIT_LOGLINEAR = 1 };
enum VoxelFlags : unsigned { FLAG_NONFINITE = 1u, FLAG_OOB = 2u };

template <typename T>
struct VoxelOpts {
  SolverOpts<T> s;
  T y_lo, y_hi;   // y_bounds (fitting.py:1065): any sample outside -> voxel skipped
  T r2_eps;       // fitting.py:763
  int init_mode;  // InitMode
  int has_bounds; // y_lo / y_hi are finite somewhere (otherwise the bounds test is skipped)
  int fast;       // 1: try the variable-projection Newton fast path first (mono-exponential only)
};

// Everything the reference does for one voxel (`_curve_fit`, fitting.py:1026-1073), on samples that
// are already in registers, in three steps so that kernels can put their own solver loop in the middle:
//   voxel_prepare  skip rules (:1065-1067), optional log-linear initial guess, non-finite check
//                  -> a final Status, or ST_PENDING: run the solver from p
//   (solver)       lm_solve, or lm_begin + rounds of lm_iterate
//   voxel_finish   r2 (:1032-1035) on success, NaN parameters and r2 = 0 otherwise (:1067, :1072)
template <class M, typename T, int EMAX, bool EXACT>
DFIT_HD int voxel_prepare(const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, const VoxelOpts<T>& vo, T (&p)[M::P],
                          unsigned& flags) {
  constexpr int P = M::P;
  bool all_zero = true, oob = false, nonfinite = false;
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (EXACT || e < E) {
      const T v = y[e];
      all_zero = all_zero && (v == (T)0);
      oob = oob || (v < vo.y_lo) || (v > vo.y_hi);
      nonfinite = nonfinite || !num<T>::finite(v);
    }
  }
  flags = 0;
  if (oob || all_zero) {
    if (oob) flags |= FLAG_OOB;
    return ST_SKIPPED;
  }
  if (vo.init_mode == INIT_LOGLINEAR && P == 2) {
    T q[2];
    loglinear_init<T, EMAX, EXACT>(y, xt.xc, xt.xbar, xt.inv_sxx, E, q);
    p[0] = q[0];
    p[P - 1] = q[1];
  }
#pragma unroll
  for (int i = 0; i < P; ++i) nonfinite = nonfinite || !num<T>::finite(p[i]);
  if (nonfinite) {
    flags |= FLAG_NONFINITE;
    return ST_NONFINITE;
  }
  return ST_PENDING;
}

template <class M, typename T, int EMAX, bool EXACT>
DFIT_HD void voxel_finish(int status, const T (&y)[EMAX], int E, const VoxelOpts<T>& vo, T F, T (&p)[M::P], T& r2) {
  constexpr int P = M::P;
  if (status >= ST_CONV_F && status <= ST_EXACT) {
    T ysum = 0;
#pragma unroll
    for (int e = 0; e < EMAX; ++e)
      if (EXACT || e < E) ysum += y[e];
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
}

// p: in = initial guess, out = fitted parameters (NaN on skip/failure).
template <class M, typename T, typename TA, int EMAX, bool EXACT>
DFIT_HD int fit_voxel(const T (&y)[EMAX], const XTab<T, EMAX>& xt, int E, const VoxelOpts<T>& vo, T (&p)[M::P],
                      T& r2, int& iters, unsigned& flags) {
  iters = 0;
  T F = 0;
  int status = voxel_prepare<M, T, EMAX, EXACT>(y, xt, E, vo, p, flags);
  if (status == ST_PENDING) status = lm_solve<M, T, TA, EMAX, EXACT>(p, y, xt, E, vo.s, F, iters);
  voxel_finish<M, T, EMAX, EXACT>(status, y, E, vo, F, p, r2);
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
  double scale[4], inv_scale[4];  // 10^decimals and its reciprocal, filled by set_post_scales()
  // fp32 twins for parameters whose post-processing is comparisons only (no ufunc, no rounding): a float v
  // satisfies (double)v < lb exactly when v < lbf with lbf the smallest float >= lb, and so on, so the
  // epilogue of such a parameter is exact in fp32.  Filled by set_post_scales().
  int simple[4];
  float lbf[4], ubf[4], r2_thresh_f, fill_f;
  // fp32 plan for the MonoExponentialFit column (ufunc 1 / |v| with bounds, threshold, fill and rounding,
  // fitting.py:725-737): `fastinv` set when every decision of the float64 evaluation can be reproduced from
  // fp32 quantities -- x = |v| is in bounds exactly when xa <= x <= xb (the floats between which
  // fl64(1 / x) lies inside [lb, ub]); the rounding is decided on a two-float quotient.  Filled by set_post_scales().
  int fastinv[4];
  float inv_xa[4], inv_xb[4], scale_f[4], inv_scale_f[4], fill_rounded_f[4];
};

// 1 / v in double.  On the device: MUFU.RCP of the fp32 image of v, refined by two Newton steps in fp64
// (relative error ~1e-28 before the final rounding, i.e. the correctly rounded quotient or its neighbour) --
// a generic fp64 division costs ~25 instructions and this epilogue runs once per voxel and parameter.
// The parameters it is applied to carry fp32 / LM-tolerance errors that are nine decades larger.
DFIT_HD double fast_recip(double v) {
#if defined(__CUDA_ARCH__)
  const double av = fabs(v);
  if (av > 1e-30 && av < 1e30) {
    float yf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)v));
    double y = (double)yf;
    y = fma(y, fma(-v, y, 1.0), y);
    y = fma(y, fma(-v, y, 1.0), y);
    return y;
  }
#endif
  return 1.0 / v;
}

DFIT_HD double apply_ufunc(int id, double v) {
  // (an if-chain on purpose: a switch becomes an indirect branch through a jump table on the device, and the
  // three reciprocal kinds share one reciprocal)
  if (id == UF_NONE) return v;
  if (id == UF_ABS) return fabs(v);
  const double r = fast_recip(id == UF_INV_ABS ? fabs(v) : v);
  return id == UF_NEG_INV ? -r : r;
}

DFIT_HD double pow10i(int d) {
  double s = 1.0;
  for (int k = 0; k < d; ++k) s *= 10.0;
  return s;
}

inline float float_at_least(double v) {  // smallest float >= v (v itself when it is +-inf / NaN)
  float f = (float)v;
  if ((double)f < v) f = nextafterf(f, INFINITY);
  return f;
}
inline float float_at_most(double v) {  // largest float <= v
  float f = (float)v;
  if ((double)f > v) f = nextafterf(f, -INFINITY);
  return f;
}

constexpr float kInvXMax = 1e30f;  // largest |v| the fp32 form of the 1 / |v| epilogue handles itself

// Smallest non-negative float (by bit pattern, 0 .. +inf) for which `pred` holds; pred must be monotone
// (false ... false true ... true).  Returns NaN when it never holds.
template <class Pred>
inline float first_float_where(Pred pred) {
  uint32_t lo = 0u, hi = 0x7f800000u;  // +0 .. +inf
  auto as_float = [](uint32_t u) {
    float f;
    memcpy(&f, &u, sizeof(f));
    return f;
  };
  if (!pred(as_float(hi))) return NAN;
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (pred(as_float(mid))) hi = mid;
    else lo = mid + 1;
  }
  return as_float(lo);
}

inline void set_post_scales(PostOpts& po) {
  po.fill_f = (float)po.fill;
  const bool fill_ok = !po.has_fill || (double)po.fill_f == po.fill;
  po.r2_thresh_f = float_at_least(po.r2_thresh);
  for (int i = 0; i < 4; ++i) {
    po.scale[i] = po.decimals[i] >= 0 ? pow10i(po.decimals[i]) : 1.0;
    po.inv_scale[i] = 1.0 / po.scale[i];
    po.simple[i] = (po.ufunc[i] == UF_NONE && po.decimals[i] < 0 && fill_ok) ? 1 : 0;
    po.lbf[i] = float_at_least(po.lb[i]);
    po.ubf[i] = float_at_most(po.ub[i]);
    // 1 / |v| in fp32 with the float64 evaluation's decisions
    po.fastinv[i] = 0;
    po.inv_xa[i] = po.inv_xb[i] = po.fill_rounded_f[i] = 0.f;
    po.scale_f[i] = (float)po.scale[i];
    po.inv_scale_f[i] = (float)po.inv_scale[i];
    if (po.ufunc[i] == UF_INV_ABS && po.decimals[i] <= 10) {
      const double lb = po.lb[i], ub = po.ub[i];
      // fl64(1 / x) decreases with x: in bounds for xa <= x <= xb
      const float xa = first_float_where([&](float x) { return !(1.0 / (double)x > ub); });
      const float xb1 = first_float_where([&](float x) { return 1.0 / (double)x < lb; });  // first x out again
      double fr = po.fill;
      if (po.has_fill && po.decimals[i] >= 0 && fabs(fr) < 1e300) fr = nearbyint(fr * po.scale[i]) / po.scale[i];
      const bool fr_ok = !po.has_fill || ((double)(float)fr == fr);
      // (x stays a normal, finite float and 10^d / x below 2^22 for every x in bounds: no per-voxel range checks)
      // (bounds-wise the upper end may be anything up to +inf; the fp32 arithmetic is used up to kInvXMax and the
      // float64 form beyond it)
      const bool range_ok = xa == xa && xa >= 1e-30f && xa < kInvXMax && po.scale[i] / (double)xa < 4.0e6;
      if (range_ok && fr_ok) {
        po.inv_xa[i] = xa;
        po.inv_xb[i] = xb1 != xb1 ? INFINITY : nextafterf(xb1, -INFINITY);  // (never out again: up to +inf)
        po.fill_rounded_f[i] = (float)fr;
        po.fastinv[i] = 1;
      }
    }
  }
}

// m / s for an integer-valued m and s = 10^d, correctly rounded like numpy's division in `around`:
// q = m (1/s), r = m - q s (exact, one fma), q + r (1/s) is the correctly rounded quotient (Markstein's
// correction step; checked exhaustively against IEEE division for |m| < 2^27 and d = 1..4).
DFIT_HD double div_pow10(double m, double s, double rs) {
#if defined(__CUDA_ARCH__)
  if (fabs(m) < 134217728.0 && s <= 10000.0) {
    const double q = m * rs;
    return fma(fma(-q, s, m), rs, q);
  }
#endif
  (void)rs;
  return m / s;
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
    const double s = po.scale[i];
    v = div_pow10(rint(v * s), s, po.inv_scale[i]);  // numpy.around: round-half-even of the scaled value, divided back
  }
  return v;
}


// a * b rounded to float and never contracted into a following add (the two-float arithmetic below relies on
// the ROUNDED product: its companion term compensates exactly that rounding)
DFIT_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float t = a * b;
  return t;
#endif
}

// top / x, x positive and finite, as two floats hi + lo (relative error ~2^-45): MUFU reciprocal, one Newton step,
// quotient and exact remainder by fma.  `top` = 1 or a power of ten.
DFIT_HD void quotient2(float top, float x, float& hi, float& lo) {
  float r = num<float>::rcp_(x);
  r = fmaf(r, fmaf(-x, r, 1.0f), r);
  hi = mul_rn(top, r);
  lo = mul_rn(fmaf(-hi, x, top), r);
}

// Epilogue of one fp32 parameter into an fp32 map: comparisons-only parameters stay in fp32 (exactly the
// decisions the float64 evaluation takes, see PostOpts::simple), and so does the MonoExponentialFit column
// 1 / |v| with bounds, r2 threshold, fill and rounding (PostOpts::fastinv): the value returned is the float64
// epilogue's result converted to float, bit for bit.  Everything else goes through post_param.
DFIT_HD float post_param_f32(const PostOpts& po, int i, float v, float r2) {
  if (!po.enabled) return v;
  if (po.simple[i] && !(fabsf(v) > 3.4028234e38f)) {  // finite or NaN (+-inf takes the nan_to_num branch in double)
    const bool bad = v < po.lbf[i] || v > po.ubf[i] || (po.has_r2_thresh && r2 < po.r2_thresh_f);
    if (bad) v = NAN;
    if (po.has_fill && v != v) v = po.fill_f;
    return v;
  }
  if (po.fastinv[i]) {
    const float x = fabsf(v);
    // fitting.py:130-141: outside the bounds (decided on x, see set_post_scales; false for NaN) or below the r2 threshold
    const bool good = x >= po.inv_xa[i] && x <= po.inv_xb[i] && !(po.has_r2_thresh && r2 < po.r2_thresh_f);
    if (!good) {
      if (po.has_fill) return po.fill_rounded_f[i];  // fitting.py:143-144, then :736-737
      if (x == x) return NAN;                        // (a NaN parameter keeps its payload through post_param)
    } else if (x <= kInvXMax) {
      const float S = po.decimals[i] >= 0 ? po.scale_f[i] : 1.0f;
      float hi, lo;
      quotient2(S, x, hi, lo);
      if (po.decimals[i] < 0) {
        // fl32(fl64(1 / x)): hi is within an ulp, lo says on which side of hi the quotient lies
        const float up = nextafterf(hi, INFINITY), dn = nextafterf(hi, -INFINITY);
        const float h_up = 0.5f * (up - hi), h_dn = 0.5f * (hi - dn);
        if (lo > h_up) return up;
        if (-lo > h_dn) return dn;
        if (lo != h_up && -lo != h_dn) return hi;
      } else {
        // np.around(1 / x, d) = rint(S / x) / S: the integer is decided on hi + lo (a quotient that is not a tie
        // is at least 2^-25 away from one, far more than either evaluation's error), ties go to post_param
        float m = rintf(hi);
        const float f = (hi - m) + lo;
        if (fabsf(f) != 0.5f) {
          if (f > 0.5f) m += 1.0f;
          else if (f < -0.5f) m -= 1.0f;
          // m / S correctly rounded (Markstein's correction): equals fl32(fl64(m / S))
          const float q = mul_rn(m, po.inv_scale_f[i]);
          return fmaf(fmaf(-q, S, m), po.inv_scale_f[i], q);
        }
      }
    }
  }
  return (float)post_param(po, i, (double)v, (double)r2);
}

// The same epilogue for the two voxels of a lane (parameter i of both): for the rounded 1 / |v| column the
// two-float quotient, the rounding (round-to-nearest-even by adding and subtracting 1.5 * 2^23) and the division by
// 10^d run on packed pairs; only the comparisons are per voxel.  Bit-identical to post_param_f32 on each half.
DFIT_HD pair2<float> post_pair_f32(const PostOpts& po, int i, pair2<float> v, pair2<float> r2) {
  if (!po.enabled) return v;
  if (po.fastinv[i] && po.decimals[i] >= 0) {
    typedef pair2<float> V;
    const V x = p2_make<float>(fabsf(v.lo), fabsf(v.hi));
    const bool good_lo = x.lo >= po.inv_xa[i] && x.lo <= po.inv_xb[i] && !(po.has_r2_thresh && r2.lo < po.r2_thresh_f);
    const bool good_hi = x.hi >= po.inv_xa[i] && x.hi <= po.inv_xb[i] && !(po.has_r2_thresh && r2.hi < po.r2_thresh_f);
    const V S = p2_bcast<float>(po.scale_f[i]), rS = p2_bcast<float>(po.inv_scale_f[i]);
    const V nx = p2_mul<float>(x, p2_bcast<float>(-1.0f));
    V r = p2_make<float>(num<float>::rcp_(x.lo), num<float>::rcp_(x.hi));
    r = p2_fma<float>(r, p2_fma<float>(nx, r, p2_bcast<float>(1.0f)), r);
    const V hi = p2_mul_rn(S, r);
    const V lo = p2_mul_rn(p2_fma<float>(p2_mul<float>(hi, p2_bcast<float>(-1.0f)), x, S), r);
    const V C = p2_bcast<float>(12582912.0f), nC = p2_bcast<float>(-12582912.0f);  // 1.5 * 2^23: rint for |t| < 2^22
    V m = p2_add_rn(p2_add_rn(hi, C), nC);
    const V f = p2_add_rn(p2_add_rn(hi, p2_mul<float>(m, p2_bcast<float>(-1.0f))), lo);
    const bool tie = fabsf(f.lo) == 0.5f || fabsf(f.hi) == 0.5f || x.lo > kInvXMax || x.hi > kInvXMax;  // (or out of the fp32 form's range)
    m = p2_add_rn(m, p2_add_rn(p2_add_rn(f, C), nC));  // + rint(f): +-1 when the low part carries over a half
    const V q = p2_mul_rn(m, rS);
    const V out = p2_fma<float>(p2_fma<float>(p2_mul<float>(q, p2_bcast<float>(-1.0f)), S, m), rS, q);
    // (out of bounds without a fill value: NaN, through the one-voxel form so that a NaN parameter keeps its payload)
    if (!tie && (good_lo || po.has_fill) && (good_hi || po.has_fill)) {
      return p2_make<float>(good_lo ? out.lo : po.fill_rounded_f[i], good_hi ? out.hi : po.fill_rounded_f[i]);
    }
  }
  return p2_make<float>(post_param_f32(po, i, v.lo, r2.lo), post_param_f32(po, i, v.hi, r2.hi));
}

}  // namespace dfit
