// Mono-exponential fast path: variable-projection Newton iteration on the projected cost.  Two voxels per lane,
// straight line, exactly two passes (mono_uniform_fast2 for uniformly spaced echoes, mono_general_fast2 for arbitrary
// echo times; entry point fit_voxel_fast2s) is what the dense kernels run; the loop forms (mono_uniform_newton, one
// voxel per lane; mono_general_newton2; entry point fit_voxel_fast) take the voxels those turn down, before the
// general LM of lm_core.cuh.  See DESIGN.md section 3.
//
// Like lm_core.cuh this header compiles for the device and, through the same portability shim, for the
// test-only host build (tests/hostsim).
#pragma once

#include "lm_core.cuh"

namespace dfit {

// ------------------------------------------------------------------------------------ one voxel per lane
// Mono-exponential fit on UNIFORMLY spaced echoes x_k = x0 + k dx, by variable projection.
//
// With q = exp(b dx) and a' = a exp(b x0) the model is a' q^k: a polynomial in q.  For a given q the
// optimal amplitude is a'(q) = N(q) / D(q) with N = sum_k y_k q^k and D = sum_k q^2k, and the projected
// cost is phi(q) = sum y^2 - N^2 / D.  Its stationary points are exactly those of the two-parameter
// least-squares problem the reference hands to MINPACK (fitting.py:1044-1062), so an exact Newton
// iteration on phi converges (quadratically) to the same minimiser -- with no exponentials at all in the
// loop: N, N', N'' and D, D', D'' come from one packed Horner recurrence (lo half: coefficients y_k,
// multiplier q; hi half: coefficients 1, multiplier q^2), 3E - 6 packed FMAs per pass.
//
// The start is the linear-prediction (Prony) estimate q0 = sum y_k y_k+1 / sum y_k^2, which is within a
// few per cent of the minimiser on decaying signals, so two to three passes suffice; the user's p0 only
// selects the basin MINPACK would start in and is not needed (the general LM of lm_core.cuh, which honours it,
// takes over whenever this path declines).  Convergence is judged like the LM's "predicted reduction
// <= ftol * F" test, applied to the error expected AFTER the step about to be taken: with the contraction
// kappa of newton_contraction() that is kappa^2 * pred <= ftol * F.
//
// Returns a Status (ST_CONV_F / ST_EXACT) or -1 when the path declines (no admissible start, curvature
// not positive, not converged in kMonoFastPasses, non-finite data): the caller then runs the general path.
constexpr int kMonoFastPasses = 6;
constexpr float kFirstStepCap = 0.2f;  // largest relative first Newton step in q the fast path accepts

// Expected error after the step about to be taken, relative to that step.  Newton's iteration contracts
// quadratically, e_k+1 = C e_k^2, and the last two steps estimate C = |dq_k| / dq_k-1^2, so the error left
// after taking dq_k is about (dq_k / dq_k-1)^2 |dq_k|.  A safety factor of 4 and a floor of 1e-3 (the size
// of C |dq| for a step that is about to be accepted) guard against a ratio that is small by accident.
template <typename T>
DFIT_HD T newton_contraction(T step2, T prev_step2) {
  return num<T>::min_(num<T>::max_((T)4 * step2 * num<T>::rcp_(prev_step2), (T)1e-3), (T)1);
}


// On the device the Newton loop is WARP-UNIFORM: every lane that entered together keeps iterating (with
// its state frozen once it has converged or declined) until all of them are finished, so the warp stays
// converged and the code after the loop runs once per warp instead of once per exit pass.
#if defined(__CUDA_ARCH__)
#define DFIT_LANES() __activemask()
#define DFIT_ANY(mask, pred) (__any_sync((mask), (pred)) != 0)
#else
#define DFIT_LANES() 0u
#define DFIT_ANY(mask, pred) (pred)
#endif

template <typename T, int E>
DFIT_HD int mono_uniform_newton(const T (&y)[E], const XTab<T, E>& xt, const SolverOpts<T>& o, T (&p)[2], T& F_out,
                                int& iters) {
  static_assert(E >= 3, "needs at least three echoes");
  typedef num<T> nm;
  const unsigned lanes = DFIT_LANES();
  (void)lanes;
  // Prony start; sum y^2 falls out of the same recurrence
  pair2<T> nd = p2_mul<T>(p2_bcast<T>(y[0]), p2_make<T>(y[1], y[0]));
#pragma unroll
  for (int k = 1; k + 1 < E; ++k) nd = p2_fma<T>(p2_bcast<T>(y[k]), p2_make<T>(y[k + 1], y[k]), nd);
  const T ysq = nm::fma_(y[E - 1], y[E - 1], nd.hi);
  // descending echo times: a decaying signal grows with the echo index; it is predicted backwards so that the large
  // samples carry the estimate (a per-launch choice)
  const T pdb = nm::fma_(-y[0], y[0], ysq);
  T q = xt.backward != 0 ? pdb * nm::rcp_(nd.lo) : nd.lo * nm::rcp_(nd.hi);
  // no admissible start (also catches NaN and all-zero voxels): this lane declines, but keeps in step
  bool active = q > xt.q_lo && q < xt.q_hi && nm::finite(ysq);
  if (!active) q = (T)0.5;

  const T floorF = o.floor_rel * ysq;
  T dprev2 = 0, qf = 0, af = 0;
  bool done = false;
  int npass = 0;
#pragma unroll 1
  for (int k = 0; k < kMonoFastPasses; ++k) {
    if (!DFIT_ANY(lanes, active)) break;
    const T s = q * q;
    const pair2<T> m = p2_make<T>(q, s);
    // Horner with first and (half) second derivative: lo = N(q) chain, hi = D(s) chain
    pair2<T> P0 = p2_make<T>(y[E - 1], (T)1), P1, P2;
    P1 = P0;
    P0 = p2_fma<T>(P0, m, p2_make<T>(y[E - 2], (T)1));
    P2 = P1;
    P1 = p2_fma<T>(P1, m, P0);
    P0 = p2_fma<T>(P0, m, p2_make<T>(y[E - 3], (T)1));
#pragma unroll
    for (int j = E - 4; j >= 0; --j) {
      P2 = p2_fma<T>(P2, m, P1);
      P1 = p2_fma<T>(P1, m, P0);
      P0 = p2_fma<T>(P0, m, p2_make<T>(y[j], (T)1));
    }
    const T N = P0.lo, N1 = P1.lo, N2h = P2.lo;  // N, dN/dq, (d2N/dq2) / 2
    const T D = P0.hi;                           // D(s), s = q^2
    const T D1 = (T)2 * q * P1.hi;               // dD/dq
    const T D2 = nm::fma_((T)8 * s, P2.hi, (T)2 * P1.hi);  // d2D/dq2
    const T rD = nm::rcp_(D);
    const T a = N * rD;                    // projected amplitude a'(q)
    const T u = nm::fma_(a, D1, -N1);      // = -D da'/dq
    const T ap = -u * rD;                  // da'/dq
    const T g = a * (u - N1);              // dphi/dq
    const T h = nm::fma_(a, nm::fma_(a, D2, (T)-4 * N2h), (T)2 * u * ap);  // d2phi/dq2
    T dq = -g * nm::rcp_(h);
    const T pred = (T)-0.5 * g * dq;       // Newton decrement: predicted reduction of phi
    const T Fest = nm::max_(nm::fma_(-N, a, ysq), (T)0);
    const T step2 = dq * dq;
    const T kappa = k == 0 ? (T)1 : newton_contraction<T>(step2, dprev2);
    // h > 0: inside the convex basin (false for NaN as well); otherwise the lane declines
    const bool convex = h > (T)0 && (k != 0 || step2 <= (T)(kFirstStepCap * kFirstStepCap) * q * q);  // see newton_lane_step
    const bool conv = convex && (pred * kappa) * kappa <= nm::fma_(o.ftol, Fest, floorF);
    dq = nm::min_(nm::max_(dq, (T)-0.5 * q), q);  // keep q positive whatever happens
    if (active) {
      npass = k + 1;
      if (conv) {
        qf = q + dq;
        af = nm::fma_(ap, dq, a);
        done = true;
      }
      active = convex && !conv;
      if (active) {
        q += dq;
        dprev2 = step2;
      }
    }
  }
  iters = npass;
  if (!done || !(qf > xt.q_lo && qf < xt.q_hi) || !nm::finite(af)) return -1;

  // cost at the returned point: r_k = y_k - a' q^k, powers two at a time
  T F;
  {
    pair2<T> ee = p2_make<T>((T)1, qf);
    const pair2<T> ss = p2_bcast<T>(qf * qf), na = p2_bcast<T>(-af);
    pair2<T> F2 = p2_bcast<T>((T)0);
#pragma unroll
    for (int e = 0; e + 1 < E; e += 2) {
      const pair2<T> r = p2_fma<T>(na, ee, p2_make<T>(y[e], y[e + 1]));
      F2 = p2_fma<T>(r, r, F2);
      if (e + 2 < E) ee = p2_mul<T>(ee, ss);
    }
    F = F2.lo + F2.hi;
    if constexpr (E & 1) {
      const T r = nm::fma_(-af, ee.lo, y[E - 1]);
      F = nm::fma_(r, r, F);
    }
  }
  // back to the reference's parameters: b = ln(q) / dx, a = a' exp(-b x0)
  const T b = nm::log_(qf) * xt.inv_dx;
  p[1] = b;
  p[0] = xt.x0 != (T)0 ? af * nm::expbx(-b, xt.x0, xt.x0s) : af;
  F_out = F;
  return F <= floorF ? ST_EXACT : ST_CONV_F;
}

// ---- two voxels per lane ---------------------------------------------------------------------------
// The same iteration with the two halves of every packed operation holding two VOXELS (lo = voxel A,
// hi = voxel B) instead of the N and D chains of one voxel: the Horner work per voxel is unchanged, but the
// Newton algebra, the convergence test, the residuals, the logarithm and the r2 all run two voxels per
// instruction.  Only MUFU, min/max, compares and selects remain per voxel.

// ln(v), v positive and normal, both halves: v = m 2^e with m in [sqrt(1/2), sqrt(2)), ln v = e ln 2 +
// log1p(m - 1); log1p(f) = f - f^2/2 + f^3 P(f), P a degree-7 fit on the interval (relative error 8e-8).
template <typename T>
DFIT_HD pair2<T> p2_log_pos(pair2<T> v) {
  return p2_make<T>(num<T>::log_(v.lo), num<T>::log_(v.hi));
}
#if defined(__CUDA_ARCH__)
template <>
__device__ __forceinline__ pair2<float> p2_log_pos<float>(pair2<float> v) {
  const int il = __float_as_int(v.lo), ih = __float_as_int(v.hi);
  const int el = (il - 0x3f3504f3) >> 23, eh = (ih - 0x3f3504f3) >> 23;
  const pair2<float> m = p2_make<float>(__int_as_float(il - (el << 23)), __int_as_float(ih - (eh << 23)));
  const pair2<float> f = p2_add<float>(m, p2_bcast<float>(-1.0f));
  pair2<float> t = p2_bcast<float>(-0.0790274366736412f);
  t = p2_fma<float>(t, f, p2_bcast<float>(0.12622319161891937f));
  t = p2_fma<float>(t, f, p2_bcast<float>(-0.12998183071613312f));
  t = p2_fma<float>(t, f, p2_bcast<float>(0.14214496314525604f));
  t = p2_fma<float>(t, f, p2_bcast<float>(-0.166412815451622f));
  t = p2_fma<float>(t, f, p2_bcast<float>(0.20001044869422913f));
  t = p2_fma<float>(t, f, p2_bcast<float>(-0.25000306963920593f));
  t = p2_fma<float>(t, f, p2_bcast<float>(0.3333333134651184f));
  const pair2<float> f2 = p2_mul<float>(f, f);
  const pair2<float> r = p2_add<float>(f, p2_fma<float>(p2_mul<float>(t, f), f2, p2_mul<float>(f2, p2_bcast<float>(-0.5f))));
  return p2_fma<float>(p2_make<float>((float)el, (float)eh), p2_bcast<float>(0.69314718055994531f), r);
}
#endif

template <bool B>
struct FirstPass {
  static constexpr bool value = B;
};

// Per-voxel (non-packable) part of one Newton pass: convergence test, step clamp, state update.
// The lane's life cycle is encoded in `dprev2` (no boolean registers to juggle): >= 0 while it iterates
// (the squared previous step), kLaneDone once it has converged, kLaneDeclined once it has given up.  `q`
// always holds the latest iterate: the step that ends the iteration is taken like any other.
constexpr float kLaneDone = -1.0f, kLaneDeclined = -2.0f;

template <typename T>
struct NewtonLane {
  T q, dprev2, af;
  int npass;
  DFIT_HD bool active() const { return dprev2 >= (T)0; }
  DFIT_HD bool done() const { return dprev2 == (T)kLaneDone; }
  DFIT_HD void start(T q0, bool ok, T fallback) {
    q = ok ? q0 : fallback;
    dprev2 = ok ? (T)0 : (T)kLaneDeclined;
    af = (T)0;
    npass = 0;
  }
};

template <bool FIRST, typename T>
DFIT_HD void newton_lane_step(NewtonLane<T>& L, int k, T h, T pred2, T tol2, T dq, T a, T ap, T step_lo, T step_hi,
                              T first_cap) {
  typedef num<T> nm;
  const bool act = L.active();
  const T step2 = dq * dq;
  T kappa = (T)1;
  // h > 0: inside the convex basin (false for NaN as well).
  bool convex = h > (T)0;
  if constexpr (FIRST) {
    // A first step beyond first_cap says the data-driven start is not near a minimum (low-SNR voxels with
    // several local minima): decline, so that the LM decides from the caller's p0 like the reference does.
    convex = convex && step2 <= first_cap * first_cap;
  } else {
    kappa = newton_contraction<T>(step2, L.dprev2);
  }
  const bool conv = convex && (pred2 * kappa) * kappa <= tol2;
  dq = nm::min_(nm::max_(dq, step_lo), step_hi);  // trust clamp (keeps q positive / b x bounded)
  L.q = (act && convex) ? L.q + dq : L.q;
  L.af = (act && conv) ? nm::fma_(ap, dq, a) : L.af;
  L.npass = act ? k + 1 : L.npass;
  L.dprev2 = act ? (conv ? (T)kLaneDone : (convex ? step2 : (T)kLaneDeclined)) : L.dprev2;
}

// ---- general echo times ------------------------------------------------------------------------------
// The same variable-projection Newton iteration for ARBITRARY echo times (T1rho spin-lock times, 10/20/40/80
// ms protocols ...), in the reference's own parameter b: e_k = exp(b x_k), N = sum y_k e_k, D = sum e_k^2 and
// their b-derivatives carry one and two factors of x_k.  One ex2 per sample and pass instead of none, so a
// pass costs about 1.5x the uniform one.  The start is a weighted log-linear fit (weights max(y^2 - c sum
// y^2, 0): samples near the noise floor drop out, no sign tests), two voxels per lane like above.
template <typename T, int E, class YS>
DFIT_HD void mono_general_newton2(const YS& Y, const XTab<T, E>& xt, const SolverOpts<T>& o, pair2<T>& pa,
                                  pair2<T>& pb, pair2<T>& F_out, int (&status)[2], int (&iters)[2]) {
  static_assert(E >= 3, "needs at least three echoes");
  typedef num<T> nm;
  typedef pair2<T> V;
  const unsigned lanes = DFIT_LANES();
  (void)lanes;
  V ysq = p2_mul<T>(Y[0], Y[0]);
#pragma unroll
  for (int e = 1; e < E; ++e) ysq = p2_fma<T>(Y[e], Y[e], ysq);
  // weighted log-linear start: minimise sum w (log2 y^2 - alpha - beta x)^2, b0 = beta ln2 / 2
  V S0 = p2_bcast<T>((T)0), S1 = S0, S2 = S0, T0 = S0, T1 = S0;
  {
    // weights relative to sum y^2 (scale-free: the moments cannot overflow whatever the signal scale)
    const V rysq = p2_make<T>(nm::rcp_(ysq.lo), nm::rcp_(ysq.hi));
    const V cut = p2_bcast<T>((T)-0.004);
    const V tiny = p2_bcast<T>(nm::tiny());
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const V y2 = p2_mul<T>(Y[e], Y[e]);
      const V y2t = p2_add<T>(y2, tiny);
#if defined(__CUDA_ARCH__)
      V l;
      if constexpr (sizeof(T) == 4) {
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l.lo) : "f"(y2t.lo));
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l.hi) : "f"(y2t.hi));
      } else {
        l = p2_make<T>(log2(y2t.lo), log2(y2t.hi));
      }
#else
      const V l = p2_make<T>((T)log2((double)y2t.lo), (T)log2((double)y2t.hi));
#endif
      V w = p2_fma<T>(y2, rysq, cut);
      w = p2_make<T>(nm::max_(w.lo, (T)0), nm::max_(w.hi, (T)0));
      // moments about the mean echo time: the determinant below must not drown in cancellation
      const V xe = p2_bcast<T>(xt.xc[e]), xxe = p2_bcast<T>(xt.xc[e] * xt.xc[e]);
      const V wl = p2_mul<T>(w, l);
      S0 = p2_add<T>(S0, w);
      S1 = p2_fma<T>(xe, w, S1);
      S2 = p2_fma<T>(xxe, w, S2);
      T0 = p2_add<T>(T0, wl);
      T1 = p2_fma<T>(xe, wl, T1);
    }
  }
  const V num_ = p2_fma<T>(S0, T1, p2_mul<T>(p2_mul<T>(S1, T0), p2_bcast<T>((T)-1)));
  const V den_ = p2_fma<T>(S0, S2, p2_mul<T>(p2_mul<T>(S1, S1), p2_bcast<T>((T)-1)));
  const V b0 = p2_mul<T>(p2_mul<T>(num_, p2_make<T>(nm::rcp_(den_.lo), nm::rcp_(den_.hi))), p2_bcast<T>((T)0.34657359027997264));
  NewtonLane<T> A, B;
  // |b x| <= 40 keeps every e_k^2 finite in fp32
  const T blim = (T)(sizeof(T) == 4 ? 40.0 : 300.0) * xt.inv_xmax;
  // den / S0^2 is the weighted variance of the echo times that carry the start: when (nearly) one sample
  // survives the cut it is zero up to rounding and the slope is noise -- such voxels decline.  1e-4 span^2
  // still admits two neighbouring samples of a 16-echo protocol.
  const V dmin = p2_mul<T>(p2_mul<T>(S0, S0), p2_bcast<T>((T)1e-4 * xt.span2));
  A.start(b0.lo, den_.lo > dmin.lo && nm::abs_(b0.lo) < blim && nm::finite(ysq.lo) && ysq.lo > (T)0, (T)0);
  B.start(b0.hi, den_.hi > dmin.hi && nm::abs_(b0.hi) < blim && nm::finite(ysq.hi) && ysq.hi > (T)0, (T)0);
  const V floor2 = p2_mul<T>(ysq, p2_bcast<T>((T)2 * o.floor_rel));
  const V ftol2 = p2_bcast<T>((T)2 * o.ftol);
  const T smax = xt.inv_xmax;  // largest step in b: exp(b x) changes by at most a factor e per pass
  const T bcap = (T)kFirstStepCap * (T)(E - 1) * xt.inv_xmax;  // the same first-step gate per mean echo spacing
  auto pass = [&](auto first_tag, int k) {
    constexpr bool FIRST = decltype(first_tag)::value;
    const V b = p2_make<T>(A.q, B.q);
    V N0 = p2_bcast<T>((T)0), N1 = N0, N2 = N0, D0 = N0, D1 = N0, D2 = N0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const V ex = p2_make<T>(nm::expbx(b.lo, xt.x[e], xt.xs[e]), nm::expbx(b.hi, xt.x[e], xt.xs[e]));
      const V ye = p2_mul<T>(Y[e], ex), ee = p2_mul<T>(ex, ex);
      const V xe = p2_bcast<T>(xt.x[e]), xxe = p2_bcast<T>(xt.xx[e]);
      N0 = p2_add<T>(N0, ye);
      N1 = p2_fma<T>(xe, ye, N1);
      N2 = p2_fma<T>(xxe, ye, N2);
      D0 = p2_add<T>(D0, ee);
      D1 = p2_fma<T>(xe, ee, D1);
      D2 = p2_fma<T>(xxe, ee, D2);
    }
    // N1 = dN/db, N2 = d2N/db2;  dD/db = 2 D1, d2D/db2 = 4 D2
    const V nDb = p2_mul<T>(D1, p2_bcast<T>((T)-2));
    const V rD = p2_make<T>(nm::rcp_(D0.lo), nm::rcp_(D0.hi));
    const V a = p2_mul<T>(N0, rD);
    const V w = p2_fma<T>(a, nDb, N1);
    const V ap = p2_mul<T>(w, rD);
    const V mg = p2_mul<T>(a, p2_add<T>(w, N1));
    const V t2 = p2_fma<T>(N2, p2_bcast<T>((T)-2), p2_mul<T>(a, p2_mul<T>(D2, p2_bcast<T>((T)4))));
    const V h = p2_fma<T>(a, t2, p2_mul<T>(p2_mul<T>(w, ap), p2_bcast<T>((T)-2)));
    const V db = p2_mul<T>(mg, p2_make<T>(nm::rcp_(h.lo), nm::rcp_(h.hi)));
    const V pred2 = p2_mul<T>(mg, db);
    const V Fe = p2_fma<T>(p2_mul<T>(N0, p2_bcast<T>((T)-1)), a, ysq);
    const V tol2 = p2_fma<T>(ftol2, p2_make<T>(nm::max_(Fe.lo, (T)0), nm::max_(Fe.hi, (T)0)), floor2);
    newton_lane_step<FIRST, T>(A, k, h.lo, pred2.lo, tol2.lo, db.lo, a.lo, ap.lo, -smax, smax, bcap);
    newton_lane_step<FIRST, T>(B, k, h.hi, pred2.hi, tol2.hi, db.hi, a.hi, ap.hi, -smax, smax, bcap);
  };
  if (DFIT_ANY(lanes, A.active() || B.active())) pass(FirstPass<true>(), 0);
#pragma unroll 1
  for (int k = 1; k < kMonoFastPasses; ++k) {
    if (!DFIT_ANY(lanes, A.active() || B.active())) break;
    pass(FirstPass<false>(), k);
  }
  iters[0] = A.npass;
  iters[1] = B.npass;
  const bool okA = A.done() && nm::abs_(A.q) < blim && nm::finite(A.af);
  const bool okB = B.done() && nm::abs_(B.q) < blim && nm::finite(B.af);
  const V bf = p2_make<T>(okA ? A.q : (T)0, okB ? B.q : (T)0);
  const V af = p2_make<T>(okA ? A.af : (T)0, okB ? B.af : (T)0);
  // cost at the returned point
  const V naf = p2_mul<T>(af, p2_bcast<T>((T)-1));
  V F = p2_bcast<T>((T)0);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const V ex = p2_make<T>(nm::expbx(bf.lo, xt.x[e], xt.xs[e]), nm::expbx(bf.hi, xt.x[e], xt.xs[e]));
    const V r = p2_fma<T>(naf, ex, Y[e]);
    F = p2_fma<T>(r, r, F);
  }
  pa = af;
  pb = bf;
  F_out = F;
  status[0] = okA ? (F.lo <= o.floor_rel * ysq.lo ? ST_EXACT : ST_CONV_F) : -1;
  status[1] = okB ? (F.hi <= o.floor_rel * ysq.hi ? ST_EXACT : ST_CONV_F) : -1;
}

// ---- straight-line two-pass variants (what the dense kernels run first) ---------------------------------
// Measured on the benchmark volume: 99.65 % of the voxels converge in exactly two Newton passes from the
// data-driven start.  The variants below therefore run exactly two passes for BOTH voxels of a lane with no
// per-lane state machine, no warp votes and no loop -- one basic block the scheduler can interleave freely --
// and simply report `ok = false` for a voxel that would have needed anything else (no admissible start, first
// step beyond the gate, curvature not positive, not converged after the second pass, non-finite result).
// Such voxels are fitted by the one-voxel path (generic Newton loop, then the LM from the caller's p0), which
// the kernels run on whole warps of deferred voxels.  Differences from the loop versions above, all within the
// parity budget (the CPU test-suite runs this very code through the test-only host build, tests/hostsim):
//   * sum (y - mean)^2 comes from the moments, sum y^2 - (sum y)^2 / E (error ~2 eps sum y^2, i.e. an r2 error of
//     (1 - r2) 2 eps sum y^2 / ss_tot), unless that is ill-conditioned (nearly constant signal: below 4 % of
//     sum y^2), where the two-pass form is used.  (The cost F itself cannot be had from the projected cost
//     sum y^2 - N a the same way: its ~2 eps sum y^2 error is divided by ss_tot without the (1 - r2) factor.)
constexpr float kSsTotCond = 0.04f;

template <typename T, int E, class YS>
DFIT_HD pair2<T> ss_total2_fast(const YS& Y, pair2<T> ysq, pair2<T> S) {
  pair2<T> t = p2_fma<T>(p2_mul<T>(S, p2_bcast<T>((T)(-1.0 / E))), S, ysq);
  const bool ill_lo = t.lo < (T)kSsTotCond * ysq.lo, ill_hi = t.hi < (T)kSsTotCond * ysq.hi;
  if (ill_lo || ill_hi) {  // rare: nearly constant signals
    const pair2<T> nmean = p2_mul<T>(S, p2_bcast<T>((T)(-1.0 / E)));
    pair2<T> u = p2_bcast<T>((T)0);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const pair2<T> d = p2_add<T>(Y[e], nmean);
      u = p2_fma<T>(d, d, u);
    }
    // per voxel: a voxel's result must not depend on its neighbour in the pair
    if (ill_lo) t.lo = u.lo;
    if (ill_hi) t.hi = u.hi;
  }
  return t;
}

// The convergence rule of newton_lane_step for the second pass, on packed quantities: kappa^2 pred2 <= tol2
// with kappa = clamp(4 step2 / prev_step2, 1e-3, 1).
template <typename T>
DFIT_HD void second_pass_converged(pair2<T> pred2, pair2<T> tol2, pair2<T> step2, pair2<T> prev2, pair2<T> h,
                                   bool (&conv)[2]) {
  typedef num<T> nm;
  const pair2<T> rr = p2_mul<T>(p2_mul<T>(step2, p2_bcast<T>((T)4)), p2_make<T>(nm::rcp_(prev2.lo), nm::rcp_(prev2.hi)));
  const pair2<T> kap = p2_make<T>(nm::min_(nm::max_(rr.lo, (T)1e-3), (T)1), nm::min_(nm::max_(rr.hi, (T)1e-3), (T)1));
  const pair2<T> lhs = p2_mul<T>(p2_mul<T>(pred2, kap), kap);
  conv[0] = h.lo > (T)0 && lhs.lo <= tol2.lo;
  conv[1] = h.hi > (T)0 && lhs.hi <= tol2.hi;
}

template <typename T, int E, class YS>
DFIT_HD void mono_uniform_fast2(const YS& Y, const XTab<T, E>& xt, const SolverOpts<T>& o, T r2_eps, pair2<T>& pa,
                                pair2<T>& pb, pair2<T>& r2, bool (&ok)[2]) {
  static_assert(E >= 3, "needs at least three echoes");
  typedef num<T> nm;
  typedef pair2<T> V;
  const V one = p2_bcast<T>((T)1);
  // Prony start q0 = sum y_k y_k+1 / sum y_k^2; sum y^2 and sum y fall out of the same recurrence
  V pn = p2_mul<T>(Y[0], Y[1]), pd = p2_mul<T>(Y[0], Y[0]), S = p2_add<T>(Y[0], Y[1]);
#pragma unroll
  for (int k = 1; k + 1 < E; ++k) {
    pn = p2_fma<T>(Y[k], Y[k + 1], pn);
    pd = p2_fma<T>(Y[k], Y[k], pd);
    S = p2_add<T>(S, Y[k + 1]);
  }
  const V ysq = p2_fma<T>(Y[E - 1], Y[E - 1], pd);
  V q;
  if (xt.backward != 0) {  // descending echo times: predicted backwards (see mono_uniform_newton)
    const V pdb = p2_fma<T>(p2_mul<T>(Y[0], p2_bcast<T>((T)-1)), Y[0], ysq);
    q = p2_mul<T>(pdb, p2_make<T>(nm::rcp_(pn.lo), nm::rcp_(pn.hi)));
  } else {
    q = p2_mul<T>(pn, p2_make<T>(nm::rcp_(pd.lo), nm::rcp_(pd.hi)));
  }
  // One evaluation of the projected problem at q: amplitude a, da/dq, Newton step, curvature, twice the
  // Newton decrement and the projected cost.
  V a, ap, dq, h, pred2, Fe;
  auto evaluate = [&]() {
    const V s = p2_mul<T>(q, q);
    // Horner with first and (half) second derivative: N(q) over the samples, D(s) over ones
    V N0 = Y[E - 1], N1 = N0, N2, D0 = one, D1 = one, D2;
    N0 = p2_fma<T>(N0, q, Y[E - 2]);
    D0 = p2_add<T>(s, one);
    N2 = N1;
    D2 = D1;
    N1 = p2_fma<T>(N1, q, N0);
    D1 = p2_add<T>(s, D0);
    N0 = p2_fma<T>(N0, q, Y[E - 3]);
    D0 = p2_fma<T>(D0, s, one);
#pragma unroll
    for (int j = E - 4; j >= 0; --j) {
      N2 = p2_fma<T>(N2, q, N1);
      D2 = p2_fma<T>(D2, s, D1);
      N1 = p2_fma<T>(N1, q, N0);
      D1 = p2_fma<T>(D1, s, D0);
      N0 = p2_fma<T>(N0, q, Y[j]);
      D0 = p2_fma<T>(D0, s, one);
    }
    const V nDq = p2_mul<T>(p2_mul<T>(q, p2_bcast<T>((T)-2)), D1);                     // -dD/dq
    const V Dqq = p2_fma<T>(p2_mul<T>(s, p2_bcast<T>((T)8)), D2, p2_add<T>(D1, D1));  // d2D/dq2
    const V rD = p2_make<T>(nm::rcp_(D0.lo), nm::rcp_(D0.hi));
    a = p2_mul<T>(N0, rD);
    const V w = p2_fma<T>(a, nDq, N1);       // = D da/dq
    ap = p2_mul<T>(w, rD);
    const V mg = p2_mul<T>(a, p2_add<T>(w, N1));  // -dphi/dq
    const V t2 = p2_fma<T>(N2, p2_bcast<T>((T)-4), p2_mul<T>(a, Dqq));
    h = p2_fma<T>(a, t2, p2_mul<T>(p2_mul<T>(w, ap), p2_bcast<T>((T)-2)));  // d2phi/dq2
    dq = p2_mul<T>(mg, p2_make<T>(nm::rcp_(h.lo), nm::rcp_(h.hi)));
    pred2 = p2_mul<T>(mg, dq);
    Fe = p2_fma<T>(p2_mul<T>(N0, p2_bcast<T>((T)-1)), a, ysq);
  };
  // pass 1: first-step gate (see newton_lane_step): inside the convex basin and no further than kFirstStepCap q
  evaluate();
  const V step2a = p2_mul<T>(dq, dq);
  {
    const V lim = p2_mul<T>(p2_mul<T>(q, q), p2_bcast<T>((T)(kFirstStepCap * kFirstStepCap)));
    ok[0] = h.lo > (T)0 && step2a.lo <= lim.lo;
    ok[1] = h.hi > (T)0 && step2a.hi <= lim.hi;
  }
  q = p2_add<T>(q, dq);
  // pass 2: convergence judged like the loop version's second pass
  evaluate();
  {
    const V tol2 = p2_fma<T>(p2_bcast<T>((T)2 * o.ftol), p2_make<T>(nm::max_(Fe.lo, (T)0), nm::max_(Fe.hi, (T)0)),
                             p2_mul<T>(ysq, p2_bcast<T>((T)2 * o.floor_rel)));
    bool conv[2];
    second_pass_converged<T>(pred2, tol2, p2_mul<T>(dq, dq), step2a, h, conv);
    ok[0] = ok[0] && conv[0];
    ok[1] = ok[1] && conv[1];
  }
  const V qf = p2_add<T>(q, dq);
  const V af = p2_fma<T>(ap, dq, a);
  ok[0] = ok[0] && qf.lo > xt.q_lo && qf.lo < xt.q_hi && nm::finite(af.lo);
  ok[1] = ok[1] && qf.hi > xt.q_lo && qf.hi < xt.q_hi && nm::finite(af.hi);
  // cost at the returned point, r_k = y_k - a' q^k (the projected cost sum y^2 - N a cancels to ~2 eps sum y^2,
  // which a nearly constant signal's r2 does not forgive); r2 of fitting.py:1032-1035
  V m = af, F;
  {
    const V r0 = p2_add<T>(Y[0], p2_mul<T>(m, p2_bcast<T>((T)-1)));
    F = p2_mul<T>(r0, r0);
  }
  const V nqf = p2_mul<T>(qf, p2_bcast<T>((T)-1));
  m = p2_mul<T>(m, nqf);  // m = -a' q^k from here on
#pragma unroll
  for (int e = 1; e < E; ++e) {
    const V r = p2_add<T>(Y[e], m);
    F = p2_fma<T>(r, r, F);
    if (e + 1 < E) m = p2_mul<T>(m, qf);
  }
  const V den = p2_add<T>(ss_total2_fast<T, E, YS>(Y, ysq, S), p2_bcast<T>(r2_eps));
  r2 = p2_fma<T>(F, p2_make<T>(-nm::rcp_(den.lo), -nm::rcp_(den.hi)), one);
  // back to the reference's parameters: b = ln(q) / dx, a = a' exp(-b x0)
  const V b = p2_mul<T>(p2_log_pos<T>(qf), p2_bcast<T>(xt.inv_dx));
  pb = b;
  pa = af;
  if (xt.x0 != (T)0) pa = p2_mul<T>(af, p2_make<T>(nm::expbx(-b.lo, xt.x0, xt.x0s), nm::expbx(-b.hi, xt.x0, xt.x0s)));
}

// The same for ARBITRARY echo times (see mono_general_newton2 for the start and the iteration).
template <typename T, int E, class YS>
DFIT_HD void mono_general_fast2(const YS& Y, const XTab<T, E>& xt, const SolverOpts<T>& o, T r2_eps, pair2<T>& pa,
                                pair2<T>& pb, pair2<T>& r2, bool (&ok)[2]) {
  static_assert(E >= 3, "needs at least three echoes");
  typedef num<T> nm;
  typedef pair2<T> V;
  V ysq = p2_mul<T>(Y[0], Y[0]), S = Y[0];
#pragma unroll
  for (int e = 1; e < E; ++e) {
    ysq = p2_fma<T>(Y[e], Y[e], ysq);
    S = p2_add<T>(S, Y[e]);
  }
  // weighted log-linear start: minimise sum w (log2 y^2 - alpha - beta x)^2, b0 = beta ln2 / 2
  V S0 = p2_bcast<T>((T)0), S1 = S0, S2 = S0, T0 = S0, T1 = S0;
  {
    const V rysq = p2_make<T>(nm::rcp_(ysq.lo), nm::rcp_(ysq.hi));
    const V cut = p2_bcast<T>((T)-0.004);
    const V tiny = p2_bcast<T>(nm::tiny());
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const V y2 = p2_mul<T>(Y[e], Y[e]);
      const V y2t = p2_add<T>(y2, tiny);
#if defined(__CUDA_ARCH__)
      V l;
      if constexpr (sizeof(T) == 4) {
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l.lo) : "f"(y2t.lo));
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l.hi) : "f"(y2t.hi));
      } else {
        l = p2_make<T>(log2(y2t.lo), log2(y2t.hi));
      }
#else
      const V l = p2_make<T>((T)log2((double)y2t.lo), (T)log2((double)y2t.hi));
#endif
      V w = p2_fma<T>(y2, rysq, cut);
      w = p2_make<T>(nm::max_(w.lo, (T)0), nm::max_(w.hi, (T)0));
      const V xe = p2_bcast<T>(xt.xc[e]), xxe = p2_bcast<T>(xt.xc[e] * xt.xc[e]);
      const V wl = p2_mul<T>(w, l);
      S0 = p2_add<T>(S0, w);
      S1 = p2_fma<T>(xe, w, S1);
      S2 = p2_fma<T>(xxe, w, S2);
      T0 = p2_add<T>(T0, wl);
      T1 = p2_fma<T>(xe, wl, T1);
    }
  }
  const V num_ = p2_fma<T>(S0, T1, p2_mul<T>(p2_mul<T>(S1, T0), p2_bcast<T>((T)-1)));
  const V den_ = p2_fma<T>(S0, S2, p2_mul<T>(p2_mul<T>(S1, S1), p2_bcast<T>((T)-1)));
  V b = p2_mul<T>(p2_mul<T>(num_, p2_make<T>(nm::rcp_(den_.lo), nm::rcp_(den_.hi))), p2_bcast<T>((T)0.34657359027997264));
  const T blim = (T)(sizeof(T) == 4 ? 40.0 : 300.0) * xt.inv_xmax;  // |b x| <= 40 keeps every e_k^2 finite in fp32
  {
    const V dmin = p2_mul<T>(p2_mul<T>(S0, S0), p2_bcast<T>((T)1e-4 * xt.span2));  // see mono_general_newton2
    ok[0] = den_.lo > dmin.lo && nm::abs_(b.lo) < blim && nm::finite(ysq.lo) && ysq.lo > (T)0;
    ok[1] = den_.hi > dmin.hi && nm::abs_(b.hi) < blim && nm::finite(ysq.hi) && ysq.hi > (T)0;
  }
  V a, ap, db, h, pred2, Fe;
  auto evaluate = [&]() {
    V N0 = p2_bcast<T>((T)0), N1 = N0, N2 = N0, D0 = N0, D1 = N0, D2 = N0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const V ex = p2_make<T>(nm::expbx(b.lo, xt.x[e], xt.xs[e]), nm::expbx(b.hi, xt.x[e], xt.xs[e]));
      const V ye = p2_mul<T>(Y[e], ex), ee = p2_mul<T>(ex, ex);
      const V xe = p2_bcast<T>(xt.x[e]), xxe = p2_bcast<T>(xt.xx[e]);
      N0 = p2_add<T>(N0, ye);
      N1 = p2_fma<T>(xe, ye, N1);
      N2 = p2_fma<T>(xxe, ye, N2);
      D0 = p2_add<T>(D0, ee);
      D1 = p2_fma<T>(xe, ee, D1);
      D2 = p2_fma<T>(xxe, ee, D2);
    }
    const V nDb = p2_mul<T>(D1, p2_bcast<T>((T)-2));
    const V rD = p2_make<T>(nm::rcp_(D0.lo), nm::rcp_(D0.hi));
    a = p2_mul<T>(N0, rD);
    const V w = p2_fma<T>(a, nDb, N1);
    ap = p2_mul<T>(w, rD);
    const V mg = p2_mul<T>(a, p2_add<T>(w, N1));
    const V t2 = p2_fma<T>(N2, p2_bcast<T>((T)-2), p2_mul<T>(a, p2_mul<T>(D2, p2_bcast<T>((T)4))));
    h = p2_fma<T>(a, t2, p2_mul<T>(p2_mul<T>(w, ap), p2_bcast<T>((T)-2)));
    db = p2_mul<T>(mg, p2_make<T>(nm::rcp_(h.lo), nm::rcp_(h.hi)));
    pred2 = p2_mul<T>(mg, db);
    Fe = p2_fma<T>(p2_mul<T>(N0, p2_bcast<T>((T)-1)), a, ysq);
  };
  evaluate();
  const V step2a = p2_mul<T>(db, db);
  {
    // first-step gate per mean echo spacing, and never further than exp(b x) changing by a factor e
    const T cap = nm::min_((T)kFirstStepCap * (T)(E - 1), (T)1) * xt.inv_xmax;
    ok[0] = ok[0] && h.lo > (T)0 && step2a.lo <= cap * cap;
    ok[1] = ok[1] && h.hi > (T)0 && step2a.hi <= cap * cap;
  }
  b = p2_add<T>(b, db);
  evaluate();
  {
    const V tol2 = p2_fma<T>(p2_bcast<T>((T)2 * o.ftol), p2_make<T>(nm::max_(Fe.lo, (T)0), nm::max_(Fe.hi, (T)0)),
                             p2_mul<T>(ysq, p2_bcast<T>((T)2 * o.floor_rel)));
    bool conv[2];
    second_pass_converged<T>(pred2, tol2, p2_mul<T>(db, db), step2a, h, conv);
    ok[0] = ok[0] && conv[0];
    ok[1] = ok[1] && conv[1];
  }
  const V bf = p2_add<T>(b, db);
  const V af = p2_fma<T>(ap, db, a);
  ok[0] = ok[0] && nm::abs_(bf.lo) < blim && nm::finite(af.lo);
  ok[1] = ok[1] && nm::abs_(bf.hi) < blim && nm::finite(af.hi);
  // cost at the returned point (explicit residuals, see mono_uniform_fast2)
  const V naf = p2_mul<T>(af, p2_bcast<T>((T)-1));
  V F = p2_bcast<T>((T)0);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const V ex = p2_make<T>(nm::expbx(bf.lo, xt.x[e], xt.xs[e]), nm::expbx(bf.hi, xt.x[e], xt.xs[e]));
    const V r = p2_fma<T>(naf, ex, Y[e]);
    F = p2_fma<T>(r, r, F);
  }
  const V den = p2_add<T>(ss_total2_fast<T, E, YS>(Y, ysq, S), p2_bcast<T>(r2_eps));
  r2 = p2_fma<T>(F, p2_make<T>(-nm::rcp_(den.lo), -nm::rcp_(den.hi)), p2_bcast<T>((T)1));
  pa = af;
  pb = bf;
}

// sum (y - mean)^2 of two voxels at once
template <typename T, int E, class YS>
DFIT_HD pair2<T> ss_total2(const YS& Y) {
  pair2<T> s = Y[0];
#pragma unroll
  for (int e = 1; e < E; ++e) s = p2_add<T>(s, Y[e]);
  const pair2<T> nmean = p2_mul<T>(s, p2_bcast<T>((T)(-1.0 / E)));
  pair2<T> t = p2_bcast<T>((T)0);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const pair2<T> d = p2_add<T>(Y[e], nmean);
    t = p2_fma<T>(d, d, t);
  }
  return t;
}

// sum (y - mean)^2 for the r2 of fitting.py:1032-1035, two samples per instruction
template <typename T, int E>
DFIT_HD T ss_total(const T (&y)[E]) {
  pair2<T> s2 = p2_bcast<T>((T)0);
#pragma unroll
  for (int e = 0; e + 1 < E; e += 2) s2 = p2_add<T>(s2, p2_make<T>(y[e], y[e + 1]));
  T s = s2.lo + s2.hi;
  if constexpr (E & 1) s += y[E - 1];
  const T mean = s / (T)E;
  const pair2<T> nm2 = p2_bcast<T>(-mean);
  pair2<T> t2 = p2_bcast<T>((T)0);
#pragma unroll
  for (int e = 0; e + 1 < E; e += 2) {
    const pair2<T> d = p2_add<T>(p2_make<T>(y[e], y[e + 1]), nm2);
    t2 = p2_fma<T>(d, d, t2);
  }
  T t = t2.lo + t2.hi;
  if constexpr (E & 1) t = num<T>::fma_(y[E - 1] - mean, y[E - 1] - mean, t);
  return t;
}


// Loop-form fast-path attempt for two voxels on ARBITRARY echo times (the one-voxel path below runs its voxel through
// it; uniformly spaced echoes have their own one-voxel solver, mono_uniform_newton): status[i] = -1 where voxel i has
// to take the general path.
template <class M, typename T, int EMAX, class YS>
DFIT_HD void fit_voxel_fast2(const YS& Y, const XTab<T, EMAX>& xt, const VoxelOpts<T>& vo, pair2<T>& pa,
                             pair2<T>& pb, pair2<T>& r2, int (&status)[2], int (&iters)[2]) {
  static_assert(M::MONO && EMAX >= 3, "mono-exponential model only");
  pair2<T> F;
  mono_general_newton2<T, EMAX, YS>(Y, xt, vo.s, pa, pb, F, status, iters);
  const pair2<T> den = p2_add<T>(ss_total2<T, EMAX, YS>(Y), p2_bcast<T>(vo.r2_eps));
  const pair2<T> nr = p2_make<T>(-num<T>::rcp_(den.lo), -num<T>::rcp_(den.hi));
  r2 = p2_fma<T>(F, nr, p2_bcast<T>((T)1));  // fitting.py:1032-1035
}

// Straight-line fast-path attempt for two voxels (what the dense kernels run): ok[i] = false where voxel i has
// to take the one-voxel path (fit_voxel_fast, then fit_voxel).  Passes spent by a voxel that is ok:
constexpr int kFast2Passes = 2;
template <class M, typename T, int EMAX, class YS>
DFIT_HD void fit_voxel_fast2s(const YS& Y, const XTab<T, EMAX>& xt, const VoxelOpts<T>& vo, pair2<T>& pa, pair2<T>& pb,
                              pair2<T>& r2, bool (&ok)[2]) {
  static_assert(M::MONO && EMAX >= 3, "mono-exponential model only");
  if (xt.uniform != 0) mono_uniform_fast2<T, EMAX, YS>(Y, xt, vo.s, vo.r2_eps, pa, pb, r2, ok);
  else mono_general_fast2<T, EMAX, YS>(Y, xt, vo.s, vo.r2_eps, pa, pb, r2, ok);
}

// Fast-path attempt for one voxel (mono-exponential model, no y_bounds).  Returns a
// successful Status with p / r2 / iters filled in, or -1: the caller then loads the initial guess and runs
// fit_voxel.  The path declines on anything unusual -- zero, non-finite or non-decaying-looking voxels
// included -- so the skip / failure rules of fitting.py:1065-1073 stay with the general path.
// Warp-collective on the device: call it with the lanes of a warp converged.
template <class M, typename T, int EMAX, bool EXACT>
DFIT_HD int fit_voxel_fast(const T (&y)[EMAX], const XTab<T, EMAX>& xt, const VoxelOpts<T>& vo, T (&p)[M::P], T& r2,
                           int& iters) {
  if constexpr (M::MONO && EXACT && EMAX >= 3) {
    if (vo.fast != 0 && vo.has_bounds == 0) {
      T F;
      int st;
      if (xt.uniform != 0) {
        st = mono_uniform_newton<T, EMAX>(y, xt, vo.s, p, F, iters);
      } else {  // the general solver is written for two voxels per lane: run it on the voxel twice
        pair2<T> Y[EMAX], pa, pb, r2p;
#pragma unroll
        for (int e = 0; e < EMAX; ++e) Y[e] = p2_bcast<T>(y[e]);
        int st2[2], it2[2];
        fit_voxel_fast2<M, T, EMAX, pair2<T>[EMAX]>(Y, xt, vo, pa, pb, r2p, st2, it2);
        iters = it2[0];
        p[0] = pa.lo;
        p[1] = pb.lo;
        r2 = r2p.lo;
        return st2[0];
      }
      if (st > 0) r2 = (T)1 - F * num<T>::rcp_(ss_total<T, EMAX>(y) + vo.r2_eps);  // fitting.py:1032-1035
      return st;
    }
  }
  return -1;
}

}  // namespace dfit
