/* CPU ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C restatement of the arithmetic that ad12/DOSMA's per-voxel fit delegates to a
 * third-party dependency that is not vendored in /root/reference:
 *
 *   dosma/core/fitting.py:1030   sop.curve_fit(func, x, y, p0=p0, ftol=ftol, maxfev=100)
 *     -> scipy.optimize.leastsq -> MINPACK `lmdif` (SciPy 1.18.1 in this image; unpinned in the
 *        reference: requirements.txt:12, setup.py:108)
 *
 * What is restated here is MINPACK's published algorithm (More', Garbow, Hillstrom, "User Guide
 * for MINPACK-1", ANL-80-74, 1980): enorm, qrfac (Householder QR with column pivoting), qrsolv,
 * lmpar (More' 1978 trust-region LM parameter), fdjac2 (forward differences) and the lmdif
 * driver, with the call-site constants SciPy's `leastsq` passes for DOSMA's call:
 *   ftol = 1e-5 (fitting.py:762), xtol = 1.49012e-8, gtol = 0, maxfev = 100 (fitting.py:761),
 *   epsfcn = DBL_EPSILON, factor = 100, mode = 1 (automatic column scaling).
 * ier in {1,2,3,4} is success; anything else is the RuntimeError the reference turns into
 * (nan,)*P, r2 = 0 (fitting.py:1069-1073).  The all-zero skip (:1065-1067) and
 * r2 = 1 - SS_res/(SS_tot + eps) (:1032-1035) are restated in `dosma_fit_voxels`.
 *
 * Parity status: PINNED against real SciPy outputs (popt, nfev, ier) by tests/test_oracle.py,
 * and through tests/golden/ against the real reference code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the shared object built from this file (oracle/Makefile -> oracle/liboracle.so).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXP 8   /* parameters  */
#define MAXE 64  /* echoes      */

typedef void (*model_fn)(const double *x, int m, const double *p, double *f);

static void f_monoexp(const double *x, int m, const double *p, double *f) {
  for (int i = 0; i < m; ++i) f[i] = p[0] * exp(p[1] * x[i]);
}
static void f_biexp(const double *x, int m, const double *p, double *f) {
  for (int i = 0; i < m; ++i) f[i] = p[0] * exp(p[1] * x[i]) + p[2] * exp(p[3] * x[i]);
}
static void f_linear(const double *x, int m, const double *p, double *f) {
  for (int i = 0; i < m; ++i) f[i] = p[0] * x[i];
}

/* Euclidean norm with MINPACK's three-accumulator over/underflow guard. */
static double enorm(int n, const double *v) {
  const double rdwarf = 3.834e-20, rgiant = 1.304e19;
  double s_big = 0, s_mid = 0, s_small = 0, big_max = 0, small_max = 0;
  const double agiant = rgiant / (double)n;
  for (int i = 0; i < n; ++i) {
    double a = fabs(v[i]);
    if (a > rdwarf && a < agiant) {
      s_mid += a * a;
    } else if (a <= rdwarf) {
      if (a > small_max) {
        double q = small_max / a;
        s_small = 1.0 + s_small * q * q;
        small_max = a;
      } else if (a != 0) {
        double q = a / small_max;
        s_small += q * q;
      }
    } else {
      if (a > big_max) {
        double q = big_max / a;
        s_big = 1.0 + s_big * q * q;
        big_max = a;
      } else {
        double q = a / big_max;
        s_big += q * q;
      }
    }
  }
  if (s_big != 0) return big_max * sqrt(s_big + (s_mid / big_max) / big_max);
  if (s_mid != 0) {
    if (s_mid >= small_max) return sqrt(s_mid * (1.0 + (small_max / s_mid) * (small_max * s_small)));
    return sqrt(small_max * ((s_mid / small_max) + (small_max * s_small)));
  }
  return small_max * sqrt(s_small);
}

/* Householder QR with column pivoting of the m x n column-major matrix a (leading dim m).
 * On return the strict upper triangle of a holds R, the lower trapezoid the Householder vectors,
 * rdiag the diagonal of R, acnorm the input column norms, ipvt the permutation. */
static void qrfac(int m, int n, double *a, int *ipvt, double *rdiag, double *acnorm, double *wa) {
  for (int j = 0; j < n; ++j) {
    acnorm[j] = enorm(m, a + (size_t)j * m);
    rdiag[j] = acnorm[j];
    wa[j] = rdiag[j];
    ipvt[j] = j;
  }
  int minmn = m < n ? m : n;
  for (int j = 0; j < minmn; ++j) {
    int kmax = j;
    for (int k = j; k < n; ++k)
      if (rdiag[k] > rdiag[kmax]) kmax = k;
    if (kmax != j) {
      for (int i = 0; i < m; ++i) {
        double t = a[i + (size_t)j * m];
        a[i + (size_t)j * m] = a[i + (size_t)kmax * m];
        a[i + (size_t)kmax * m] = t;
      }
      rdiag[kmax] = rdiag[j];
      wa[kmax] = wa[j];
      int t = ipvt[j];
      ipvt[j] = ipvt[kmax];
      ipvt[kmax] = t;
    }
    double *cj = a + (size_t)j * m;
    double ajnorm = enorm(m - j, cj + j);
    if (ajnorm != 0) {
      if (cj[j] < 0) ajnorm = -ajnorm;
      for (int i = j; i < m; ++i) cj[i] /= ajnorm;
      cj[j] += 1.0;
      for (int k = j + 1; k < n; ++k) {
        double *ck = a + (size_t)k * m;
        double sum = 0;
        for (int i = j; i < m; ++i) sum += cj[i] * ck[i];
        double t = sum / cj[j];
        for (int i = j; i < m; ++i) ck[i] -= t * cj[i];
        if (rdiag[k] != 0) {
          double q = ck[j] / rdiag[k];
          double d = 1.0 - q * q;
          rdiag[k] *= sqrt(d > 0 ? d : 0);
          double w = rdiag[k] / wa[k];
          if (0.05 * (w * w) <= DBL_EPSILON) {
            rdiag[k] = enorm(m - j - 1, ck + j + 1);
            wa[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
}

/* Solve  [R P^T; D P^T] x ~ [Q^T b; 0]  by Givens elimination of the diagonal D.
 * r is n x n column-major (leading dim ldr) holding R in its upper triangle; the strict lower
 * triangle is used as workspace for S^T.  sdiag receives the diagonal of S. */
static void qrsolv(int n, double *r, int ldr, const int *ipvt, const double *diag, const double *qtb,
                   double *x, double *sdiag, double *wa) {
  for (int j = 0; j < n; ++j) {
    for (int i = j; i < n; ++i) r[i + (size_t)j * ldr] = r[j + (size_t)i * ldr];
    x[j] = r[j + (size_t)j * ldr];
    wa[j] = qtb[j];
  }
  for (int j = 0; j < n; ++j) {
    int l = ipvt[j];
    if (diag[l] != 0) {
      for (int k = j; k < n; ++k) sdiag[k] = 0;
      sdiag[j] = diag[l];
      double qtbpj = 0;
      for (int k = j; k < n; ++k) {
        if (sdiag[k] == 0) continue;
        double rkk = r[k + (size_t)k * ldr], c, s;
        if (fabs(rkk) < fabs(sdiag[k])) {
          double cot = rkk / sdiag[k];
          s = 0.5 / sqrt(0.25 + 0.25 * cot * cot);
          c = s * cot;
        } else {
          double tn = sdiag[k] / rkk;
          c = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
          s = c * tn;
        }
        r[k + (size_t)k * ldr] = c * rkk + s * sdiag[k];
        double t = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = t;
        for (int i = k + 1; i < n; ++i) {
          double rik = r[i + (size_t)k * ldr];
          t = c * rik + s * sdiag[i];
          sdiag[i] = -s * rik + c * sdiag[i];
          r[i + (size_t)k * ldr] = t;
        }
      }
    }
    sdiag[j] = r[j + (size_t)j * ldr];
    r[j + (size_t)j * ldr] = x[j];
  }
  int nsing = n;
  for (int j = 0; j < n; ++j) {
    if (sdiag[j] == 0 && nsing == n) nsing = j;
    if (nsing < n) wa[j] = 0;
  }
  for (int j = nsing - 1; j >= 0; --j) {
    double sum = 0;
    for (int i = j + 1; i < nsing; ++i) sum += r[i + (size_t)j * ldr] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
}

/* More' (1978): find par >= 0 such that ||D x(par)|| is within 10 % of delta (or par = 0 when the
 * Gauss-Newton step already is). */
static void lmpar(int n, double *r, int ldr, const int *ipvt, const double *diag, const double *qtb,
                  double delta, double *par, double *x, double *sdiag, double *wa1, double *wa2) {
  const double dwarf = DBL_MIN;
  int nsing = n;
  for (int j = 0; j < n; ++j) {
    wa1[j] = qtb[j];
    if (r[j + (size_t)j * ldr] == 0 && nsing == n) nsing = j;
    if (nsing < n) wa1[j] = 0;
  }
  for (int j = nsing - 1; j >= 0; --j) {
    wa1[j] /= r[j + (size_t)j * ldr];
    double t = wa1[j];
    for (int i = 0; i < j; ++i) wa1[i] -= r[i + (size_t)j * ldr] * t;
  }
  for (int j = 0; j < n; ++j) x[ipvt[j]] = wa1[j];

  int iter = 0;
  for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = enorm(n, wa2);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    *par = 0;
    return;
  }
  double parl = 0;
  if (nsing >= n) {
    for (int j = 0; j < n; ++j) {
      int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < n; ++j) {
      double sum = 0;
      for (int i = 0; i < j; ++i) sum += r[i + (size_t)j * ldr] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j + (size_t)j * ldr];
    }
    double t = enorm(n, wa1);
    parl = ((fp / delta) / t) / t;
  }
  for (int j = 0; j < n; ++j) {
    double sum = 0;
    for (int i = 0; i <= j; ++i) sum += r[i + (size_t)j * ldr] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  double gnorm = enorm(n, wa1);
  double paru = gnorm / delta;
  if (paru == 0) paru = dwarf / (delta < 0.1 ? delta : 0.1);
  if (*par < parl) *par = parl;
  if (*par > paru) *par = paru;
  if (*par == 0) *par = gnorm / dxnorm;

  for (;;) {
    ++iter;
    if (*par == 0) *par = dwarf > 0.001 * paru ? dwarf : 0.001 * paru;
    double t = sqrt(*par);
    for (int j = 0; j < n; ++j) wa1[j] = t * diag[j];
    qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = enorm(n, wa2);
    t = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0 && fp <= t && t < 0) || iter == 10) break;
    for (int j = 0; j < n; ++j) {
      int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < n; ++j) {
      wa1[j] /= sdiag[j];
      double w = wa1[j];
      for (int i = j + 1; i < n; ++i) wa1[i] -= r[i + (size_t)j * ldr] * w;
    }
    t = enorm(n, wa1);
    double parc = ((fp / delta) / t) / t;
    if (fp > 0 && *par > parl) parl = *par;
    if (fp < 0 && *par < paru) paru = *par;
    *par = parl > *par + parc ? parl : *par + parc;
  }
}

typedef struct {
  model_fn f;
  const double *x;
  const double *y;
  int m;
} problem;

static void residual(const problem *pb, const double *p, double *r) {
  pb->f(pb->x, pb->m, p, r);
  for (int i = 0; i < pb->m; ++i) r[i] -= pb->y[i]; /* scipy: func(x, *p) - ydata */
}

/* The lmdif driver.  Returns MINPACK's info (ier); p is updated in place; *nfev_out counts
 * model evaluations including the n per forward-difference Jacobian. */
static int lmdif(const problem *pb, int n, double *p, double ftol, double xtol, double gtol, int maxfev,
                 double epsfcn, double factor, int *nfev_out) {
  const int m = pb->m;
  double fvec[MAXE], fjac[MAXE * MAXP], diag[MAXP], qtf[MAXP];
  double wa1[MAXP], wa2[MAXP], wa3[MAXP], wa4[MAXE];
  int ipvt[MAXP];
  int info = 0, nfev = 0, iter = 1;
  double par = 0, delta = 0, xnorm = 0, gnorm = 0, fnorm, ratio;

  if (n <= 0 || m < n || ftol < 0 || xtol < 0 || gtol < 0 || maxfev <= 0 || factor <= 0) {
    *nfev_out = 0;
    return 0;
  }
  residual(pb, p, fvec);
  nfev = 1;
  fnorm = enorm(m, fvec);

  for (;;) {
    { /* forward-difference Jacobian */
      double h0 = sqrt(epsfcn > DBL_EPSILON ? epsfcn : DBL_EPSILON);
      for (int j = 0; j < n; ++j) {
        double t = p[j], h = h0 * fabs(t);
        if (h == 0) h = h0;
        p[j] = t + h;
        residual(pb, p, wa4);
        p[j] = t;
        for (int i = 0; i < m; ++i) fjac[i + (size_t)j * m] = (wa4[i] - fvec[i]) / h;
      }
      nfev += n;
    }
    qrfac(m, n, fjac, ipvt, wa1, wa2, wa3);
    if (iter == 1) {
      for (int j = 0; j < n; ++j) {
        diag[j] = wa2[j];
        if (wa2[j] == 0) diag[j] = 1;
      }
      for (int j = 0; j < n; ++j) wa3[j] = diag[j] * p[j];
      xnorm = enorm(n, wa3);
      delta = factor * xnorm;
      if (delta == 0) delta = factor;
    }
    /* first n components of Q^T fvec */
    for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
    for (int j = 0; j < n; ++j) {
      double *cj = fjac + (size_t)j * m;
      if (cj[j] != 0) {
        double sum = 0;
        for (int i = j; i < m; ++i) sum += cj[i] * wa4[i];
        double t = -sum / cj[j];
        for (int i = j; i < m; ++i) wa4[i] += cj[i] * t;
      }
      cj[j] = wa1[j];
      qtf[j] = wa4[j];
    }
    gnorm = 0;
    if (fnorm != 0) {
      for (int j = 0; j < n; ++j) {
        int l = ipvt[j];
        if (wa2[l] == 0) continue;
        double sum = 0;
        for (int i = 0; i <= j; ++i) sum += fjac[i + (size_t)j * m] * (qtf[i] / fnorm);
        double g = fabs(sum / wa2[l]);
        if (g > gnorm) gnorm = g;
      }
    }
    if (gnorm <= gtol) {
      info = 4;
      break;
    }
    for (int j = 0; j < n; ++j)
      if (wa2[j] > diag[j]) diag[j] = wa2[j];

    do {
      lmpar(n, fjac, m, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa4);
      for (int j = 0; j < n; ++j) {
        wa1[j] = -wa1[j];
        wa2[j] = p[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
      }
      double pnorm = enorm(n, wa3);
      if (iter == 1 && pnorm < delta) delta = pnorm;
      residual(pb, wa2, wa4);
      ++nfev;
      double fnorm1 = enorm(m, wa4);
      double actred = -1;
      if (0.1 * fnorm1 < fnorm) {
        double q = fnorm1 / fnorm;
        actred = 1.0 - q * q;
      }
      for (int j = 0; j < n; ++j) {
        wa3[j] = 0;
        double t = wa1[ipvt[j]];
        for (int i = 0; i <= j; ++i) wa3[i] += fjac[i + (size_t)j * m] * t;
      }
      double t1 = enorm(n, wa3) / fnorm;
      double t2 = (sqrt(par) * pnorm) / fnorm;
      double prered = t1 * t1 + t2 * t2 / 0.5;
      double dirder = -(t1 * t1 + t2 * t2);
      ratio = prered != 0 ? actred / prered : 0;
      if (ratio <= 0.25) {
        double t = 0.5;
        if (actred < 0) t = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || t < 0.1) t = 0.1;
        double pn = pnorm / 0.1;
        delta = t * (delta < pn ? delta : pn);
        par /= t;
      } else if (par == 0 || ratio >= 0.75) {
        delta = pnorm / 0.5;
        par *= 0.5;
      }
      if (ratio >= 1e-4) {
        for (int j = 0; j < n; ++j) {
          p[j] = wa2[j];
          wa2[j] = diag[j] * p[j];
        }
        for (int i = 0; i < m; ++i) fvec[i] = wa4[i];
        xnorm = enorm(n, wa2);
        fnorm = fnorm1;
        ++iter;
      }
      int small = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1;
      if (small) info = 1;
      if (delta <= xtol * xnorm) info = 2;
      if (small && info == 2) info = 3;
      if (info != 0) goto done;
      if (nfev >= maxfev) info = 5;
      if (fabs(actred) <= DBL_EPSILON && prered <= DBL_EPSILON && 0.5 * ratio <= 1) info = 6;
      if (delta <= DBL_EPSILON * xnorm) info = 7;
      if (gnorm <= DBL_EPSILON) info = 8;
      if (info != 0) goto done;
    } while (ratio < 1e-4);
  }
done:
  *nfev_out = nfev;
  return info;
}

/* ---- DOSMA per-voxel wrapper -------------------------------------------------------------
 * model: 0 monoexponential, 1 biexponential, 2 linear (a*x).
 * y: planar (E, N) float64, row stride ldy.  p0: (n_p0, P) with n_p0 in {1, N}.
 * Outputs popt (N, P), r2 (N), info (N, 2) = (nfev, ier) [may be NULL].
 * Voxels [v0, v1) are processed so callers can thread over ranges. */
int dosma_fit_voxels(int model, int m, const double *x, const double *y, int64_t ldy, int64_t v0, int64_t v1,
                     const double *p0, int64_t n_p0, double ftol, int maxfev, double r2_eps, double *popt,
                     double *r2, int32_t *info) {
  model_fn f;
  int n;
  switch (model) {
    case 0: f = f_monoexp; n = 2; break;
    case 1: f = f_biexp; n = 4; break;
    case 2: f = f_linear; n = 1; break;
    default: return -1;
  }
  if (m > MAXE || m < n) return -2;
  for (int64_t v = v0; v < v1; ++v) {
    double yv[MAXE], p[MAXP], fit[MAXE];
    int all_zero = 1;
    for (int i = 0; i < m; ++i) {
      yv[i] = y[(size_t)i * ldy + v];
      if (yv[i] != 0) all_zero = 0;
    }
    int nfev = 0, ier = 0;
    if (!all_zero) {
      const double *pv = p0 + (n_p0 > 1 ? (size_t)v * n : 0);
      for (int j = 0; j < n; ++j) p[j] = pv[j];
      problem pb = {f, x, yv, m};
      ier = lmdif(&pb, n, p, ftol, 1.49012e-8, 0.0, maxfev, DBL_EPSILON, 100.0, &nfev);
    }
    if (ier >= 1 && ier <= 4) {
      f(x, m, p, fit);
      double mean = 0, ss_res = 0, ss_tot = 0;
      for (int i = 0; i < m; ++i) mean += yv[i];
      mean /= m;
      for (int i = 0; i < m; ++i) {
        double d = yv[i] - fit[i], c = yv[i] - mean;
        ss_res += d * d;
        ss_tot += c * c;
      }
      for (int j = 0; j < n; ++j) popt[(size_t)v * n + j] = p[j];
      r2[v] = 1.0 - ss_res / (ss_tot + r2_eps);
    } else {
      for (int j = 0; j < n; ++j) popt[(size_t)v * n + j] = NAN;
      r2[v] = 0.0;
    }
    if (info) {
      info[2 * v] = nfev;
      info[2 * v + 1] = ier;
    }
  }
  return 0;
}
