"""Helpers to read the fixtures under tests/golden/ (written by tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "meta"}
    d["meta"] = json.loads(str(z["meta"]))
    return d


def p0_of(case):
    if "p0_arr" in case:
        return case["p0_arr"]
    p0 = case["meta"]["p0"]
    return tuple(p0) if isinstance(p0, list) else p0


def same_nan(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b))


# ---- bi-exponential fp32 parity (BASELINE config 4's arithmetic) --------------------------------------------
# The 4-parameter fit is ill-conditioned: MINPACK at the reference's ftol = 1e-5 stops up to ~1e-3 (relative) short
# of the minimiser on noisy data, and fp32 arithmetic resolves the parameters to ~1e-4 even on noise-free data.
# What both solvers agree on tightly is the VALUE of the minimum, i.e. r2.  Tolerances, measured on the two
# fixtures (engine fp32 vs the reference's float64 SciPy output) and stated here once for the CPU (host build of
# the solver) and GPU tests:
BIEXP_F32_TOL = {
    # fixture: (median, p90, p99 of the per-voxel max relative parameter error, r2 atol, allowed NaN-set symmetric difference)
    "curvefit_biexp16_clean_f32": (5e-5, 3e-4, 1e-3, 1e-6, 0.0),
    "curvefit_biexp16_snr100_f32": (1.5e-4, 2e-3, 2e-2, 1e-6, 0.02),
}


def check_biexp_f32(case_name, popt, r2):
    """Assert the stated fp32 tolerances of `popt`, `r2` against the golden reference output of a bi-exponential fixture."""
    c = load(case_name)
    med, p90, p99, r2_atol, nan_frac = BIEXP_F32_TOL[case_name]
    ref_nan, nan = np.isnan(c["popt"][:, 0]), np.isnan(popt[:, 0])
    assert np.isnan(popt[nan]).all() and (r2[nan] == 0).all()  # fitting.py:1069-1073
    assert (ref_nan ^ nan).mean() <= nan_frac, ((ref_nan ^ nan).sum(), ref_nan.sum(), nan.sum())
    ok = ~ref_nan & ~nan
    rel = (np.abs(popt[ok] - c["popt"][ok]) / np.abs(c["popt"][ok])).max(axis=1)
    q = np.percentile(rel, [50, 90, 99])
    assert q[0] < med and q[1] < p90 and q[2] < p99, q
    assert np.abs(r2[ok] - c["r2"][ok]).max() < r2_atol, np.abs(r2[ok] - c["r2"][ok]).max()
    return q
