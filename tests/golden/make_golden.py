"""Generate the golden fixtures in this directory from the REAL reference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture is an .npz holding the seeded synthetic inputs AND the outputs of
`/root/reference/dosma/core/fitting.py` (loaded verbatim through `ref_loader.py`; SciPy 1.18.1,
numpy 2.3.5) on those inputs.  The reference's own test-suite has no golden vectors for this
path (all checks are unseeded analytic-truth checks), so these files are
what pins the oracle -- and through it the CUDA path -- to the reference's actual outputs.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference_fitting  # noqa: E402

F, MV = load_reference_fitting()
warnings.filterwarnings("ignore")


def save(name, meta, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name:34s} {os.path.getsize(path) / 1024:8.1f} KiB")


def mono_data(rng, x, n, snr=None, dtype=np.float32, a_rng=(500, 1500), t_rng=(10, 80)):
    a = rng.uniform(*a_rng, n)
    t = rng.uniform(*t_rng, n)
    y = a * np.exp(-np.asarray(x)[:, None] / t)
    if snr:
        y = y + rng.normal(0, 1000.0 / snr, y.shape)
    return y.astype(dtype), a, t


# ---------------------------------------------------------------- array level: curve_fit
def case_curve_fit(name, func, x, y, p0, seed, **kw):
    popt, r2 = F.curve_fit(func, x, y.astype(np.float64) if kw.pop("upcast", False) else y, p0=p0, **kw)
    meta = {"kind": "curve_fit", "func": func.__name__, "seed": seed,
            "p0": p0 if not isinstance(p0, np.ndarray) else "array", "kwargs": {k: v for k, v in kw.items()}}
    extra = {"p0_arr": p0} if isinstance(p0, np.ndarray) else {}
    save(name, meta, x=np.asarray(x, dtype=np.float64), y=y, popt=popt, r2=r2, **extra)


rng = np.random.default_rng(0)
x4 = [10.0, 20.0, 40.0, 80.0]
x8 = [10.0 * i for i in range(1, 9)]
x7 = [0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]  # MAPSS echo/spin-lock times, tests/.../test_mapss.py:43
x16 = [5.0 * i for i in range(1, 17)]

# C1-like parity volume (BASELINE config 1 distribution), fixed p0 = (1, -1/30) (tc0 = 30 default)
y, _, _ = mono_data(rng, x4, 4096)
case_curve_fit("curvefit_mono4_clean_f32", F.monoexponential, x4, y, (1.0, -1 / 30), 0)
y, _, _ = mono_data(rng, x8, 4096)
case_curve_fit("curvefit_mono8_clean_f32", F.monoexponential, x8, y, (1.0, -1 / 30), 0)
y, _, _ = mono_data(rng, x8, 4096, snr=100)
case_curve_fit("curvefit_mono8_snr100_f32", F.monoexponential, x8, y, (1.0, -1 / 30), 0)
y, _, _ = mono_data(rng, x8, 2048, snr=30)
case_curve_fit("curvefit_mono8_snr30_f32", F.monoexponential, x8, y, (1.0, -1 / 30), 0)
y, _, _ = mono_data(rng, x7, 2048, snr=100, t_rng=(20, 120))
case_curve_fit("curvefit_mono7_t1rho_snr100_f32", F.monoexponential, x7, y, (1.0, -1 / 30), 0)

# the reference tests' own distribution: growing exponentials, default p0 (ones) and (1, -1/30)
xt = [0.5, 1.0, 2.0, 4.0]
b = rng.random(2000) + 0.1
yt = 1.0 * np.exp(b * np.asarray(xt)[:, None])
case_curve_fit("curvefit_mono4_growing_f64_p0ones", F.monoexponential, xt, yt, None, 0)
case_curve_fit("curvefit_mono4_growing_f64_tc30", F.monoexponential, xt, yt, (1.0, -1 / 30), 0)
# test_fitting.py:71-84 distribution: x = 1..4, a, b ~ U(0,1)
xs = [1.0, 2.0, 3.0, 4.0]
ys = np.stack([rng.random() * np.exp(rng.random() * np.asarray(xs)) for _ in range(1000)], axis=-1)
case_curve_fit("curvefit_mono4_unit_f64_p0ones", F.monoexponential, xs, ys, None, 0)
case_curve_fit("curvefit_mono4_unit_f64_p0dict", F.monoexponential, xs, ys, {"b": 0.5}, 0)

# per-voxel p0 (N, P) and mixed [array, scalar]
y, a, t = mono_data(rng, x8, 1024, snr=100)
p0 = np.stack([a * rng.uniform(0.7, 1.3, a.size), -1 / (t * rng.uniform(0.7, 1.3, t.size))], axis=-1)
case_curve_fit("curvefit_mono8_snr100_p0voxel", F.monoexponential, x8, y, p0, 0)

# degenerate voxels: all-zero rows, single zero samples, negative samples, constant rows
y, _, _ = mono_data(rng, x8, 512, snr=30)
y[:, :16] = 0
y[3, 16:32] = 0
y[5:, 32:48] = -np.abs(y[5:, 32:48])
y[:, 48:64] = 100.0
case_curve_fit("curvefit_mono8_degenerate_f32", F.monoexponential, x8, y, (1.0, -1 / 30), 0)
# y_bounds skip rule (fitting.py:1065)
y, _, _ = mono_data(rng, x8, 512, snr=100)
case_curve_fit("curvefit_mono8_ybounds", F.monoexponential, x8, y, (1.0, -1 / 30), 0, y_bounds=(0, 1400))

# low SNR: exercises the maxfev=100 failure path (NaN, r2 = 0)
y, _, _ = mono_data(rng, x8, 2048, snr=5)
case_curve_fit("curvefit_mono8_snr5_f32", F.monoexponential, x8, y, (1.0, -1 / 30), 0)

# bi-exponential (BASELINE config 4 distribution)
n = 1024
A = rng.uniform(500, 1500, n)
fs = rng.uniform(0.3, 0.7, n)
ts, tl = rng.uniform(8, 20, n), rng.uniform(50, 100, n)
yb = A * fs * np.exp(-np.asarray(x16)[:, None] / ts) + A * (1 - fs) * np.exp(-np.asarray(x16)[:, None] / tl)
p0b = (500.0, -1 / 10, 500.0, -1 / 60)
case_curve_fit("curvefit_biexp16_clean_f32", F.biexponential, x16, yb.astype(np.float32), p0b, 0)
ybn = (yb + rng.normal(0, 10.0, yb.shape)).astype(np.float32)
case_curve_fit("curvefit_biexp16_snr100_f32", F.biexponential, x16, ybn, p0b, 0)


# linear 1-parameter model of the reference tests (test_fitting.py:52-53)
def _linear(x, a):
    return a * x


al = rng.random(512) + 0.1
yl = al * np.asarray(xt)[:, None]
popt, r2 = F.curve_fit(_linear, xt, yl)
save("curvefit_linear4_f64", {"kind": "curve_fit", "func": "linear", "seed": 0, "p0": None, "kwargs": {}},
     x=np.asarray(xt), y=yl, popt=popt, r2=r2)


# ---------------------------------------------------------------- class level
def vols(y, shape, affine=None):
    affine = np.eye(4) if affine is None else affine
    return [MV(y[e].reshape(shape), affine) for e in range(y.shape[0])]


shape = (16, 16, 8)
n = int(np.prod(shape))
aff = np.array([[0, 0, 1.5, -61.7], [-0.3125, 0, 0, 50.9], [0, -0.3125, 0, 88.6], [0, 0, 0, 1.0]])

for tag, snr, dtype in (("clean", None, np.float32), ("snr100", 100, np.float32), ("snr30_i16", 30, np.int16)):
    y, _, _ = mono_data(rng, x8, n, snr=snr, dtype=np.float64)
    if dtype == np.int16:
        y = np.clip(np.round(y), -32768, 32767)
    y = y.astype(dtype)
    mask = (rng.random(shape) > 0.6)
    for tc0 in (30.0, "polyfit"):
        for use_mask in (False, True):
            fitter = F.MonoExponentialFit(bounds=(0, 100), tc0=tc0, decimal_precision=3)
            tc, r2 = fitter.fit(x8, vols(y, shape, aff), mask=mask if use_mask else None)
            name = f"monoexpfit_{tag}_{'polyfit' if tc0 == 'polyfit' else 'tc30'}_{'mask' if use_mask else 'nomask'}"
            save(name, {"kind": "MonoExponentialFit", "tc0": tc0, "bounds": [0, 100], "decimal_precision": 3,
                        "r2_threshold": 0.9, "use_mask": use_mask, "shape": shape},
                 x=np.asarray(x8), y=y, mask=mask, affine=aff, tc=tc.volume, r2=r2.volume)

# polyfit init with zero / negative samples (SURVEY Appendix B semantics)
y, _, _ = mono_data(rng, x4, n, snr=100, dtype=np.float64)
y[0, :40] = 0.0
y[3, 40:80] = -np.abs(y[3, 40:80])
y[:, 80:100] = 0.0
y = y.astype(np.float32)
tc, r2 = F.MonoExponentialFit(bounds=(0, 100), tc0="polyfit", decimal_precision=3).fit(x4, vols(y, shape))
save("monoexpfit_polyfit_zeros_negatives", {"kind": "MonoExponentialFit", "tc0": "polyfit", "bounds": [0, 100],
     "decimal_precision": 3, "r2_threshold": 0.9, "use_mask": False, "shape": shape},
     x=np.asarray(x4), y=y, affine=np.eye(4), tc=tc.volume, r2=r2.volume)

# CurveFitter: mask -> NaN outside, out_bounds, ufuncs, nan_to_num, r2 threshold
y, _, _ = mono_data(rng, x8, n, snr=30)
mask = rng.random(shape) > 0.5
cf = F.CurveFitter(F.monoexponential, p0=(1.0, -1 / 30))
popt, r2 = cf.fit(x8, vols(y, shape, aff), mask=mask)
save("curvefitter_mask_nan", {"kind": "CurveFitter", "p0": [1.0, -1 / 30], "shape": shape, "r2_threshold": 0.9},
     x=np.asarray(x8), y=y, mask=mask, affine=aff, popt=popt.volume, r2=r2.volume)
cf = F.CurveFitter(F.monoexponential, p0=(1.0, -1 / 30), out_ufuncs=[None, lambda v: 1 / np.abs(v)],
                   out_bounds=[(600, 1400), (0, 60)], r2_threshold=0.99, nan_to_num=-1.0)
popt, r2 = cf.fit(x8, vols(y, shape, aff))
save("curvefitter_post", {"kind": "CurveFitter", "p0": [1.0, -1 / 30], "shape": shape, "r2_threshold": 0.99,
                          "out_bounds": [[600, 1400], [0, 60]], "nan_to_num": -1.0, "ufuncs": [None, "inv_abs"]},
     x=np.asarray(x8), y=y, affine=aff, popt=popt.volume, r2=r2.volume)
print("done")
