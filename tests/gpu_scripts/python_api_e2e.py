"""End-to-end time of the Python drop-in API on host (numpy, pageable) data: curve_fit and MonoExponentialFit
on BASELINE config 2 (384 x 384 x 160, 8 echoes).  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D  # noqa: E402

shape = (384, 384, 160)
n = int(np.prod(shape))
rng = np.random.default_rng(1)
x = np.arange(1, 9) * 10.0
a = rng.uniform(500, 1500, n).astype(np.float32)
t2 = rng.uniform(10, 80, n).astype(np.float32)
y = np.empty((8, n), dtype=np.float32)
for e in range(8):
    y[e] = a * np.exp(-np.float32(x[e]) / t2) + rng.normal(0, 10, n).astype(np.float32)
out = {"voxels": n}
D.curve_fit(D.monoexponential, x, y[:, :100000], p0=(1.0, -1 / 30))  # warm-up: context, handle, buffers
for rep in range(2):
    t0 = time.perf_counter()
    popt, r2 = D.curve_fit(D.monoexponential, x, y, p0=(1.0, -1 / 30))
    dt = time.perf_counter() - t0
out["curve_fit_s"] = dt
out["curve_fit_voxels_per_s"] = n / dt
for rep in range(2):
    t0 = time.perf_counter()
    popt32, r232 = D.curve_fit(D.monoexponential, x, y, p0=(1.0, -1 / 30), out_dtype="f32")
    dt = time.perf_counter() - t0
out["curve_fit_f32_maps_s"] = dt
out["curve_fit_f32_maps_voxels_per_s"] = n / dt
vols = [D.MedicalVolume(y[e].reshape(shape), np.eye(4)) for e in range(8)]
for rep in range(2):
    t0 = time.perf_counter()
    tc, r2v = D.MonoExponentialFit(tc0="polyfit", decimal_precision=3).fit(x, vols)
    dt = time.perf_counter() - t0
out["monoexpfit_s"] = dt
out["monoexpfit_voxels_per_s"] = n / dt
out["tc_median_abs_err_ms"] = float(np.median(np.abs(tc.volume.reshape(-1) - t2)))
print(json.dumps(out))
