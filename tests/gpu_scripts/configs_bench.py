"""Time and sanity-check every BASELINE.json configuration on one GPU (device-resident samples).

    python tests/gpu_scripts/configs_bench.py > gpurun_out/configs.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dosma_b200 as D  # noqa: E402
from dosma_b200 import _cabi, device_api as A  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
h = _cabi.get_handle(0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def mono(n, x, seed, t_rng=(10, 80), sigma=10.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    xt = torch.tensor(x, device=dev, dtype=torch.float32)[:, None]
    a = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    t = t_rng[0] + (t_rng[1] - t_rng[0]) * torch.rand(n, device=dev, generator=g)
    y = a * torch.exp(-xt / t)
    if sigma:
        y += sigma * torch.randn(len(x), n, device=dev, generator=g)
    return y, a, t


out = {}

# C1: 64x64x16, 4 echoes -- the reference's CPU-runnable case: parity against the C oracle on the spot
x4 = [10.0, 20.0, 40.0, 80.0]
y, a, t = mono(64 * 64 * 16, x4, 0, sigma=0)
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
p, r = A.fit_device(o, P, x4, y)
from oracle import c_oracle  # noqa: E402

ref, _ = c_oracle.curve_fit("monoexponential", np.asarray(x4), y.cpu().numpy().astype(np.float64), p0=(1.0, -1 / 30),
                            num_threads=8)
rel = np.abs(p.cpu().numpy() - ref) / np.abs(ref)
out["C1_64x64x16_4echo"] = {"voxels": y.shape[1], "ms": timed(lambda: A.fit_device(o, P, x4, y, popt=p, r2=r)),
                            "max_rel_err_vs_oracle": float(rel.max())}

# C2: 384x384x160, 8 echoes, fp32
x8 = [10.0 * i for i in range(1, 9)]
n = 384 * 384 * 160
y, a, t = mono(n, x8, 1)
p = torch.empty((n, 2), device=dev)
r = torch.empty(n, device=dev)
for name, kw in (("p0_tc30", {}), ("loglinear_init", {"init": "loglinear"})):
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), **kw)
    ms = timed(lambda: A.fit_device(o, P, x8, y, popt=p, r2=r))
    st = h.stats()
    out[f"C2_384x384x160_8echo_{name}"] = {"voxels": n, "ms": ms, "voxels_per_s": n / ms * 1e3,
                                           "mean_passes": st["sum_iters"] / st["n_fitted"], "failed": st["n_failed"],
                                           "median_rel_err_b": float((((p[:, 1] + 1 / t) * t).abs()).median())}
# fused MonoExponentialFit epilogue (ufunc, bounds, r2 threshold, fill, rounding)
post = {"ufunc": [0, 1], "lb": [-np.inf, 0.0], "ub": [np.inf, 100.0], "decimals": [-1, 3], "r2_threshold": 0.9,
        "nan_to_num": 0.0}
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post)
ms = timed(lambda: A.fit_device(o, P, x8, y, popt=p, r2=r))
out["C2_fused_monoexpfit_epilogue"] = {"voxels": n, "ms": ms, "voxels_per_s": n / ms * 1e3,
                                       "tc_median_abs_err_ms": float((p[:, 1] - t).abs().median())}
# int16 samples (DICOM): half the read traffic
y16 = y.clamp(-32768, 32767).round().to(torch.int16)
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
ms = timed(lambda: A.fit_device(o, P, x8, y16, popt=p, r2=r))
out["C2_int16_samples"] = {"voxels": n, "ms": ms, "voxels_per_s": n / ms * 1e3}
del y, y16, p, r

# C3: 512x512x256, 7-echo T1rho, ellipsoid-shell tissue mask (~10 %)
x7 = [0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]
shape = (512, 512, 256)
n = int(np.prod(shape))
y, a, t = mono(n, x7, 2, t_rng=(20, 120))
zz, yy, xx = torch.meshgrid(*[torch.linspace(-1, 1, s, device=dev) for s in shape], indexing="ij")
rad = (zz ** 2 + yy ** 2 + (xx * 1.6) ** 2).sqrt()
mask = ((rad > 0.55) & (rad < 0.62)).reshape(-1)
del zz, yy, xx, rad
p = torch.empty((n, 2), device=dev)
r = torch.empty(n, device=dev)
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
ms_dense = timed(lambda: A.fit_device(o, P, x7, y, popt=p, r2=r), reps=3)
p_dense = p.clone()
ms_mask = timed(lambda: A.fit_device(o, P, x7, y, mask=mask, popt=p, r2=r), reps=3)
same = torch.equal(p[mask], p_dense[mask]) and bool(torch.isnan(p[~mask]).all())
out["C3_512x512x256_7echo_t1rho"] = {"voxels": n, "mask_fraction": float(mask.float().mean()), "ms_dense": ms_dense,
                                     "voxels_per_s_dense": n / ms_dense * 1e3, "ms_masked": ms_mask,
                                     "masked_voxels_per_s": float(mask.sum()) / ms_mask * 1e3,
                                     "masked_equals_dense_inside_and_nan_outside": same}
del y, p, r, p_dense, mask

# C4: 256x256x128, 16-echo bi-exponential
x16 = [5.0 * i for i in range(1, 17)]
n = 256 * 256 * 128
g = torch.Generator(device=dev).manual_seed(3)
xt = torch.tensor(x16, device=dev, dtype=torch.float32)[:, None]
Aamp = 500 + 1000 * torch.rand(n, device=dev, generator=g)
fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
ts = 8 + 12 * torch.rand(n, device=dev, generator=g)
tl = 50 + 50 * torch.rand(n, device=dev, generator=g)
y = Aamp * fs * torch.exp(-xt / ts) + Aamp * (1 - fs) * torch.exp(-xt / tl)
p = torch.empty((n, 4), device=dev)
r = torch.empty(n, device=dev)
for cd in ("f32", "f64"):
    o, P = A.make_opts(D.biexponential, p0=(500.0, -1 / 10, 500.0, -1 / 60), compute_dtype=cd)
    ms = timed(lambda: A.fit_device(o, P, x16, y, popt=p, r2=r), reps=3)
    st = h.stats()
    ok = ~torch.isnan(p[:, 0])
    err = ((p[ok, 1] + 1 / ts[ok]) * ts[ok]).abs()
    out[f"C4_256x256x128_16echo_biexp_{cd}"] = {"voxels": n, "ms": ms, "voxels_per_s": n / ms * 1e3,
                                                "mean_passes": st["sum_iters"] / max(st["n_fitted"], 1),
                                                "failed_fraction": st["n_failed"] / n,
                                                "frac_b1_within_1e-3_of_truth": float((err < 1e-3).float().mean())}
del y, p, r

# C5: one GPU's share of "32 subjects x 384x384x64 x 8 echoes over 8 GPUs" = 4 subjects
n = 4 * 384 * 384 * 64
y, a, t = mono(n, x8, 100)
p = torch.empty((n, 2), device=dev)
r = torch.empty(n, device=dev)
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
ms = timed(lambda: A.fit_device(o, P, x8, y, popt=p, r2=r))
out["C5_4subjects_384x384x64_8echo_per_gpu"] = {"voxels": n, "ms": ms, "voxels_per_s": n / ms * 1e3}
print(json.dumps(out, indent=1))
