"""Config 4 alone (256 x 256 x 128 voxels, 16 echoes x 5 ms, bi-exponential, SNR 100, fp32): time the kernel variants.
Usage: python tests/gpu_scripts/biexp_c4.py [reps]   -> gpurun_out/biexp_c4.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D  # noqa: E402
from dosma_b200 import _cabi, device_api as A  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
g = torch.Generator(device=dev).manual_seed(3)
x16 = [5.0 * i for i in range(1, 17)]
n = 256 * 256 * 128
xt = torch.tensor(x16, device=dev, dtype=torch.float32)[:, None]
amp = 500 + 1000 * torch.rand(n, device=dev, generator=g)
fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
ts = 8 + 12 * torch.rand(n, device=dev, generator=g)
tl = 50 + 50 * torch.rand(n, device=dev, generator=g)
y = amp * fs * torch.exp(-xt / ts) + amp * (1 - fs) * torch.exp(-xt / tl) + 10 * torch.randn(16, n, device=dev, generator=g)
p0 = (500.0, -1 / 10, 500.0, -1 / 60)
out = {}
ref = None
for name, kw in (("default", {}), ("tma", dict(use_tma=1))):
    o, P = A.make_opts(D.biexponential, p0=p0, **kw)
    popt = torch.empty((n, 4), device=dev)
    r2 = torch.empty((n,), device=dev)
    for _ in range(2):
        A.fit_device(o, P, x16, y, popt=popt, r2=r2)
    torch.cuda.synchronize()
    ts_ = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A.fit_device(o, P, x16, y, popt=popt, r2=r2)
        e1.record()
        torch.cuda.synchronize()
        ts_.append(e0.elapsed_time(e1))
    st = _cabi.get_handle(0).stats()
    ms = float(np.median(ts_))
    rec = {"ms": ms, "voxels_per_s": n / ms * 1e3, "mean_passes": st["sum_iters"] / max(st["n_fitted"], 1), "max_passes": st["max_iters"],
           "failed_fraction": st["n_failed"] / n, "checksum": float(popt.double().nan_to_num(0).sum())}
    if ref is None:
        ref = popt.clone()
    else:
        rec["identical_to_default"] = bool(torch.equal(popt.nan_to_num(-1), ref.nan_to_num(-1)))
    out[name] = rec
    print(name, rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/biexp_c4.json", "w"), indent=1)
