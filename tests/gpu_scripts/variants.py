"""Times the mono-exponential fit kernel variants on the bench workload (384^3 x 8 echoes, SNR 100):
fast path on/off x plain coalesced loads / TMA-staged ring, and checks that they agree.
Writes gpurun_out/variants.json.  Usage: python tests/gpu_scripts/variants.py [n_voxels]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D  # noqa: E402
from dosma_b200 import _cabi, device_api as A  # noqa: E402

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384 ** 3
g = torch.Generator(device="cuda").manual_seed(1)
x = np.arange(1, 9) * 10.0
xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
a = 500 + 1000 * torch.rand(n, device="cuda", generator=g)
t2 = 10 + 70 * torch.rand(n, device="cuda", generator=g)
y = a * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device="cuda", generator=g)
del a, t2
out = {"voxels": n}
ref = None
popt = torch.empty((n, 2), device="cuda")
r2 = torch.empty((n,), device="cuda")
for name, kw in (("lm_ldg", dict(fast_path=0, use_tma=0)), ("fast1_ldg", dict(fast_path=2, use_tma=0)),
                 ("fast2_ldg", dict(fast_path=1, use_tma=0)), ("fast2_tma", dict(fast_path=1, use_tma=1)),
                 ("default", dict())):
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), **kw)
    A.fit_device(o, P, x, y, popt=popt, r2=r2)
    torch.cuda.synchronize()
    st = _cabi.get_handle(0).stats()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A.fit_device(o, P, x, y, popt=popt, r2=r2)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    rec = {"ms": ms, "voxels_per_s": n / ms * 1e3, "GBps": 44 * n / ms / 1e6, "mean_passes": st["sum_iters"] / max(st["n_fitted"], 1),
           "max_passes": st["max_iters"], "failed": st["n_failed"], "fitted": st["n_fitted"]}
    if ref is None:
        ref = (popt.clone(), r2.clone())
    else:
        rel = ((popt - ref[0]).abs() / ref[0].abs()).nan_to_num(0)
        rec["max_rel_vs_lm"] = float(rel.max())
        rec["p999_rel_vs_lm"] = float(torch.quantile(rel.flatten()[:8_000_000].float(), 0.999))
        rec["max_r2_diff"] = float((r2 - ref[1]).abs().max())
    out[name] = rec
    print(name, rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/variants.json", "w") as f:
    json.dump(out, f, indent=1)
