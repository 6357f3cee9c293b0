"""One noise-dominated volume (see noise_volume.py) fitted through the default dense path: for profiling.
Usage: python tests/gpu_scripts/noise_one.py [background fraction] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D  # noqa: E402
from dosma_b200 import _cabi, device_api as A  # noqa: E402

dev = torch.device("cuda", 0)
bg = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 384 * 384 * 160
x = np.arange(1, 9) * 10.0
xt = torch.tensor(x, device=dev, dtype=torch.float32)[:, None]
g = torch.Generator(device=dev).manual_seed(5)
a = 500 + 1000 * torch.rand(n, device=dev, generator=g)
t2 = 10 + 70 * torch.rand(n, device=dev, generator=g)
blk = torch.rand((n + 4095) // 4096, device=dev, generator=g) < bg
air = blk.repeat_interleave(4096)[:n]
y = torch.where(air, torch.zeros((), device=dev), a * torch.exp(-xt / t2)) + 10 * torch.randn(8, n, device=dev, generator=g)
popt = torch.empty((n, 2), device=dev)
r2 = torch.empty((n,), device=dev)
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    A.fit_device(o, P, x, y, popt=popt, r2=r2)
    e1.record()
    torch.cuda.synchronize()
    st = _cabi.get_handle(0).stats()
    print(round(e0.elapsed_time(e1), 3), "ms", {k: st[k] for k in ("n_fitted", "n_failed", "n_deferred", "sum_iters", "n_launches")}, flush=True)
