"""Mono-exponential LM from p0 (fast_path=0) on pure noise (384 x 384 x 160 voxels x 8 echoes): the workload of the LM tail.
Usage: python tests/gpu_scripts/lm_noise.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D  # noqa: E402
from dosma_b200 import _cabi, device_api as A  # noqa: E402

dev = torch.device("cuda", 0)
n = 384 * 384 * 160
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator(device=dev).manual_seed(5)
x = np.arange(1, 9) * 10.0
y = 10 * torch.randn(8, n, device=dev, generator=g)
popt = torch.empty((n, 2), device=dev)
r2 = torch.empty((n,), device=dev)
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), fast_path=0)
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    A.fit_device(o, P, x, y, popt=popt, r2=r2)
    e1.record()
    torch.cuda.synchronize()
    st = _cabi.get_handle(0).stats()
    print(e0.elapsed_time(e1), "ms", st["sum_iters"] / max(st["n_fitted"], 1), "passes", st["n_failed"] / n, "failed", flush=True)
