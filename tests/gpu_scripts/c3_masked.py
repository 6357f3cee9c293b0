"""Config 3 alone (512 x 512 x 256, 7-echo T1rho with non-uniform spin-lock times, 2.3 % tissue mask): the masked fit.
Usage: python tests/gpu_scripts/c3_masked.py [reps] [raw|fused]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D  # noqa: E402
from dosma_b200 import _cabi, device_api as A  # noqa: E402

dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
fused = len(sys.argv) > 2 and sys.argv[2] == "fused"
x7 = [0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]
shape = (512, 512, 256)
n = int(np.prod(shape))
g = torch.Generator(device=dev).manual_seed(2)
xt = torch.tensor(x7, device=dev, dtype=torch.float32)[:, None]
a = 500 + 1000 * torch.rand(n, device=dev, generator=g)
t = 20 + 100 * torch.rand(n, device=dev, generator=g)
y = a * torch.exp(-xt / t) + 10 * torch.randn(7, n, device=dev, generator=g)
zz, yy, xx = torch.meshgrid(*[torch.linspace(-1, 1, s, device=dev) for s in shape], indexing="ij")
rad = (zz ** 2 + yy ** 2 + (xx * 1.6) ** 2).sqrt()
mask = ((rad > 0.55) & (rad < 0.62)).reshape(-1)
del zz, yy, xx, rad, a, t
post = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, 3], r2_threshold=0.9, nan_to_num=0.0) if fused else None
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post, **(dict(out_param=1) if fused else {}))
p = torch.empty((n,) if fused else (n, 2), device=dev)
r = torch.empty(n, device=dev)
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    A.fit_device(o, P, x7, y, mask=mask, popt=p, r2=r)
    e1.record()
    torch.cuda.synchronize()
    print(round(e0.elapsed_time(e1), 4), "ms", _cabi.get_handle(0).stats()["n_fitted"], flush=True)
