"""Small fits through every dense / masked kernel variant, for compute-sanitizer runs:
    compute-sanitizer --tool memcheck python tests/gpu_scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python tests/gpu_scripts/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D  # noqa: E402
from dosma_b200 import device_api as A  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for x in (np.arange(1, 9) * 10.0, np.array([0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0])):
    xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
    for n in (20_032, 20_033, 20_036):
        y = (500 + 1000 * torch.rand(n, device="cuda", generator=g)) * torch.exp(
            -xt / (10 + 70 * torch.rand(n, device="cuda", generator=g))) + 10 * torch.randn(len(x), n, device="cuda", generator=g)
        y[:, 7] = 0
        mask = torch.rand(n, device="cuda", generator=g) > 0.5
        for kw in (dict(), dict(use_tma=0), dict(fast_path=2), dict(fast_path=0)):
            o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), **kw)
            p, r = A.fit_device(o, P, x, y)
            pm, rm = A.fit_device(o, P, x, y, mask=mask)
            pi, ri = A.fit_device(o, P, x, y.round().to(torch.int16))
            torch.cuda.synchronize()
            assert torch.isfinite(p[mask & (torch.arange(n, device="cuda") != 7)]).all()
# the LM in rounds (bi-exponential; dense, masked, ragged; budgets that suspend often), float64 maps with status bytes
x16 = np.arange(1, 13) * 6.0
xt = torch.tensor(x16, device="cuda", dtype=torch.float32)[:, None]
for n in (5_000, 5_003):
    amp = 500 + 1000 * torch.rand(n, device="cuda", generator=g)
    y = amp * 0.5 * torch.exp(-xt / (8 + 12 * torch.rand(n, device="cuda", generator=g))) + amp * 0.5 * torch.exp(
        -xt / (50 + 50 * torch.rand(n, device="cuda", generator=g))) + 10 * torch.randn(12, n, device="cuda", generator=g)
    mask = torch.rand(n, device="cuda", generator=g) > 0.7
    for budgets in ("5,2", "1,1"):
        os.environ["DFIT_LMQ"] = budgets
        o, P = A.make_opts(D.biexponential, p0=(500.0, -1 / 10, 500.0, -1 / 60))
        A.fit_device(o, P, x16, y)
        A.fit_device(o, P, x16, y, mask=mask)
        A.fit_device(o, P, x16, y, mask=mask, out_dtype=torch.float64, status=torch.empty(n, device="cuda", dtype=torch.uint8),
                     niter=torch.empty(n, device="cuda", dtype=torch.uint8))
        torch.cuda.synchronize()
os.environ.pop("DFIT_LMQ", None)
# the LM tail of the dense kernel: a volume that is half noise, above the 2^20 voxels from which the tail is used
n = (1 << 20) + 4096
x = np.arange(1, 9) * 10.0
xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
air = (torch.rand(n // 1024 + 1, device="cuda", generator=g) < 0.5).repeat_interleave(1024)[:n]
y = torch.where(air, torch.zeros((), device="cuda"), 1000 * torch.exp(-xt / (10 + 70 * torch.rand(n, device="cuda", generator=g)))) \
    + 10 * torch.randn(8, n, device="cuda", generator=g)
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
for _ in range(2):
    A.fit_device(o, P, x, y)
torch.cuda.synchronize()
print("sanitize_small ok")
