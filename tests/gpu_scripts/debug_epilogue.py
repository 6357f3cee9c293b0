"""Debug: where does the fp32 fused epilogue differ from the float64 one on the GPU?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import dosma_b200 as D
from dosma_b200 import device_api as A
g = torch.Generator(device="cuda").manual_seed(21)
x = np.arange(1, 9) * 10.0
xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
n = 1_000_000
t2 = 5 + 120 * torch.rand(n, device="cuda", generator=g)
y = (500 + 1000 * torch.rand(n, device="cuda", generator=g)) * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device="cuda", generator=g)
o_raw, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
praw, rraw = A.fit_device(o_raw, P, x, y)
for decimals in (1, 3, -1):
    post = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, decimals], r2_threshold=0.9, nan_to_num=0.0)
    for kw in (dict(), dict(use_tma=0), dict(fast_path=2, use_tma=0), dict(fast_path=0, use_tma=0)):
        o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post, **kw)
        p32, r32 = A.fit_device(o, P, x, y, out_dtype=torch.float32)
        p64, r64 = A.fit_device(o, P, x, y, out_dtype=torch.float64)
        torch.cuda.synchronize()
        d0 = (p32[:, 0] != p64[:, 0].float()); d1 = (p32[:, 1] != p64[:, 1].float()); dr = (r32 != r64.float())
        print(decimals, kw, "mismatch a/tc/r2:", int(d0.sum()), int(d1.sum()), int(dr.sum()), flush=True)
        for name, d in (("a", d0), ("tc", d1), ("r2", dr)):
            idx = torch.nonzero(d)[:6, 0]
            for i in idx.tolist():
                print("   ", name, i, "f32:", p32[i].tolist(), float(r32[i]), "f64:", p64[i].tolist(), float(r64[i]), "raw:", praw[i].tolist(), float(rraw[i]),
                      "1/|b| f64:", 1.0 / abs(float(praw[i, 1].double())))
