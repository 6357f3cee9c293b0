import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import dosma_b200 as D
from dosma_b200 import device_api as A, _cabi
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(3)
x = np.arange(1, 9) * 10.0
xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
for n in (1000, 100_003 // 4 * 4, 384 * 384 * 384):
    a = 500 + 1000 * torch.rand(n, device="cuda", generator=g)
    t2 = 10 + 70 * torch.rand(n, device="cuda", generator=g)
    y = a * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device="cuda", generator=g)
    o0, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), use_tma=0)
    o1, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), use_tma=1)
    p0_, r0_ = A.fit_device(o0, P, x, y)
    torch.cuda.synchronize()
    s0 = _cabi.get_handle(0).stats()
    p1_, r1_ = A.fit_device(o1, P, x, y)
    torch.cuda.synchronize()
    s1 = _cabi.get_handle(0).stats()
    same = torch.equal(p0_.nan_to_num(-1), p1_.nan_to_num(-1)) and torch.equal(r0_, r1_)
    print(n, "tma == ldg:", same, s0["kernel_ms"], s1["kernel_ms"], s0["sum_iters"] == s1["sum_iters"], flush=True)
    if n > 1e7:
        for name, o in (("ldg", o0), ("tma", o1)):
            ts = []
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); A.fit_device(o, P, x, y, popt=p0_, r2=r0_); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(name, "ms:", [round(t, 3) for t in ts], flush=True)
