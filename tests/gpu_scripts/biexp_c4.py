"""Config 4 alone (256 x 256 x 128 voxels, 16 echoes x 5 ms, bi-exponential, fp32), noise-free (the configuration table's
line) and at SNR 100: time the plain one-voxel-per-lane LM kernel (DFIT_LMQ=0) against the LM-in-rounds kernel
(fit_kernel_lmq) at several round budgets, and check that the results are bit-identical.
Usage: python tests/gpu_scripts/biexp_c4.py [reps] [budget,budget ...]   -> gpurun_out/biexp_c4.json"""
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
budgets = sys.argv[2:] or ["4,3", "4,2", "5,3", "3,3", "6,4", "4,4"]
g = torch.Generator(device=dev).manual_seed(3)
x16 = [5.0 * i for i in range(1, 17)]
n = 256 * 256 * 128
xt = torch.tensor(x16, device=dev, dtype=torch.float32)[:, None]
amp = 500 + 1000 * torch.rand(n, device=dev, generator=g)
fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
ts = 8 + 12 * torch.rand(n, device=dev, generator=g)
tl = 50 + 50 * torch.rand(n, device=dev, generator=g)
clean = amp * fs * torch.exp(-xt / ts) + amp * (1 - fs) * torch.exp(-xt / tl)
p0 = (500.0, -1 / 10, 500.0, -1 / 60)
out = {}
for data, y in (("clean", clean), ("snr100", clean + 10 * torch.randn(16, n, device=dev, generator=g))):
    ref = None
    for name in ["0"] + budgets:
        os.environ["DFIT_LMQ"] = name
        o, P = A.make_opts(D.biexponential, p0=p0)
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
        rec = {"ms": ms, "voxels_per_s": n / ms * 1e3, "mean_passes": st["sum_iters"] / max(st["n_fitted"], 1),
               "max_passes": st["max_iters"], "failed_fraction": st["n_failed"] / n,
               "checksum": float(popt.double().nan_to_num(0).sum())}
        if ref is None:
            ref = (popt.clone(), r2.clone())
        else:
            rec["identical_to_plain"] = bool(torch.equal(popt.view(torch.int32), ref[0].view(torch.int32)) and
                                             torch.equal(r2.view(torch.int32), ref[1].view(torch.int32)))
        out[f"{data}_lmq_{name}"] = rec
        print(data, name, rec, flush=True)
os.environ.pop("DFIT_LMQ", None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/biexp_c4.json", "w"), indent=1)
