"""A/B timing of library variants (built by `python -m dosma_b200.build --variant=...`) on the bench workload.
Each variant runs in its own process (DOSMA_B200_LIB selects the library).  Usage:
    python tests/gpu_scripts/ab_libs.py [lib ...]         (default: every dosma_b200/libdfit*.so)
Writes gpurun_out/ab_libs.json."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")

CHILD = r'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1])
import dosma_b200 as D
from dosma_b200 import _cabi, device_api as A
n = 384 ** 3
g = torch.Generator(device="cuda").manual_seed(1)
x = np.arange(1, 9) * 10.0
xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
a = 500 + 1000 * torch.rand(n, device="cuda", generator=g)
t2 = 10 + 70 * torch.rand(n, device="cuda", generator=g)
y = a * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device="cuda", generator=g)
popt = torch.empty((n, 2), device="cuda"); r2 = torch.empty((n,), device="cuda")
o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
out = {}
for name, post in (("raw", None), ("fused_epilogue", dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, 1], r2_threshold=0.9, nan_to_num=0.0))):
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post)
    for _ in range(3):
        A.fit_device(o, P, x, y, popt=popt, r2=r2)
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A.fit_device(o, P, x, y, popt=popt, r2=r2); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = _cabi.get_handle(0).stats()
    out[name] = {"ms_median": float(np.median(ts)), "ms_min": float(np.min(ts)), "mean_passes": st["sum_iters"] / max(st["n_fitted"], 1),
                 "fitted": st["n_fitted"], "checksum": float(popt[:, 1].double().nan_to_num(0).sum())}
print(json.dumps(out))
'''

libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "dosma_b200", "libdfit*.so")))
res = {}
for lib in libs:
    env = dict(os.environ, DOSMA_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(ROOT)], env=env, capture_output=True, text=True)
    name = os.path.basename(lib)
    try:
        res[name] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        res[name] = {"error": (r.stderr or r.stdout)[-600:]}
    print(name, res[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "ab_libs.json"), "w") as f:
    json.dump(res, f, indent=1)
