"""A/B of library variants on the LM workloads: pure-noise LM (fast_path=0), tissue LM (fast_path=0), the 70 %-air volume
through the default path, config 4 (bi-exponential, clean and SNR 100).  One child process per library.
Usage: python tests/gpu_scripts/ab_lm.py [lib ...]  -> gpurun_out/ab_lm.json"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
CHILD = r'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1])
import dosma_b200 as D
from dosma_b200 import _cabi, device_api as A
dev = torch.device("cuda", 0)
def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
out = {}
n = 384 * 384 * 160
x = np.arange(1, 9) * 10.0
xt = torch.tensor(x, device=dev, dtype=torch.float32)[:, None]
g = torch.Generator(device=dev).manual_seed(5)
a = 500 + 1000 * torch.rand(n, device=dev, generator=g); t2 = 10 + 70 * torch.rand(n, device=dev, generator=g)
noise = 10 * torch.randn(8, n, device=dev, generator=g)
tissue = a * torch.exp(-xt / t2) + noise
air = (torch.rand((n + 4095) // 4096, device=dev, generator=g) < 0.7).repeat_interleave(4096)[:n]
mixed = torch.where(air, torch.zeros((), device=dev), a * torch.exp(-xt / t2)) + noise
popt = torch.empty((n, 2), device=dev); r2 = torch.empty((n,), device=dev)
o_lm, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), fast_path=0)
o_def, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
out["lm_noise"] = timed(lambda: A.fit_device(o_lm, P, x, noise, popt=popt, r2=r2))
out["lm_tissue"] = timed(lambda: A.fit_device(o_lm, P, x, tissue, popt=popt, r2=r2))
out["default_air70"] = timed(lambda: A.fit_device(o_def, P, x, mixed, popt=popt, r2=r2))
out["default_tissue"] = timed(lambda: A.fit_device(o_def, P, x, tissue, popt=popt, r2=r2), reps=15)
del noise, tissue, mixed, popt, r2
x16 = [5.0 * i for i in range(1, 17)]; n = 256 * 256 * 128
g = torch.Generator(device=dev).manual_seed(3)
xt = torch.tensor(x16, device=dev, dtype=torch.float32)[:, None]
amp = 500 + 1000 * torch.rand(n, device=dev, generator=g); fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
ts = 8 + 12 * torch.rand(n, device=dev, generator=g); tl = 50 + 50 * torch.rand(n, device=dev, generator=g)
clean = amp * fs * torch.exp(-xt / ts) + amp * (1 - fs) * torch.exp(-xt / tl)
o_b, P = A.make_opts(D.biexponential, p0=(500.0, -1 / 10, 500.0, -1 / 60))
popt = torch.empty((n, 4), device=dev); r2 = torch.empty((n,), device=dev)
out["biexp_clean"] = timed(lambda: A.fit_device(o_b, P, x16, clean, popt=popt, r2=r2))
noisy = clean + 10 * torch.randn(16, n, device=dev, generator=g)
out["biexp_snr100"] = timed(lambda: A.fit_device(o_b, P, x16, noisy, popt=popt, r2=r2))
print(json.dumps(out))
'''
libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "dosma_b200", "libdfit*.so")))
res = {}
for lib in libs:
    env = dict(os.environ, DOSMA_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True)
    name = os.path.basename(lib)
    try:
        res[name] = {k: round(v, 4) for k, v in json.loads(r.stdout.strip().splitlines()[-1]).items()}
    except Exception:
        res[name] = {"error": (r.stderr or r.stdout)[-800:]}
    print(name, res[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_lm.json"), "w"), indent=1)
