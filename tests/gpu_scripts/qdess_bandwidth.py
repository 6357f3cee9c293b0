"""Device-resident bandwidth of the qDESS T2-map kernel (8 B read + 4 B written per voxel)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dosma_b200 import _cabi  # noqa: E402
from dosma_b200.qdess import qdess_constants  # noqa: E402

n = 512 * 512 * 512
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
s1 = 200 + 1000 * torch.rand(n, device=dev, generator=g)
s2 = s1 * (0.1 + 0.3 * torch.rand(n, device=dev, generator=g))
out = torch.empty(n, device=dev)
out64 = torch.empty(n, device=dev, dtype=torch.float64)
lib = _cabi.load()
o = _cabi.DfitQdessOpts()
_cabi.check(lib.dfit_default_qdess_opts(ctypes.byref(o)))
o.k, o.c1, o.tr_minus_te = qdess_constants(20.36, 6.43, 3400.0, 3132.0, 20.0, 1200.0)
h = _cabi.get_handle(0)
res = {}
for name, fat, cd in (("fast_f32", 0, _cabi.F32), ("fast_f32_suppress_fat+fluid", 1, _cabi.F32), ("exact_f64", 0, _cabi.F64)):
    o.suppress_fat = o.suppress_fluid = fat
    o.compute_dtype = cd
    dst, odt, bpv = (out, _cabi.F32, 12) if cd == _cabi.F32 else (out64, _cabi.F64, 16)
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check(lib.dfit_qdess_t2_device(h.ptr, ctypes.byref(o), n, s1.data_ptr(), s2.data_ptr(), _cabi.F32,
                                             dst.data_ptr(), odt, None))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    res[name] = {"ms": ms, "voxels_per_s": n / ms * 1e3, "bytes_per_voxel": bpv, "GBps_algorithmic": bpv * n / ms / 1e6}
print(json.dumps({"kernel": "qdess_kernel", "voxels": n, **res}))
