"""TEST-ONLY: g++ build of the device solver header, see hostsim.cpp.  Never imported by dosma_b200/."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostsim.so")
_lib = None
MODELS = {"monoexponential": (0, 2), "biexponential": (1, 4), "linear": (2, 1)}


def _load():
    global _lib
    if _lib is None:
        csrc = os.path.join(_HERE, "..", "..", "dosma_b200", "csrc")
        srcs = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(csrc, "lm_core.cuh"), os.path.join(csrc, "mono_fast.cuh")]
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=fast", "-march=native",
                                   "-o", _SO, srcs[0]])
        _lib = ctypes.CDLL(_SO)
        _lib.hostsim_fit.restype = ctypes.c_int
        _lib.hostsim_post_param.restype = ctypes.c_double
    return _lib


def fit(model, x, y, p0=None, dtype="f32", acc64=False, init_mode=0, init_linear=1, ftol=None, xtol=None,
        lambda0=1e-3, floor_rel=None, maxfev=100, r2_eps=1e-8, y_bounds=None, fast=1):
    lib = _load()
    mid, P = MODELS[model]
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    E, N = y.shape
    p0 = np.ones((1, P)) if p0 is None else np.ascontiguousarray(np.atleast_2d(np.asarray(p0, dtype=np.float64)))
    f32 = dtype == "f32"
    eps = 1.19e-7 if f32 else 2.2e-16
    ftol = (1e-7 if f32 else 1e-12) if ftol is None else ftol
    xtol = (1e-6 if f32 else 1e-10) if xtol is None else xtol
    floor_rel = (32 * eps) ** 2 if floor_rel is None else floor_rel
    lo, hi = (-np.inf, np.inf) if y_bounds is None else y_bounds
    popt = np.empty((N, P))
    r2 = np.empty(N)
    status = np.empty(N, dtype=np.int32)
    iters = np.empty(N, dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int32)
    rc = lib.hostsim_fit(ctypes.c_int(mid), ctypes.c_int(0 if f32 else 1), ctypes.c_int(int(acc64)), ctypes.c_int(E),
                         ctypes.c_int64(N), x.ctypes.data_as(dp), y.ctypes.data_as(dp), p0.ctypes.data_as(dp),
                         ctypes.c_int64(p0.shape[0]), ctypes.c_int(init_mode), ctypes.c_int(init_linear), ctypes.c_int(int(fast)),
                         ctypes.c_double(ftol), ctypes.c_double(xtol), ctypes.c_double(lambda0),
                         ctypes.c_double(floor_rel), ctypes.c_int(maxfev), ctypes.c_double(r2_eps),
                         ctypes.c_double(lo), ctypes.c_double(hi), popt.ctypes.data_as(dp), r2.ctypes.data_as(dp),
                         status.ctypes.data_as(ip), iters.ctypes.data_as(ip))
    assert rc == 0, rc
    return popt, r2, status, iters
