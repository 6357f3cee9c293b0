"""CPU ORACLE (ctypes front-end of oracle/minpack_lmdif.c) -- test infrastructure, NOT product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  See the header of minpack_lmdif.c for what is restated and how it is pinned.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
MODELS = {"monoexponential": (0, 2), "biexponential": (1, 4), "linear": (2, 1)}
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "minpack_lmdif.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.dosma_fit_voxels.restype = ctypes.c_int
        _lib.dosma_fit_voxels.argtypes = [
            ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, dp,
            ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_double, dp, dp,
            ctypes.POINTER(ctypes.c_int32),
        ]
    return _lib


def curve_fit(model, x, y, p0=None, ftol=1e-5, maxfev=100, eps=1e-8, num_threads=1, want_info=False):
    """(E, N) planar float64 `y` -> popt (N, P), r2 (N,) [, info (N, 2) = (nfev, ier)].

    Same contract as `dosma_oracle.curve_fit` for the built-in models, MINPACK restated in C.
    `p0` is None (ones), a length-P sequence or an (N, P) array.
    """
    lib = _load()
    mid, P = MODELS[model]
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    E, N = y.shape
    p0 = np.ones((1, P)) if p0 is None else np.ascontiguousarray(np.atleast_2d(np.asarray(p0, dtype=np.float64)))
    assert p0.shape in ((1, P), (N, P)), p0.shape
    popt = np.empty((N, P))
    r2 = np.empty(N)
    info = np.zeros((N, 2), dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)

    def run(v0, v1):
        rc = lib.dosma_fit_voxels(mid, E, x.ctypes.data_as(dp), y.ctypes.data_as(dp), N, v0, v1,
                                  p0.ctypes.data_as(dp), p0.shape[0], ftol, maxfev, eps,
                                  popt.ctypes.data_as(dp), r2.ctypes.data_as(dp),
                                  info.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        if rc != 0:
            raise RuntimeError(f"dosma_fit_voxels failed: {rc}")

    if num_threads <= 1 or N < 1024:
        run(0, N)
    else:
        edges = np.linspace(0, N, num_threads * 4 + 1).astype(np.int64)
        with ThreadPoolExecutor(num_threads) as ex:
            list(ex.map(lambda ab: run(int(ab[0]), int(ab[1])), zip(edges[:-1], edges[1:])))
    return (popt, r2, info) if want_info else (popt, r2)
