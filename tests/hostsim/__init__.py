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


def set_rounds(k_first, k_next=None):
    """LM in rounds (what fit_kernel_lmq does): lm_begin, then lm_iterate with `k_first` / `k_next` evaluations per round,
    the solver state parked in between.  0 switches back to lm_solve."""
    _load().hostsim_set_rounds(ctypes.c_int(int(k_first)), ctypes.c_int(int(k_first if k_next is None else k_next)))


def set_uniform_recurrence(on):
    """With set_rounds(k > 0): take the model's exponentials from the two-echo recurrence on uniformly spaced echoes, as
    fit_kernel_lmq does (fp32, >= 4 echoes)."""
    _load().hostsim_set_uniform_recurrence(ctypes.c_int(int(bool(on))))


def engine_fit(model_id, nparams, x, planes, mask, p0_cols, *, init_mode=0, y_bounds=None, maxfev=100, ftol=1e-5,
               eps=1e-8, post=None, engine=None, out_param=None):
    """TEST-ONLY stand-in for `dosma_b200.fitting._engine_fit` with the same signature and return values: the device
    solver headers compiled by g++ run the fit, the fused epilogue and the mask fill on the host.  It lets the CPU
    suite drive the drop-in's Python layer (argument handling, marshalling, headers) end to end -- e.g. through
    the reference's own test-suite -- in the build container, where there is no GPU.  Never used by the product."""
    lib = _load()
    engine = dict(engine or {})
    model = {0: "monoexponential", 1: "biexponential", 2: "linear"}[model_id]
    planes = [np.asarray(p).reshape(-1) for p in planes]
    dt = np.result_type(*[p.dtype for p in planes])
    cd = engine.get("compute_dtype") or "auto"
    if cd == "auto":
        cd = "f64" if dt == np.float64 else "f32"
    y = np.stack([p.astype(np.float64) for p in planes])
    if cd == "f32":
        y = y.astype(np.float32).astype(np.float64)
    E, N = y.shape
    sel = np.ones(N, dtype=bool) if mask is None else (np.asarray(mask).reshape(-1) != 0)
    if not np.isfinite(y[:, sel]).all():
        raise ValueError("array must not contain infs or NaNs")
    per_voxel = any(isinstance(c, np.ndarray) for c in p0_cols)
    if init_mode != 0:
        p0 = np.ones((1, nparams))
    elif per_voxel:
        p0 = np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), (N,)) for c in p0_cols], axis=1)[sel]
    else:
        p0 = np.asarray([[float(c) for c in p0_cols]])
    f32 = cd == "f32"
    e = 1.1920929e-7 if f32 else 2.220446049250313e-16
    eng_ftol = max(ftol * float(engine.get("ftol_scale") or 1e-2), 1e-8 if f32 else 1e-14)
    fast_opt = engine.get("fast_path")
    fast_on = model_id == 0 and fast_opt != 0 and maxfev >= 21
    fast = (2 if f32 and fast_opt != 2 else 1) if fast_on else 0
    il = engine.get("init_linear")
    il = (0 if model_id == 1 else 1) if il is None or il < 0 else il
    popt_s, r2_s, st, it = fit(model, x, y[:, sel], p0=p0, dtype=cd, init_mode=init_mode, init_linear=il, ftol=eng_ftol,
                               xtol=float(engine.get("xtol") or (1e-6 if f32 else 1e-10)),
                               lambda0=float(engine.get("lambda0") or 1e-3), floor_rel=(8 * e) ** 2, maxfev=maxfev,
                               r2_eps=eps, y_bounds=y_bounds, fast=fast)
    if post:
        I4, D4 = ctypes.c_int * 4, ctypes.c_double * 4
        pad = lambda a, v: list(a) + [v] * (4 - len(a))  # noqa: E731
        popt_s = np.ascontiguousarray(popt_s)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.hostsim_post_params(ctypes.c_int64(popt_s.shape[0]), ctypes.c_int(nparams), popt_s.ctypes.data_as(dp),
                                np.ascontiguousarray(r2_s).ctypes.data_as(dp), I4(*pad(post["ufunc"], 0)),
                                D4(*pad(post["lb"], -np.inf)), D4(*pad(post["ub"], np.inf)),
                                int(post.get("r2_threshold") is not None), ctypes.c_double(post.get("r2_threshold") or 0.0),
                                int(post.get("nan_to_num") is not None), ctypes.c_double(post.get("nan_to_num") or 0.0),
                                I4(*pad(post["decimals"], -1)))
    fill = np.nan
    if post and post.get("nan_to_num") is not None:
        fill = float(post["nan_to_num"])
    popt = np.full((N, nparams), fill)
    r2 = np.full(N, fill)
    if post and not np.isnan(fill):
        for i, d in enumerate(post["decimals"]):
            if d >= 0:
                popt[:, i] = np.around(fill, d)
    popt[sel] = popt_s
    r2[sel] = r2_s
    out_dt = np.float32 if engine.get("out_dtype") == "f32" else np.float64
    stats = {"n_voxels": N, "n_fitted": int(((st >= 1) & (st <= 4)).sum()), "n_failed": int((st >= 5).sum()),
             "n_nonfinite": 0, "n_oob": 0, "sum_iters": int(it.sum()), "max_iters": int(it.max()) if it.size else 0,
             "n_launches": 0, "kernel_ms": -1.0, "total_ms": -1.0}
    if out_param is not None:
        popt = np.ascontiguousarray(popt[:, out_param])
    return popt.astype(out_dt), r2.astype(out_dt), stats
