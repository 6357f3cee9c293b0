"""ctypes binding of libdfit.so (include/dfit.h).  Thin on purpose: structures, prototypes, errors.

The library is loaded lazily and *loudly*: if it is missing or no CUDA device is visible the
compute entry points raise -- there is no CPU fallback anywhere in this package.
"""
import ctypes
import os
import threading

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# DOSMA_B200_LIB: developer hook for kernel A/B runs (a variant built by `python -m dosma_b200.build --variant=...`)
LIB_PATH = os.environ.get("DOSMA_B200_LIB") or os.path.join(_PKG, "libdfit.so")

MAX_PARAMS = 4
MAX_ECHOES = 32

# enums of include/dfit.h
MODEL_MONOEXP, MODEL_BIEXP, MODEL_LINEAR = 0, 1, 2
F32, F64, I16, U16, I32, U8 = 0, 1, 2, 3, 4, 5
PLANAR, ECHO_FASTEST = 0, 1
INIT_GIVEN, INIT_LOGLINEAR = 0, 1
UFUNC_NONE, UFUNC_INV_ABS, UFUNC_NEG_INV, UFUNC_ABS, UFUNC_INV = 0, 1, 2, 3, 4

NP_TO_DTYPE = {
    np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int16): I16,
    np.dtype(np.uint16): U16, np.dtype(np.int32): I32, np.dtype(np.uint8): U8,
}

EXPORTED_SYMBOLS = (
    "dfit_version", "dfit_device_count", "dfit_strerror", "dfit_last_error", "dfit_default_opts",
    "dfit_model_nparams", "dfit_create", "dfit_destroy", "dfit_fit_device", "dfit_fit_host", "dfit_get_stats",
    "dfit_set_gather", "dfit_set_gather_ex", "dfit_ipc_alloc", "dfit_ipc_open", "dfit_ipc_close", "dfit_ipc_free",
    "dfit_default_qdess_opts", "dfit_qdess_t2_device", "dfit_qdess_t2_host", "dfit_region_metrics_host",
)


class DfitOpts(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("model", ctypes.c_int32),
        ("compute_dtype", ctypes.c_int32),
        ("init_mode", ctypes.c_int32),
        ("init_linear", ctypes.c_int32),
        ("maxfev", ctypes.c_int32),
        ("ftol", ctypes.c_double),
        ("ftol_scale", ctypes.c_double),
        ("xtol", ctypes.c_double),
        ("lambda0", ctypes.c_double),
        ("r2_eps", ctypes.c_double),
        ("y_lo", ctypes.c_double),
        ("y_hi", ctypes.c_double),
        ("p0", ctypes.c_double * MAX_PARAMS),
        ("post_enabled", ctypes.c_int32),
        ("ufunc", ctypes.c_int32 * MAX_PARAMS),
        ("lb", ctypes.c_double * MAX_PARAMS),
        ("ub", ctypes.c_double * MAX_PARAMS),
        ("has_r2_threshold", ctypes.c_int32),
        ("r2_threshold", ctypes.c_double),
        ("has_nan_fill", ctypes.c_int32),
        ("nan_fill", ctypes.c_double),
        ("decimals", ctypes.c_int32 * MAX_PARAMS),
        ("fast_path", ctypes.c_int32),
        ("use_tma", ctypes.c_int32),
        ("out_param", ctypes.c_int32),
    ]


class DfitQdessOpts(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("k", ctypes.c_double),
        ("c1", ctypes.c_double),
        ("tr_minus_te", ctypes.c_double),
        ("has_bounds", ctypes.c_int32),
        ("lb", ctypes.c_double),
        ("ub", ctypes.c_double),
        ("has_nan_fill", ctypes.c_int32),
        ("nan_fill", ctypes.c_double),
        ("decimals", ctypes.c_int32),
        ("suppress_fat", ctypes.c_int32),
        ("suppress_fluid", ctypes.c_int32),
        ("beta", ctypes.c_double),
        ("compute_dtype", ctypes.c_int32),
    ]


class DfitGatherDesc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("world", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("maps", ctypes.c_void_p),
        ("multicast", ctypes.c_void_p),
        ("rows", ctypes.c_int64),
        ("row0", ctypes.c_int64),
        ("param_mask", ctypes.c_uint32),
        ("split_list", ctypes.c_int32),
        ("fit_lo", ctypes.c_int64),
        ("fit_hi", ctypes.c_int64),
        ("y_voxel0", ctypes.c_int64),
    ]


class DfitStats(ctypes.Structure):
    _fields_ = [
        ("n_voxels", ctypes.c_int64),
        ("n_fitted", ctypes.c_int64),
        ("n_failed", ctypes.c_int64),
        ("n_nonfinite", ctypes.c_int64),
        ("n_oob", ctypes.c_int64),
        ("sum_iters", ctypes.c_int64),
        ("max_iters", ctypes.c_int32),
        ("n_launches", ctypes.c_int32),
        ("kernel_ms", ctypes.c_float),
        ("total_ms", ctypes.c_float),
        ("n_deferred", ctypes.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DfitError(RuntimeError):
    def __init__(self, code, detail):
        super().__init__(f"libdfit error {code}: {detail}")
        self.code = code


_lib = None
_lock = threading.Lock()


def load():
    """Load libdfit.so (building nothing: see dosma_b200.build).  Raises if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found. Build it with `python -m dosma_b200.build` (needs nvcc). "
                "dosma_b200 has no CPU fallback: the CUDA library is required."
            )
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        lib.dfit_version.restype = i32
        lib.dfit_device_count.restype = i32
        lib.dfit_strerror.restype = ctypes.c_char_p
        lib.dfit_strerror.argtypes = [i32]
        lib.dfit_last_error.restype = ctypes.c_char_p
        lib.dfit_default_opts.argtypes = [ctypes.POINTER(DfitOpts), i32]
        lib.dfit_model_nparams.argtypes = [i32]
        lib.dfit_create.argtypes = [i32, ctypes.POINTER(vp)]
        lib.dfit_destroy.argtypes = [vp]
        lib.dfit_fit_device.argtypes = [vp, ctypes.POINTER(DfitOpts), i32, i64, vp, vp, i32, i32, i64, vp, vp, i32,
                                        vp, vp, i32, vp, vp, vp]
        lib.dfit_fit_host.argtypes = [vp, ctypes.POINTER(DfitOpts), i32, i64, vp, vp, i32, vp, vp, i32, vp, vp, i32,
                                      vp, vp]
        lib.dfit_get_stats.argtypes = [vp, ctypes.POINTER(DfitStats)]
        lib.dfit_set_gather.argtypes = [vp, i32, i32, vp, i64]
        lib.dfit_ipc_alloc.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp), ctypes.c_char_p]
        lib.dfit_ipc_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
        lib.dfit_ipc_close.argtypes = [vp, vp]
        lib.dfit_ipc_free.argtypes = [vp, vp]
        lib.dfit_default_qdess_opts.argtypes = [ctypes.POINTER(DfitQdessOpts)]
        lib.dfit_qdess_t2_device.argtypes = [vp, ctypes.POINTER(DfitQdessOpts), i64, vp, vp, i32, vp, i32, vp]
        lib.dfit_qdess_t2_host.argtypes = [vp, ctypes.POINTER(DfitQdessOpts), i64, vp, vp, i32, vp, i32]
        lib.dfit_region_metrics_host.argtypes = [vp, i64, vp, i32, vp, i32, i32, vp, i32, ctypes.c_double,
                                                 ctypes.c_double, i32, i32, vp]
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        lib = load()
        detail = lib.dfit_last_error().decode() or lib.dfit_strerror(rc).decode()
        raise DfitError(rc, detail)


def default_opts(model):
    o = DfitOpts()
    check(load().dfit_default_opts(ctypes.byref(o), model))
    return o


class Handle:
    """RAII wrapper around a dfit_handle (one per thread and device; created lazily on first fit so
    that importing this package never creates a CUDA context -- users may fork first)."""

    def __init__(self, device=0):
        lib = load()
        h = ctypes.c_void_p()
        check(lib.dfit_create(device, ctypes.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load().dfit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ptr(self):
        return self._h

    def stats(self):
        s = DfitStats()
        check(load().dfit_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()


_tls = threading.local()


def get_handle(device=0):
    cache = getattr(_tls, "handles", None)
    if cache is None:
        cache = _tls.handles = {}
    key = (os.getpid(), device)
    if key not in cache:
        cache[key] = Handle(device)
    return cache[key]
