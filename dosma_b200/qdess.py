"""qDESS analytic T2 map on the GPU -- the element-wise step of `QDess.generate_t2_map`
(`dosma/scan_sequences/mri/qdess.py:105-263`), SURVEY.md section 8 row f3.

`qdess_t2_map` takes the two echo volumes and the sequence parameters the reference reads from
the DICOM header (or from its keyword arguments) and returns the T2 map with the reference's
post-processing (bounds -> NaN, NaN fill, rounding, optional fat / fluid suppression).  The scalar
constants `k` and `c1` are computed here exactly as at qdess.py:204-223; the per-voxel arithmetic
(:225-255) runs in `dfit_qdess_t2_host` (include/dfit.h).
"""
import ctypes
import math
import numbers
import warnings

import numpy as np

from . import _cabi
from .fitting import _as_plane, _default_device
from .med_volume import is_volume

__all__ = ["qdess_constants", "qdess_t2_map"]

_GAMMA = 4258 * 2 * math.pi  # rad / (G s), qdess.py:214


def qdess_constants(tr, te, tg, gl_area, alpha, t1, diffusivity=1.25e-9):
    """(k, c1, TR - TE [s]) from the sequence parameters -- qdess.py:193-223.

    tr, te, t1 in ms; tg in microseconds; alpha in degrees; gl_area as in DICOM tag 0x001910b6.
    """
    TR, TE, Tg, T1 = tr * 1e-3, te * 1e-3, tg * 1e-6, t1 * 1e-3
    a = math.radians(alpha)
    if np.allclose(math.sin(a / 2), 0):
        warnings.warn("sin(flip angle) is close to 0 - t2 map may fail.")  # qdess.py:208-209
    gl = gl_area / (Tg * 1e6) * 100
    dkl = _GAMMA * gl * Tg
    decay = math.exp(-TR / T1 - TR * dkl ** 2 * diffusivity)
    k = math.sin(a / 2) ** 2 * (1 + decay) / (1 - math.cos(a) * decay)
    c1 = (TR - Tg / 3) * dkl ** 2 * diffusivity
    return k, c1, TR - TE


def qdess_t2_map(echo1, echo2, *, tr, te, tg, gl_area, alpha, t1, diffusivity=1.25e-9, nan_bounds=(0, 100),
                 nan_to_num=0.0, decimals=1, suppress_fat=False, suppress_fluid=False, beta=1.2, device=None,
                 precision="exact"):
    """T2 map from the two qDESS echoes (ndarrays or volumes of equal shape).

    Keyword arguments and defaults are those of `QDess.generate_t2_map` (qdess.py:105-121).
    Returns an ndarray, or a volume made with `echo1._partial_clone` when volumes are passed
    (qdess.py:257).  precision="exact" (default) uses the reference's float64 arithmetic and returns
    float64 like the reference (its `mask = ones(...)` promotes every volume to float64, :226-228);
    precision="fast" computes and returns float32 (HBM-bound kernel, ~2e-6 relative before rounding).
    """
    if precision not in ("exact", "fast"):
        raise ValueError("precision must be 'exact' or 'fast'")
    if not all(isinstance(v, numbers.Real) for v in (alpha, t1, diffusivity)):  # (numpy scalars included)
        raise NotImplementedError("array-valued alpha / t1 / diffusivity are not supported by the CUDA path")
    vol = echo1 if is_volume(echo1) else None
    a1 = np.asarray(echo1.volume if is_volume(echo1) else echo1)
    a2 = np.asarray(echo2.volume if is_volume(echo2) else echo2)
    if a1.shape != a2.shape:
        raise ValueError("echo volumes must have the same shape")
    p1, p2 = _as_plane(a1), _as_plane(a2)
    if p1.dtype != p2.dtype:
        dt = np.result_type(p1.dtype, p2.dtype)
        p1, p2 = _as_plane(p1.astype(dt)), _as_plane(p2.astype(dt))
    k, c1, dt_s = qdess_constants(tr, te, tg, gl_area, alpha, t1, diffusivity)
    lib = _cabi.load()
    o = _cabi.DfitQdessOpts()
    _cabi.check(lib.dfit_default_qdess_opts(ctypes.byref(o)))
    o.k, o.c1, o.tr_minus_te = k, c1, dt_s
    if nan_bounds is None:
        o.has_bounds = 0
    else:
        o.has_bounds, o.lb, o.ub = 1, float(nan_bounds[0]), float(nan_bounds[1])
    if nan_to_num is None:
        o.has_nan_fill = 0
    else:
        o.has_nan_fill = 1
        o.nan_fill = 0.0 if isinstance(nan_to_num, bool) else float(nan_to_num)  # qdess.py:241-245
    o.decimals = -1 if decimals is None else int(decimals)
    o.suppress_fat, o.suppress_fluid, o.beta = int(bool(suppress_fat)), int(bool(suppress_fluid)), float(beta)
    fast = precision == "fast"
    if fast and p1.dtype != np.float32:
        p1, p2 = p1.astype(np.float32), p2.astype(np.float32)
    o.compute_dtype = _cabi.F32 if fast else _cabi.F64
    out_dtype = np.float32 if fast else np.float64
    out = np.empty(p1.shape[0], dtype=out_dtype)
    h = _cabi.get_handle(_default_device() if device is None else device)
    _cabi.check(lib.dfit_qdess_t2_host(h.ptr, ctypes.byref(o), p1.shape[0], p1.ctypes.data, p2.ctypes.data,
                                       _cabi.NP_TO_DTYPE[p1.dtype], out.ctypes.data,
                                       _cabi.F32 if out_dtype == np.float32 else _cabi.F64))
    out = out.reshape(a1.shape)
    return vol._partial_clone(volume=out, headers=True) if vol is not None else out
