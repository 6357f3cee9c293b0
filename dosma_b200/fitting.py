"""Drop-in replacements for `dosma.core.fitting.{curve_fit, CurveFitter, MonoExponentialFit}`.

Same constructor / `fit()` signatures, argument meaning, return types and error behaviour as the
reference (`dosma/core/fitting.py` @ bd5efec; line numbers cited per function), but the N-voxel
loop that calls `scipy.optimize.curve_fit` once per voxel (fitting.py:855-868, 1026-1073) is ONE
call into the hand-written sm_100a CUDA engine behind the C-ABI of `include/dfit.h`.  Everything
the reference does around that loop -- dtype up-cast, mask select + scatter, the log-linear
("polyfit") initial guess, `_process_params`, rounding -- is fused into the same launch.

There is no CPU fallback: without `libdfit.so` and a CUDA device these functions raise.
"""
import ctypes
import inspect
import os
import warnings
from collections.abc import Mapping, Sequence
from copy import deepcopy
from numbers import Number

import numpy as np

from . import _cabi
from .med_volume import MedicalVolume, is_volume
from .models import biexponential, inv_abs, monoexponential, param_names, resolve_model, resolve_ufunc

__all__ = ["CurveFitter", "MonoExponentialFit", "curve_fit", "monoexponential", "biexponential"]

_R2_THRESHOLD_TEMPLATE = 0.9  # dosma/resources/templates/.preferences.yml:3-4 (fitting/r2.threshold)
_AFFINE_DECIMAL_PRECISION = 4  # dosma/defaults.py AFFINE_DECIMAL_PRECISION, used at fitting.py:104
_ENGINE_KWARGS = ("compute_dtype", "device", "xtol", "lambda0", "ftol_scale", "init_linear", "use_tma",
                  "fast_path", "out_dtype", "return_stats")

_default_compute_dtype = "auto"


def set_default_compute_dtype(value):
    """'auto' (float64 samples -> fp64 arithmetic, everything else fp32), 'f32' or 'f64'."""
    global _default_compute_dtype
    if value not in ("auto", "f32", "f64"):
        raise ValueError("compute dtype must be 'auto', 'f32' or 'f64'")
    _default_compute_dtype = value


def _preferences_r2_threshold():
    """`r2_threshold="preferences"` (fitting.py:85-93): DOSMA's preference if DOSMA is importable,
    else the value its template ships with."""
    try:  # pragma: no cover - dosma is not installed on the build/GPU boxes
        from dosma.defaults import preferences

        return preferences.fitting_r2_threshold
    except Exception:
        return _R2_THRESHOLD_TEMPLATE


def _default_device():
    for key in ("DOSMA_B200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


# ------------------------------------------------------------------------------------------------
# engine call
# ------------------------------------------------------------------------------------------------
def _as_plane(arr):
    """1-D contiguous array of a dtype the kernels convert in-register (include/dfit.h dfit_dtype)."""
    arr = np.asarray(arr).reshape(-1)
    if arr.dtype not in _cabi.NP_TO_DTYPE:
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        elif np.issubdtype(arr.dtype, np.integer) and arr.dtype.itemsize <= 2:
            arr = arr.astype(np.int32)
        elif arr.dtype == np.float16:
            arr = arr.astype(np.float32)
        elif np.issubdtype(arr.dtype, np.integer) or np.issubdtype(arr.dtype, np.floating):
            arr = arr.astype(np.float64)
        else:
            raise TypeError(f"Unsupported sample dtype {arr.dtype}")
    return np.ascontiguousarray(arr)


def _engine_fit(model_id, nparams, x, planes, mask, p0_cols, *, init_mode=_cabi.INIT_GIVEN, y_bounds=None,
                maxfev=100, ftol=1e-5, eps=1e-8, post=None, engine=None, out_param=None):
    """Run the CUDA engine on host buffers.

    planes: list of E arrays (one per echo), each flattened to [N]; mask: None or [N] (truthy =
    fit); p0_cols: length-P list of scalars or [N] arrays.  post: dict for the fused epilogue.
    out_param: None, or the index of the one parameter to return (popt is then (N,): `dfit_opts.out_param`).
    Returns popt (N, P) float64, r2 (N,) float64, stats dict.
    """
    engine = dict(engine or {})
    lib = _cabi.load()
    planes = [_as_plane(p) for p in planes]
    dt = np.result_type(*[p.dtype for p in planes])
    planes = [p if p.dtype == dt else _as_plane(p.astype(dt)) for p in planes]
    N = planes[0].shape[0]
    if any(p.shape[0] != N for p in planes):
        raise ValueError("All echoes must have the same number of voxels")
    E = len(planes)
    if E > _cabi.MAX_ECHOES:
        raise NotImplementedError(f"{E} echoes > {_cabi.MAX_ECHOES} supported by the CUDA engine")
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    if x.shape[0] != E:
        raise ValueError(f"Dimension mismatch: len(x)={x.shape[0]}, but {E} echoes")

    o = _cabi.default_opts(model_id)
    cd = engine.get("compute_dtype") or _default_compute_dtype
    if cd == "auto":
        cd = "f64" if dt == np.float64 else "f32"
    if cd not in ("f32", "f64"):
        raise ValueError("compute_dtype must be 'auto', 'f32' or 'f64'")
    o.compute_dtype = _cabi.F64 if cd == "f64" else _cabi.F32
    o.init_mode = init_mode
    o.maxfev = int(maxfev)
    o.ftol = float(ftol)
    o.r2_eps = float(eps)
    for k in ("xtol", "lambda0", "ftol_scale"):
        if engine.get(k) is not None:
            setattr(o, k, float(engine[k]))
    for k in ("init_linear", "use_tma", "fast_path"):
        if engine.get(k) is not None:
            setattr(o, k, int(engine[k]))
    if y_bounds is not None:
        o.y_lo, o.y_hi = float(y_bounds[0]), float(y_bounds[1])

    per_voxel = [i for i, c in enumerate(p0_cols) if isinstance(c, np.ndarray)]
    p0v = None
    for i, c in enumerate(p0_cols):
        o.p0[i] = float("nan") if i in per_voxel else float(c)
    if per_voxel and init_mode == _cabi.INIT_GIVEN:
        p0v = np.empty((N, nparams), dtype=np.float64 if cd == "f64" else np.float32)
        for i, c in enumerate(p0_cols):
            p0v[:, i] = c if i in per_voxel else 0
    elif init_mode != _cabi.INIT_GIVEN:
        for i in range(nparams):
            o.p0[i] = 1.0

    if post:
        o.post_enabled = 1
        for i in range(nparams):
            o.ufunc[i] = post["ufunc"][i]
            o.lb[i], o.ub[i] = post["lb"][i], post["ub"][i]
            o.decimals[i] = post["decimals"][i]
        if post.get("r2_threshold") is not None:
            o.has_r2_threshold, o.r2_threshold = 1, float(post["r2_threshold"])
        if post.get("nan_to_num") is not None:
            o.has_nan_fill, o.nan_fill = 1, float(post["nan_to_num"])

    mask_u8 = None
    if mask is not None:
        mask_u8 = np.ascontiguousarray(np.asarray(mask).reshape(-1) != 0).view(np.uint8)
        if mask_u8.shape[0] != N:
            raise ValueError("mask size mismatch")

    # float64 like the reference's results (fitting.py:870) unless the caller opts into float32 maps
    # (out_dtype="f32": half the device-to-host bytes; fp32 arithmetic carries no more information anyway)
    out_dt = {"f64": np.float64, "f32": np.float32}.get(engine.get("out_dtype") or "f64")
    if out_dt is None:
        raise ValueError("out_dtype must be 'f64' or 'f32'")
    if out_param is not None:
        o.out_param = int(out_param)
    popt = np.empty((N, nparams) if out_param is None else (N,), dtype=out_dt)
    r2 = np.empty(N, dtype=out_dt)
    plane_ptrs = (ctypes.c_void_p * E)(*[p.ctypes.data for p in planes])
    h = _cabi.get_handle(engine.get("device", _default_device()))
    _cabi.check(lib.dfit_fit_host(
        h.ptr, ctypes.byref(o), E, N, x.ctypes.data, ctypes.cast(plane_ptrs, ctypes.c_void_p),
        _cabi.NP_TO_DTYPE[planes[0].dtype], mask_u8.ctypes.data if mask_u8 is not None else None,
        p0v.ctypes.data if p0v is not None else None, _cabi.F64 if cd == "f64" else _cabi.F32,
        popt.ctypes.data, r2.ctypes.data, _cabi.F64 if out_dt == np.float64 else _cabi.F32, None, None))
    stats = h.stats()
    if stats["n_nonfinite"]:
        # SciPy's asarray_chkfinite aborts the whole reference fit the same way
        raise ValueError("array must not contain infs or NaNs")
    return popt, r2, stats


# ------------------------------------------------------------------------------------------------
# p0 handling
# ------------------------------------------------------------------------------------------------
def _split_p0(p0, names, N):
    """Normalise an initial guess to a length-P list of scalars / per-voxel [N] arrays.

    Accepts what the reference accepts (fitting.py:1106-1161): None, a Number (broadcast), an (N, P)
    ndarray (split by column), a length-P sequence or a (possibly partial) mapping whose entries are
    numbers, None (-> 1.0) or [N] arrays.  Same ValueErrors for wrong length / unknown keys.
    """
    P = len(names)
    if p0 is None:
        return [1.0] * P
    if isinstance(p0, Number):
        return [float(p0)] * P
    if isinstance(p0, np.ndarray) and p0.ndim > 1:
        p0 = tuple(p0[..., i] for i in range(p0.shape[-1]))
    if isinstance(p0, Mapping):
        extra = set(p0) - set(names)
        if extra:
            raise ValueError(f"`p0` has unknown keys: {extra}. Function signature has parameters {names}.")
        p0 = [p0.get(k, 1.0) for k in names]
    elif isinstance(p0, (np.ndarray, Sequence)):
        if len(p0) != P:
            raise ValueError(f"`p0` has length {len(p0)} but function has {P} parameters")
        p0 = list(p0)
    else:
        raise ValueError(f"p0={p0} not supported")
    cols = []
    for k, v in zip(names, p0):
        if v is None:
            v = 1.0
        if isinstance(v, np.ndarray) and v.ndim > 0:
            if len(v) != N:
                raise ValueError(f"Got {len(v)} values for param '{k}'. Expected {N}")
            v = np.asarray(v, dtype=np.float64).reshape(-1)
        else:
            v = float(v)
        cols.append(v)
    return cols


# ------------------------------------------------------------------------------------------------
# curve_fit
# ------------------------------------------------------------------------------------------------
def curve_fit(func, x, y, y_bounds=None, p0=None, maxfev=100, ftol=1e-5, eps=1e-8, show_pbar=False,
              num_workers=0, chunksize=None, **kwargs):
    """Fit ``func`` to N independent sequences -- signature of `dosma.curve_fit` (fitting.py:755-768).

    Args:
        func: model function ``f(x, *params)``; must be one the engine implements (see
            :mod:`dosma_b200.models`), recognised by identity or numeric fingerprint.
        x: (E,) independent variable.  y: (E,) or (E, N) dependent data, any real dtype.
        y_bounds: sequences with a sample outside are not fit -> NaN, r2 = 0 (fitting.py:1065-1067).
        p0: None | Number | length-P sequence | (N, P) ndarray | mapping (fitting.py:1106-1161).
        maxfev, ftol, eps: as in the reference (defaults 100, 1e-5, 1e-8).
        show_pbar, num_workers, chunksize: accepted for compatibility; the GPU fits all N sequences
            in one launch, so they have no effect.
        kwargs: engine options (``compute_dtype``, ``device``, ``xtol``, ``lambda0``, ``fast_path``,
            ``out_dtype`` ...).
            ``fast_path=0`` forces the Levenberg-Marquardt iteration from ``p0`` for every sequence; by
            default mono-exponential sequences are fitted by a variable-projection Newton iteration from
            a data-driven start (same minimiser; ``p0`` then only serves the sequences that path
            declines, e.g. low-SNR ones).  SciPy options that change the algorithm (``bounds``,
            ``method``, ``sigma``, ``jac``) are not supported and raise NotImplementedError.

    Returns:
        popts (N, P) float64 and r2 (N,) float64, NaN / 0 where the fit failed (fitting.py:870).
    """
    model_id, nparams = resolve_model(func)
    engine = {k: kwargs.pop(k) for k in list(kwargs) if k in _ENGINE_KWARGS}
    if kwargs:
        bad = sorted(kwargs)
        if any(k in ("bounds", "method", "sigma", "absolute_sigma", "jac", "max_nfev", "loss") for k in bad):
            raise NotImplementedError(f"SciPy options {bad} are not supported by the CUDA engine")
        raise TypeError(f"curve_fit() got unexpected keyword arguments {bad}")
    want_stats = engine.pop("return_stats", False)

    x = np.asarray(x)
    y = np.asarray(y)
    if y.ndim == 1:
        y = y.reshape(y.shape + (1,))
    if y.ndim != 2:
        raise ValueError("`y` must have shape (M,) or (M, N)")
    N = y.shape[-1]
    cols = _split_p0(p0, param_names(func), N)

    if y_bounds is not None and ((y < y_bounds[0]).any() or (y > y_bounds[1]).any()):
        warnings.warn("Out of bounds values found. Failure in fit will result in np.nan")  # fitting.py:845-847

    popt, r2, stats = _engine_fit(model_id, nparams, x, [y[e] for e in range(y.shape[0])], None, cols,
                                  y_bounds=y_bounds, maxfev=maxfev, ftol=ftol, eps=eps, engine=engine)
    return (popt, r2, stats) if want_stats else (popt, r2)


# ------------------------------------------------------------------------------------------------
# _Fitter machinery (MedicalVolume marshalling + post-processing)
# ------------------------------------------------------------------------------------------------
def _check_ufuncs(out_ufuncs, nparams):
    """fitting.py:58-75."""
    if not callable(out_ufuncs) and not all(callable(u) or u is None for u in out_ufuncs):
        raise TypeError(f"`out_ufuncs` must be callable or sequence of callables. Got {out_ufuncs}")
    if isinstance(out_ufuncs, Sequence) and len(out_ufuncs) > nparams:
        warnings.warn(
            f"len(out_ufuncs)={len(out_ufuncs)}, but only {nparams} parameters. Extra ufuncs will be ignored."
        )
    return out_ufuncs


def _check_bounds(out_bounds):
    """fitting.py:77-83."""
    b = np.asarray(out_bounds)
    if b.shape[-1] != 2 or b.ndim > 2:
        raise ValueError("Invalid `out_bounds` - shape must be ([num_params,] 2)")
    if np.any(b[..., 0] > b[..., 1]):
        raise ValueError("Invalid `out_bounds` - lower bound must be <= upper bound")
    return b


def _check_r2_threshold(value):
    """fitting.py:85-93."""
    if isinstance(value, str):
        if value != "preferences":
            raise ValueError(
                f"Invalid value r2_threshold='{value}'. Expected `None`, a number between [0, 1], or 'preferences'."
            )
        value = _preferences_r2_threshold()
    return value


def _bounds_per_param(out_bounds, nparams):
    """1-D bounds apply to every parameter; 2-D are padded with (-inf, inf) (fitting.py:130-137)."""
    lb = [-np.inf] * nparams
    ub = [np.inf] * nparams
    if out_bounds is not None:
        b = np.asarray(out_bounds, dtype=np.float64)
        if b.ndim == 1:
            lb, ub = [b[0]] * nparams, [b[1]] * nparams
        else:
            for i in range(min(nparams, b.shape[0])):
                lb[i], ub[i] = b[i, 0], b[i, 1]
    return lb, ub


def _process_params_host(x, r2, out_ufuncs, out_bounds, r2_threshold, nan_to_num):
    """Host restatement of `_process_params` (fitting.py:109-146), used only when a post-processing
    callable is not one the fused epilogue recognises (arbitrary Python cannot run in the kernel)."""
    nparams = x.shape[-1]
    if callable(out_ufuncs):
        x = out_ufuncs(x)
    elif out_ufuncs is not None:
        for i in range(min(nparams, len(out_ufuncs))):
            if out_ufuncs[i] is not None:
                x[..., i] = out_ufuncs[i](x[..., i])
    if out_bounds is not None:
        lb, ub = _bounds_per_param(out_bounds, nparams)
        x[(x < np.asarray(lb)) | (x > np.asarray(ub))] = np.nan
    if r2_threshold is not None:
        x[r2 < r2_threshold] = np.nan
    if nan_to_num is not None:
        x = np.nan_to_num(x, nan=nan_to_num, copy=False)
    return x


def _mask_to_volume(mask, y0):
    """fitting.py:95-107: ndarray -> volume with y's affine; reorient; dimension check; `> 0`."""
    if isinstance(mask, np.ndarray):
        mask = y0._partial_clone(volume=mask, headers=None)
    elif not is_volume(mask):
        raise TypeError("`mask` must be a MedicalVolume or ndarray")
    mask = mask.reformat_as(y0)
    if not mask.is_same_dimensions(y0, _AFFINE_DECIMAL_PRECISION):
        raise RuntimeError("`mask` and `y` dimension mismatch")
    return mask


def _on_cpu(obj):
    dev = getattr(obj, "device", "cpu")
    return str(dev).lower() in ("cpu", "device(type='cpu')") or getattr(dev, "type", None) == "cpu" or \
        getattr(dev, "id", 0) == -1


def _format_p0_volumes(p0, ref, mask_flat, depth=0):
    """Bring MedicalVolume / ndarray initial guesses to flat (masked-order irrelevant here: the mask
    is applied in-kernel) [N] arrays -- fitting.py:344-380."""
    if p0 is None or isinstance(p0, Number):
        return p0
    if is_volume(p0) and depth > 0:
        p0 = p0.reformat_as(ref)
        p0.is_same_dimensions(ref, err=True)
        return np.asarray(p0.volume).reshape(-1)
    if isinstance(p0, np.ndarray) and depth > 0:
        if p0.shape != tuple(ref.shape):
            raise ValueError(f"Got p0.shape={p0.shape}, but y.shape={ref.shape}")
        return p0.reshape(-1)
    if isinstance(p0, Mapping):
        return {k: _format_p0_volumes(v, ref, mask_flat, depth + 1) for k, v in p0.items()}
    if isinstance(p0, Sequence):
        return tuple(_format_p0_volumes(v, ref, mask_flat, depth + 1) for v in p0)
    if isinstance(p0, np.ndarray) or is_volume(p0):
        arr = np.asarray(p0.volume) if is_volume(p0) else p0
        return tuple(_format_p0_volumes(arr[..., i], ref, mask_flat, depth + 1) for i in range(arr.shape[-1]))
    raise ValueError(f"p0={p0} not supported")


class CurveFitter:
    """Non-linear least squares over MedicalVolumes -- `dosma.core.fitting.CurveFitter`
    (fitting.py:238-458) on the CUDA engine.

    Constructor and `fit()` arguments, validation errors and warnings are the reference's
    (fitting.py:304-342, 382-420).  ``num_workers``, ``chunksize`` and ``verbose`` are accepted
    and have no effect.  ``kwargs`` may carry engine options (``compute_dtype``, ``device``...).
    """

    def __init__(self, func, p0=None, y_bounds=None, out_ufuncs=None, out_bounds=None, r2_threshold="preferences",
                 nan_to_num=None, num_workers=0, chunksize=None, verbose=False, **kwargs):
        self._func = func
        self._func_name = func.__name__ if hasattr(func, "__name__") else type(func).__name__
        self._param_names = param_names(func)
        nparams = len(self._param_names)
        if out_ufuncs is not None:
            out_ufuncs = _check_ufuncs(out_ufuncs, nparams)
        if out_bounds is not None:
            out_bounds = _check_bounds(out_bounds)
        self.p0 = p0
        self.y_bounds = y_bounds
        self.out_ufuncs = out_ufuncs
        self.out_bounds = out_bounds
        self.r2_threshold = _check_r2_threshold(r2_threshold)
        self.nan_to_num = nan_to_num
        self.num_workers = num_workers
        self.chunksize = chunksize
        self.verbose = verbose
        self.kwargs = kwargs
        self._decimals = None  # set by MonoExponentialFit to fuse its rounding
        self._out_param = None  # set by MonoExponentialFit: the one parameter it keeps (fitting.py:734)

    # -- epilogue planning -----------------------------------------------------------------------
    def _plan_post(self, nparams):
        """Fused epilogue description, or None when a ufunc must run on the host."""
        ids = [_cabi.UFUNC_NONE] * nparams
        uf = self.out_ufuncs
        if callable(uf):
            return None
        if uf is not None:
            for i in range(min(nparams, len(uf))):
                uid = resolve_ufunc(uf[i])
                if uid is None:
                    return None
                ids[i] = uid
        lb, ub = _bounds_per_param(self.out_bounds, nparams)
        dec = [-1] * nparams
        if self._decimals:
            for i, d in self._decimals.items():
                dec[i] = int(d)
        return {"ufunc": ids, "lb": lb, "ub": ub, "decimals": dec, "r2_threshold": self.r2_threshold,
                "nan_to_num": self.nan_to_num}

    def fit(self, x, y, mask=None, p0=np._NoValue, copy_headers=True, _init_mode=_cabi.INIT_GIVEN):
        """Fit every voxel; returns (popt volume with a trailing parameter axis, r2 volume)
        (fitting.py:382-420 -> :157-235)."""
        if not _on_cpu(x):
            raise RuntimeError("`x` must be on the CPU")
        if (not isinstance(y, (list, tuple))) or (not all(is_volume(_y) for _y in y)):
            raise TypeError("`y` must be sequence of MedicalVolumes.")
        if any(not _on_cpu(_y) for _y in y):
            raise RuntimeError("All elements in `y` must be on the CPU")
        x = np.asarray(x)
        if x.shape[-1] != len(y):
            raise ValueError("Dimension mismatch: x.shape[-1]={:d}, but len(y)={:d}".format(x.shape[-1], len(y)))

        model_id, nparams = resolve_model(self._func)
        orientation = y[0].orientation
        y = [_y.reformat(orientation) for _y in y]
        y0 = y[0]
        mask_flat = None
        if mask is not None:
            mask_flat = (np.asarray(_mask_to_volume(mask, y0).volume) > 0).reshape(-1)

        if p0 is np._NoValue:
            p0 = self.p0
        p0 = _format_p0_volumes(p0, y0, mask_flat)
        original_shape = tuple(y0.shape)
        N = int(np.prod(original_shape))
        cols = _split_p0(p0, self._param_names, N)

        planes = [np.asarray(_y.volume).reshape(-1) for _y in y]
        if self.y_bounds is not None:
            # (the reference tests the samples `curve_fit` receives, i.e. the masked ones: fitting.py:199-200, 845-847)
            sel = slice(None) if mask_flat is None else mask_flat
            if any((p[sel] < self.y_bounds[0]).any() or (p[sel] > self.y_bounds[1]).any() for p in planes):
                warnings.warn("Out of bounds values found. Failure in fit will result in np.nan")

        post = self._plan_post(nparams)
        engine = {k: v for k, v in self.kwargs.items() if k in _ENGINE_KWARGS and k != "return_stats"}
        fit_kwargs = {k: self.kwargs[k] for k in ("maxfev", "ftol", "eps") if k in self.kwargs}
        unknown = set(self.kwargs) - set(_ENGINE_KWARGS) - {"maxfev", "ftol", "eps"}
        if unknown:
            raise NotImplementedError(f"curve_fit options {sorted(unknown)} are not supported by the CUDA engine")
        fused = post if post is not None else {
            "ufunc": [0] * nparams, "lb": [-np.inf] * nparams, "ub": [np.inf] * nparams, "decimals": [-1] * nparams,
            "r2_threshold": None, "nan_to_num": None}
        # (a single kept parameter is selected on the device when the whole epilogue is fused: a third less to write
        # and to bring back over PCIe, and no host pass to slice the column out)
        only = self._out_param if post is not None else None
        popt, r2, stats = _engine_fit(model_id, nparams, x, planes, mask_flat, cols, init_mode=_init_mode,
                                      y_bounds=self.y_bounds, post=fused, engine=engine, out_param=only, **fit_kwargs)
        self.last_stats = stats
        if post is None:
            # arbitrary Python ufunc: finish `_process_params` on the host, then redo the mask fill
            sel = slice(None) if mask_flat is None else mask_flat
            with np.errstate(all="ignore"):
                popt[sel] = _process_params_host(popt[sel], r2[sel], self.out_ufuncs, self.out_bounds,
                                                 self.r2_threshold, self.nan_to_num)
            if mask_flat is not None:
                fill = np.nan if self.nan_to_num is None else self.nan_to_num
                popt[~mask_flat] = fill
                r2[~mask_flat] = fill
            if self._decimals:
                for i, d in self._decimals.items():
                    popt[:, i] = np.around(popt[:, i], d)

        if self._out_param is not None and only is None:  # (host epilogue: the column is sliced here)
            popt = np.ascontiguousarray(popt[:, self._out_param])
        one = self._out_param is not None
        popt = popt.reshape(original_shape if one else original_shape + (nparams,))
        r2 = r2.reshape(original_shape)
        if copy_headers:
            headers = y0.headers()
            if headers is not None:
                # (parameter maps get a trailing axis, fitting.py:503-507; a single parameter -- what `popt[..., i]`
                # leaves, :734 -- keeps the shape of the inputs' headers)
                headers = deepcopy(headers) if one else np.expand_dims(deepcopy(headers), axis=-1)
            popt_headers, r2_headers = headers, True
        else:
            popt_headers, r2_headers = None, None
        return y0._partial_clone(volume=popt, headers=popt_headers), y0._partial_clone(volume=r2, headers=r2_headers)

    def __str__(self):
        attrs = ["p0", "y_bounds", "out_bounds", "r2_threshold", "nan_to_num", "num_workers", "chunksize", "verbose"]
        vals = [f"func={self._func_name}"] + [f"{k}={getattr(self, k)}" for k in attrs]
        vals += [f"{k}={v}" for k, v in self.kwargs.items()]
        return f"{self.__class__.__name__}(\n\t" + "\n\t".join(v + "," for v in vals) + "\n)"


class MonoExponentialFit:
    """Mono-exponential time-constant map (T2, T1rho, T2*) -- `dosma.core.fitting.MonoExponentialFit`
    (fitting.py:606-749).  One fused launch: optional in-kernel log-linear initial guess
    (``tc0="polyfit"``, fitting.py:701-718), LM fit, ``tc = 1/|b|`` (:725), bounds on tc (:726),
    r2 threshold, NaN -> 0 (:731) and rounding to ``decimal_precision`` (:736-737)."""

    def __init__(self, x=None, y=None, mask=None, bounds=(0, 100.0), tc0=30.0, r2_threshold="preferences",
                 decimal_precision=1, num_workers=0, chunksize=1000, verbose=False, **engine_kwargs):
        self.x = x
        if y is not None:
            warnings.warn(
                f"Setting `y` in the constructor can result in significant memory overhead. "
                f"Specify `y` in `{type(self).__name__}.fit(y=...)` instead."
            )
            self._check_y(x, y)
        self.y = y
        if mask is not None:
            warnings.warn(
                f"Setting `mask` in the constructor can result in significant memory overhead. "
                f"Specify `mask` in `{type(self).__name__}.fit(mask=...)` instead."
            )
        self.mask = mask
        if not (isinstance(tc0, Number) or (isinstance(tc0, str) and tc0 == "polyfit")):
            raise ValueError("`tc0` must either be a float or the string 'polyfit'.")
        self.verbose = verbose
        self.num_workers = num_workers
        if len(bounds) != 2:
            raise ValueError("`bounds` should provide lower/upper bound in format (lb, ub)")
        self.bounds = bounds
        self.chunksize = chunksize
        self.r2_threshold = r2_threshold
        self.tc0 = tc0
        self.decimal_precision = decimal_precision
        self._engine_kwargs = engine_kwargs

    def fit(self, x=None, y=None, mask=None):
        """Returns (time-constant volume, r2 volume) (fitting.py:678-739)."""
        x = self.x if x is None else x
        y = self.y if y is None else y
        mask = self.mask if mask is None else mask
        self._check_y(x, y)
        if isinstance(mask, np.ndarray):
            mask = MedicalVolume(mask, affine=y[0].affine) if isinstance(y[0], MedicalVolume) else \
                y[0]._partial_clone(volume=mask, headers=None)

        polyfit = isinstance(self.tc0, str)
        fitter = CurveFitter(
            monoexponential,
            y_bounds=None,
            out_ufuncs=(None, inv_abs),  # 1 / |b| (fitting.py:725)
            out_bounds=((-np.inf, np.inf), self.bounds),
            r2_threshold=self.r2_threshold,
            num_workers=self.num_workers,
            chunksize=self.chunksize,
            verbose=self.verbose,
            nan_to_num=0.0,
            **self._engine_kwargs,
        )
        if self.decimal_precision is not None:
            fitter._decimals = {1: self.decimal_precision}
        fitter._out_param = 1  # `popt[..., 1]` (fitting.py:734), selected on the device
        p0 = None if polyfit else {"a": 1.0, "b": -1 / self.tc0}
        popt, r_squared = fitter.fit(
            x, y, mask=mask, p0=p0, _init_mode=_cabi.INIT_LOGLINEAR if polyfit else _cabi.INIT_GIVEN)
        self.last_stats = fitter.last_stats
        return popt, r_squared

    def _check_y(self, x, y):
        """fitting.py:741-749."""
        if (not isinstance(y, Sequence)) or (not all(is_volume(sv) for sv in y)):
            raise TypeError("`y` must be list of MedicalVolumes.")
        if any(not _on_cpu(sv) for sv in y):
            raise RuntimeError("All MedicalVolumes must be on the CPU")
        if len(x) != len(y):
            raise ValueError("`len(x)`={:d}, but `len(y)`={:d}".format(len(x), len(y)))
