"""HBM-resident entry point: fit samples that already live in device memory (torch CUDA tensors).

torch is used here for what the task brief calls plumbing -- device allocations, streams,
`torch.distributed` -- never for arithmetic: the fit is `dfit_fit_device` (include/dfit.h) launched
on torch's current stream, so `torch.cuda.Event` timings around a call time exactly the kernel.
"""
import ctypes

import numpy as np

from . import _cabi
from .models import resolve_model

_TORCH_TO_DTYPE = None


def _torch_dtypes():
    global _TORCH_TO_DTYPE
    if _TORCH_TO_DTYPE is None:
        import torch

        _TORCH_TO_DTYPE = {torch.float32: _cabi.F32, torch.float64: _cabi.F64, torch.int16: _cabi.I16,
                           torch.int32: _cabi.I32, torch.uint8: _cabi.U8}
        if hasattr(torch, "uint16"):
            _TORCH_TO_DTYPE[torch.uint16] = _cabi.U16
    return _TORCH_TO_DTYPE


def make_opts(func_or_model, *, p0=None, compute_dtype="f32", init="given", maxfev=100, ftol=1e-5, eps=1e-8,
              y_bounds=None, post=None, **engine):
    """Build a DfitOpts for `fit_device`.  `post` = dict(ufunc=[...], lb=[...], ub=[...],
    decimals=[...], r2_threshold=..., nan_to_num=...) enables the fused epilogue."""
    model_id, P = (func_or_model, _cabi.load().dfit_model_nparams(func_or_model)) if isinstance(
        func_or_model, int) else resolve_model(func_or_model)
    o = _cabi.default_opts(model_id)
    o.compute_dtype = _cabi.F64 if compute_dtype == "f64" else _cabi.F32
    o.init_mode = _cabi.INIT_LOGLINEAR if init in ("loglinear", "polyfit") else _cabi.INIT_GIVEN
    o.maxfev, o.ftol, o.r2_eps = int(maxfev), float(ftol), float(eps)
    if p0 is not None:
        for i in range(P):
            o.p0[i] = float("nan") if p0[i] is None else float(p0[i])
    if y_bounds is not None:
        o.y_lo, o.y_hi = float(y_bounds[0]), float(y_bounds[1])
    for k, v in engine.items():
        if v is not None:
            setattr(o, k, type(getattr(o, k))(v))
    if post:
        o.post_enabled = 1
        for i in range(P):
            o.ufunc[i] = post.get("ufunc", [0] * P)[i]
            o.lb[i] = post.get("lb", [-np.inf] * P)[i]
            o.ub[i] = post.get("ub", [np.inf] * P)[i]
            o.decimals[i] = post.get("decimals", [-1] * P)[i]
        if post.get("r2_threshold") is not None:
            o.has_r2_threshold, o.r2_threshold = 1, float(post["r2_threshold"])
        if post.get("nan_to_num") is not None:
            o.has_nan_fill, o.nan_fill = 1, float(post["nan_to_num"])
    return o, P


def fit_device(opts, nparams, x, y, *, layout="planar", mask=None, p0_voxel=None, out_dtype=None, popt=None,
               r2=None, status=None, niter=None, handle=None):
    """Launch one fit on torch's current CUDA stream.

    y: CUDA tensor, (E, N) for layout="planar" or (N, E) for "echo_fastest" (last dim contiguous).
    mask: uint8/bool CUDA tensor [N] or None; p0_voxel: (N, P) float32/float64 CUDA tensor or None.
    Returns (popt (N, P), r2 (N,)) CUDA tensors (float32 unless out_dtype=torch.float64).
    Asynchronous: results are ordered on the current stream.
    """
    import torch

    assert y.is_cuda and y.dim() == 2 and y.stride(1) == 1
    dev = y.device.index
    planar = layout == "planar"
    E, N = (y.shape if planar else (y.shape[1], y.shape[0]))
    ld = y.stride(0)
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    assert x.shape[0] == E
    out_dtype = out_dtype or torch.float32
    if popt is None:
        popt = torch.empty((N, nparams), dtype=out_dtype, device=y.device)
    if r2 is None:
        r2 = torch.empty((N,), dtype=out_dtype, device=y.device)
    if mask is not None:
        mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous() and mask.numel() == N
    if p0_voxel is not None:
        assert p0_voxel.is_cuda and p0_voxel.is_contiguous() and tuple(p0_voxel.shape) == (N, nparams)
    h = handle or _cabi.get_handle(dev)
    stream = torch.cuda.current_stream(y.device).cuda_stream
    dt = _torch_dtypes()
    _cabi.check(_cabi.load().dfit_fit_device(
        h.ptr, ctypes.byref(opts), E, N, x.ctypes.data, y.data_ptr(), dt[y.dtype],
        _cabi.PLANAR if planar else _cabi.ECHO_FASTEST, ld,
        mask.data_ptr() if mask is not None else None,
        p0_voxel.data_ptr() if p0_voxel is not None else None,
        dt[p0_voxel.dtype] if p0_voxel is not None else _cabi.F32,
        popt.data_ptr(), r2.data_ptr(), dt[popt.dtype],
        status.data_ptr() if status is not None else None,
        niter.data_ptr() if niter is not None else None,
        ctypes.c_void_p(stream)))
    return popt, r2
