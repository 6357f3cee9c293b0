"""Region statistics of a fitted map on the GPU -- `QuantitativeValue.to_metrics`
(`dosma/core/quant_vals.py:145-229`), SURVEY.md section 8 row f4.

`region_metrics` returns the table the reference builds (columns "Category", "Mean", "Std",
"Median", "# Voxels"; one row per label plus "total"), computed by `dfit_region_metrics_host`:
streaming moment passes and an exact radix-selection median, all in float64.
"""
import numpy as np

from . import _cabi
from .fitting import _default_device
from .med_volume import is_volume

__all__ = ["region_metrics"]

_MAX_REGIONS = 16


def region_metrics(volumetric_map, mask=None, labels=None, bounds=None, closed="right", device=None, as_frame=True):
    """Per-region mean / std / median / voxel count of `volumetric_map` (ndarray or volume).

    Arguments as `QuantitativeValue.to_metrics` (quant_vals.py:145-152); `fns` (arbitrary Python
    reducers) is not supported on the GPU.  Returns a pandas DataFrame (or a dict of columns with
    as_frame=False).
    """
    vol = np.asarray(volumetric_map.volume if is_volume(volumetric_map) else volumetric_map)
    if bounds is not None:
        assert len(bounds) == 2, len(bounds)
        lb, ub = float(bounds[0]), float(bounds[1])
        assert lb <= ub, f"lower:{lb}, upper: {ub}"
        assert closed in ("right", "left", "both", "neither"), closed
    else:
        lb, ub = -np.inf, np.inf
    lab = None
    if mask is not None:
        if is_volume(mask):
            if is_volume(volumetric_map):
                mask = mask.reformat(volumetric_map.orientation)  # quant_vals.py:193
            mask = mask.volume
        lab = np.asarray(mask)
        if lab.shape != vol.shape:
            raise ValueError("mask and map must have the same shape")
        if labels is None:  # quant_vals.py:196-198
            labels = {int(i): f"label_{int(i)}" for i in np.unique(lab) if i > 0}
        labels = dict(labels)
        labels.update({-1: "total"})  # :199
    else:
        labels = {-2: "total"}  # :201
    if len(labels) > _MAX_REGIONS:
        raise NotImplementedError(f"at most {_MAX_REGIONS} regions per call")

    m = np.ascontiguousarray(vol.reshape(-1))
    if m.dtype not in (np.float32, np.float64):
        m = m.astype(np.float64)
    if lab is not None:
        lab = np.ascontiguousarray(lab.reshape(-1))
        if lab.dtype == np.bool_:
            lab = lab.view(np.uint8)
        elif lab.dtype not in _cabi.NP_TO_DTYPE:
            lab = lab.astype(np.int32)
    regions = np.asarray(list(labels.keys()), dtype=np.int32)
    out = np.empty((len(regions), 4), dtype=np.float64)
    h = _cabi.get_handle(_default_device() if device is None else device)
    _cabi.check(_cabi.load().dfit_region_metrics_host(
        h.ptr, m.shape[0], m.ctypes.data, _cabi.NP_TO_DTYPE[m.dtype],
        lab.ctypes.data if lab is not None else None, _cabi.NP_TO_DTYPE[lab.dtype] if lab is not None else 0,
        len(regions), regions.ctypes.data, int(bounds is not None), lb, ub,
        int(closed in ("left", "both")), int(closed in ("right", "both")), out.ctypes.data))
    cols = {"Category": list(labels.values()), "Mean": out[:, 1].tolist(), "Std": out[:, 2].tolist(),
            "Median": out[:, 3].tolist(), "# Voxels": [int(c) for c in out[:, 0]]}
    if as_frame:
        import pandas as pd

        return pd.DataFrame(cols)
    return cols
