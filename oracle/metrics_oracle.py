"""CPU ORACLE for region statistics -- test infrastructure, NOT product code.

numpy restatement of `QuantitativeValue.to_metrics` (dosma/core/quant_vals.py:181-229) at array
level.  Parity status: PINNED -- tests/test_metrics_oracle.py checks it, bit for bit, against
tests/golden/metrics_*.npz, the tables of the real reference method (tests/golden/make_golden_next.py loads
quant_vals.py verbatim through tests/golden/ref_loader.py::load_reference_quant_vals), and in the build container
against the live reference on fresh seeds.
"""
import warnings

import numpy as np


def to_metrics(volume, mask=None, labels=None, bounds=None, closed="right"):
    volume = np.asarray(volume)
    valid_mask = np.isfinite(volume)  # :182
    if bounds:
        lb, ub = bounds
        lb_mask = volume >= lb if closed in ("left", "both") else volume > lb  # :188
        ub_mask = volume <= ub if closed in ("right", "both") else volume < ub  # :189
        valid_mask &= lb_mask & ub_mask
    if mask is not None:
        mask = np.asarray(mask)
        if labels is None:
            labels = {int(i): f"label_{int(i)}" for i in np.unique(mask) if i > 0}  # :196-198
        labels = dict(labels)
        labels.update({-1: "total"})
        mask = mask.copy()
        mask[~valid_mask] = 0  # :205-206
    else:
        labels = {-2: "total"}
    out = {"Category": [], "Mean": [], "Std": [], "Median": [], "# Voxels": []}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for label, name in labels.items():
            if label == -2:
                vals = volume[valid_mask]
            elif label == -1:
                vals = volume[mask > 0]
            else:
                vals = volume[mask == label]
            out["Category"].append(name)
            out["Mean"].append(float(np.nanmean(vals)))
            out["Std"].append(float(np.nanstd(vals)))
            out["Median"].append(float(np.nanmedian(vals)))
            out["# Voxels"].append(int(np.prod(vals.shape)))
    return out
