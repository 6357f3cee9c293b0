"""Golden fixtures for the "next" rows of the scope table, generated from the REAL
reference methods.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_next.py

  qdess_*.npz    inputs + output of `QDess.generate_t2_map` (dosma/scan_sequences/mri/qdess.py:100-258), the module
                 loaded verbatim through `ref_loader.load_reference_qdess()`
  metrics_*.npz  inputs + the table of `QuantitativeValue.to_metrics` (dosma/core/quant_vals.py:145-229), loaded
                 verbatim through `ref_loader.load_reference_quant_vals()`
These pin oracle/qdess_oracle.py and oracle/metrics_oracle.py (tests/test_qdess_oracle.py, tests/test_metrics_oracle.py)
and, through the same files, the CUDA kernels (tests/test_gpu_qdess.py, tests/test_gpu_metrics.py).
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader as R  # noqa: E402

F, MV = R.load_reference_fitting()
Q = R.load_reference_quant_vals()
QD = R.load_reference_qdess()
warnings.filterwarnings("ignore")

PARAMS = dict(tr=20.36, te=6.43, tg=3400.0, gl_area=3132.0, alpha=20.0, t1=1200.0)


def save(name, meta, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name:36s} {os.path.getsize(path) / 1024:8.1f} KiB")


def qdess_echoes(rng, shape, dtype):
    # signal ratios that map to T2 in (2, 130) ms for PARAMS (some beyond the (0, 100) bounds), plus edge voxels:
    # zero / negative / tiny / equal echoes (division by zero, log of 0, negative ratio)
    t2 = rng.uniform(2, 130, shape)
    s1 = rng.uniform(200, 1200, shape)
    s2 = s1 * 0.33 * np.exp(-27.86 / t2) * (1 + 0.02 * rng.standard_normal(shape))
    s1.flat[:8] = [0, 0, 5, -3, 1e-30, 800, 800, 1200]
    s2.flat[:8] = [0, 4, 0, 2, 1e-30, -50, 800, 1]
    if np.issubdtype(dtype, np.integer):
        s1, s2 = np.round(s1), np.round(s2)
    return s1.astype(dtype), s2.astype(dtype)


def case_qdess(name, dtype, seed, shape=(24, 20, 6), **kw):
    rng = np.random.default_rng(seed)
    s1, s2 = qdess_echoes(rng, shape, dtype)
    scan = QD.QDess([MV(s1, np.eye(4)), MV(s2, np.eye(4))])
    out = scan.generate_t2_map(**PARAMS, **kw)
    assert type(out).__name__ == "T2"
    meta = {"kind": "qdess", "seed": seed, "params": PARAMS,
            "kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}}
    save(name, meta, echo1=s1, echo2=s2, t2=np.asarray(out.volumetric_map.volume))


case_qdess("qdess_default_f32", np.float32, 40)
case_qdess("qdess_default_f64", np.float64, 41)
case_qdess("qdess_default_i16", np.int16, 42)
case_qdess("qdess_suppress_fat_fluid_f32", np.float32, 43, suppress_fat=True, suppress_fluid=True)
case_qdess("qdess_suppress_fluid_beta_i16", np.int16, 44, suppress_fluid=True, beta=0.9)
case_qdess("qdess_raw_f64", np.float64, 45, nan_bounds=None, nan_to_num=None, decimals=None)
case_qdess("qdess_bounds_nofill_3dec_f32", np.float32, 46, nan_bounds=(5, 60), nan_to_num=None, decimals=3)
case_qdess("qdess_fill_true_f32", np.float32, 47, nan_to_num=True, decimals=2)
case_qdess("qdess_fill_value_f64", np.float64, 48, nan_to_num=-1.0, decimals=0)


def case_metrics(name, seed, dtype=np.float64, shape=(40, 36, 12), with_mask=True, labels=None, **kw):
    rng = np.random.default_rng(seed)
    vol = np.round(rng.uniform(-5, 120, shape), 1).astype(dtype)  # a rounded tc map: many ties
    vol[rng.random(shape) < 0.05] = np.nan
    vol[rng.random(shape) < 0.01] = np.inf
    vol[rng.random(shape) < 0.01] = -np.inf
    vol[rng.random(shape) < 0.2] = 0.0
    lab = rng.integers(0, 5, shape).astype(np.uint8)
    qv = Q.T2(MV(vol, np.eye(4)))
    lab_arg = dict(labels) if labels is not None else None
    df = qv.to_metrics(mask=MV(lab, np.eye(4)) if with_mask else None, labels=lab_arg, **kw)
    meta = {"kind": "metrics", "seed": seed, "with_mask": with_mask, "labels": {str(k): v for k, v in (labels or {}).items()} or None,
            "kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
            "categories": [str(c) for c in df["Category"]]}
    save(name, meta, volume=vol, mask=lab, mean=df["Mean"].to_numpy(dtype=np.float64), std=df["Std"].to_numpy(dtype=np.float64),
         median=df["Median"].to_numpy(dtype=np.float64), count=df["# Voxels"].to_numpy(dtype=np.int64))


case_metrics("metrics_nomask_f64", 50, with_mask=False)
case_metrics("metrics_nomask_bounds_f64", 51, with_mask=False, bounds=(0, 100))
case_metrics("metrics_mask_f64", 52)
case_metrics("metrics_mask_bounds_right_f64", 53, bounds=(0, 100), closed="right")
case_metrics("metrics_mask_bounds_left_f64", 54, bounds=(0, 100), closed="left")
case_metrics("metrics_mask_bounds_both_f64", 55, bounds=(0, 100), closed="both")
case_metrics("metrics_mask_bounds_neither_f64", 56, bounds=(10, 90), closed="neither")
case_metrics("metrics_mask_labels_f64", 57, labels={2: "femoral", 4: "tibial", 9: "absent"}, bounds=(0, 100))
case_metrics("metrics_mask_f32", 58, dtype=np.float32, bounds=(0, 100))
