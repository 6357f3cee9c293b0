"""Helpers to read the fixtures under tests/golden/ (written by tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "meta"}
    d["meta"] = json.loads(str(z["meta"]))
    return d


def p0_of(case):
    if "p0_arr" in case:
        return case["p0_arr"]
    p0 = case["meta"]["p0"]
    return tuple(p0) if isinstance(p0, list) else p0


def same_nan(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b))
