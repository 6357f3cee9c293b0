"""Conformance of the drop-in with the REFERENCE'S OWN test-suite and container type (build container only).

`/root/reference/tests/core/test_fitting.py` (TestCurveFit :70-140, TestMonoExponentialFit :199-277, TestCurveFitter
:280-571) is loaded verbatim and run against `dosma_b200.{curve_fit, MonoExponentialFit, CurveFitter}` swapped into
`dosma.core.fitting` by `dosma_b200.patch_dosma()`, on the REAL `dosma.core.med_volume.MedicalVolume`.  There is no GPU
in the build container, so the engine call (`dosma_b200.fitting._engine_fit`, one ctypes call into libdfit.so) is
replaced by `tests.hostsim.engine_fit`: the same device solver headers compiled by g++.  Everything above that call
-- argument handling, p0 forms, reorientation, masks, `_process_params` planning, header propagation, the
MedicalVolume wrapping -- is the product's code.  The GPU suite repeats the header / 4-D / patch cases on the real
engine with the package's own container (tests/test_gpu_parity.py).
"""
import importlib.util
import os
import sys
import types
import unittest
from copy import deepcopy

import numpy as np
import pytest

pytestmark = pytest.mark.needs_reference

REF_TEST = "/root/reference/tests/core/test_fitting.py"


class _Header(dict):
    """Stand-in for a pydicom FileDataset (not installed here): attribute and `.get` access, deep-copyable."""

    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return _Header({k: deepcopy(v, memo) for k, v in self.items()})


def _build_dummy_headers(shape, fields=None):
    """tests/util.py:136-153 with `_Header` objects instead of pydicom datasets."""
    if isinstance(shape, int):
        shape = (shape,)
    n = int(np.prod(shape))
    arr = np.empty(n, dtype=object)
    for i in range(n):
        arr[i] = _Header(fields or {})
    return arr.reshape(shape)


@pytest.fixture(scope="module")
def ref_tests():
    import dosma_b200
    from dosma_b200 import fitting as DF
    from tests import hostsim as H
    from tests.golden import ref_loader as R

    F, MV = R.load_reference_fitting()
    saved_engine = DF._engine_fit
    saved = {k: getattr(F, k) for k in ("CurveFitter", "MonoExponentialFit", "curve_fit")}
    DF._engine_fit = H.engine_fit  # no GPU here: the host build of the device solver stands in for libdfit.so
    patched = dosma_b200.patch_dosma()
    assert {"dosma.core.fitting.CurveFitter", "dosma.core.fitting.MonoExponentialFit",
            "dosma.core.fitting.curve_fit"} <= set(patched), patched
    assert F.CurveFitter is dosma_b200.CurveFitter and F.MonoExponentialFit is dosma_b200.MonoExponentialFit

    pkg = types.ModuleType("dosma_ref_tests")
    pkg.__path__ = []
    core = types.ModuleType("dosma_ref_tests.core")
    core.__path__ = []
    util = types.ModuleType("dosma_ref_tests.util")
    util.build_dummy_headers = _build_dummy_headers
    util.num_workers = lambda: 2
    pkg.util, pkg.core = util, core
    sys.modules.update({"dosma_ref_tests": pkg, "dosma_ref_tests.core": core, "dosma_ref_tests.util": util})
    spec = importlib.util.spec_from_file_location("dosma_ref_tests.core.test_fitting", REF_TEST)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    assert mod.CurveFitter is dosma_b200.CurveFitter and mod.MedicalVolume is MV
    yield mod
    DF._engine_fit = saved_engine
    for k, v in saved.items():
        setattr(F, k, v)


def _run(mod, cls_name, test_name):
    suite = unittest.TestSuite([getattr(mod, cls_name)(test_name)])
    res = unittest.TestResult()
    suite.run(res)
    problems = res.errors + res.failures
    assert not problems, problems[0][1]


REF_CASES = [
    ("TestCurveFit", "test_multiple_workers"), ("TestCurveFit", "test_p0"),
    ("TestMonoExponentialFit", "test_basic"), ("TestMonoExponentialFit", "test_headers"),
    ("TestMonoExponentialFit", "test_mask"), ("TestMonoExponentialFit", "test_polyfit_initialization"),
    ("TestCurveFitter", "test_basic"), ("TestCurveFitter", "test_mask"), ("TestCurveFitter", "test_bounds"),
    ("TestCurveFitter", "test_out_ufuncs"), ("TestCurveFitter", "test_nan_to_num"),
    ("TestCurveFitter", "test_matches_monoexponential_fit"), ("TestCurveFitter", "test_headers"),
    ("TestCurveFitter", "test_p0"), ("TestCurveFitter", "test_str"),
]


# `TestCurveFitter.test_headers` indexes the result with `popt[..., 0]`, which raises inside the REFERENCE's
# `MedicalVolume.__getitem__` (med_volume.py:1246 indexes the header array with a list; numpy >= 1.23 rejects that) --
# with the reference's own CurveFitter just the same.  The checks of that test are repeated
# below without that indexing step.
REF_SIDE_FAILURE = {("TestCurveFitter", "test_headers")}


@pytest.mark.parametrize("cls_name,test_name", REF_CASES)
def test_reference_test_fitting(ref_tests, cls_name, test_name):
    if (cls_name, test_name) in REF_SIDE_FAILURE:
        with pytest.raises(AssertionError, match="med_volume.py.*__getitem__"):
            _run(ref_tests, cls_name, test_name)
        return
    _run(ref_tests, cls_name, test_name)


def test_curvefitter_headers_4d_on_real_medicalvolume(ref_tests):
    """The assertions of the reference's `TestCurveFitter.test_headers` (:433-477) -- 4-D inputs, header-bearing
    volumes, a one-parameter model, `copy_headers=False` -- on the real MedicalVolume, slicing the parameter axis
    on the arrays instead of through `MedicalVolume.__getitem__`."""
    m = ref_tests
    x, y, b = m._generate_monoexp_data((10, 10, 20, 4))
    for idx, _y in enumerate(y):
        _y._headers = _build_dummy_headers((1, 1) + _y.shape[2:], fields={"EchoNumbers": idx})
    popt, r2 = m.CurveFitter(m.monoexponential).fit(x, y)
    assert type(popt) is m.MedicalVolume and popt.shape == (10, 10, 20, 4, 2) and r2.shape == (10, 10, 20, 4)
    assert popt.volume.dtype == np.float64 and r2.volume.dtype == np.float64
    assert np.allclose(popt.volume[..., 0], 1.0) and np.allclose(popt.volume[..., 1], b)
    assert popt.headers() is not None and popt.headers().shape == (1, 1, 20, 4, 1)  # fitting.py:503-507
    assert r2.headers() is not None and r2.headers().shape == (1, 1, 20, 4)
    assert all(h.get("EchoNumbers") == 0 for h in popt.headers().flatten())  # y[0]'s headers, deep-copied
    assert popt.headers().flatten()[0] is not y[0].headers().flatten()[0]
    assert np.all(popt.affine == y[0].affine)
    # one-parameter model
    x, y, a = m._generate_linear_data((10, 10, 20, 4))
    for idx, _y in enumerate(y):
        _y._headers = _build_dummy_headers((1, 1) + _y.shape[2:], fields={"EchoNumbers": idx})
    popt, _ = m.CurveFitter(m._linear).fit(x, y)
    assert popt.shape == (10, 10, 20, 4, 1) and np.allclose(popt.volume[..., 0], a)
    # not copying headers
    x, y, b = m._generate_monoexp_data((10, 10, 20, 4))
    for idx, _y in enumerate(y):
        _y._headers = _build_dummy_headers((1, 1) + _y.shape[2:], fields={"EchoNumbers": idx})
    popt, r2 = m.CurveFitter(m.monoexponential).fit(x, y, copy_headers=False)
    assert np.allclose(popt.volume[..., 1], b) and popt.headers() is None and r2.headers() is None


def test_outputs_feed_quant_vals_unchanged(ref_tests):
    """`north_star`: the maps must be MedicalVolumes that `dosma.core.quant_vals` consumes unchanged -- wrap the drop-in's
    T2 map in the real `quant_vals.T2` and run the real `to_metrics` on it."""
    from tests.golden import ref_loader as R

    m = ref_tests
    Q = R.load_reference_quant_vals()
    x, y, b = m._generate_monoexp_data((8, 8, 6))
    tc, r2 = m.MonoExponentialFit(decimal_precision=3).fit(x, y)
    assert type(tc) is m.MedicalVolume and tc.volume.dtype == np.float64
    qv = Q.T2(tc)
    qv.add_additional_volume("r2", r2)
    df = qv.to_metrics(bounds=(0, 100), closed="right")
    assert int(df["# Voxels"][0]) == 8 * 8 * 6
    assert abs(float(df["Mean"][0]) - np.around(1 / np.abs(b), 3).mean()) < 1e-9
