"""GPU parity of region statistics (dfit_region_metrics_host) against the numpy oracle."""
import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _check(got, ref, rtol=1e-11):
    assert got["Category"] == ref["Category"], (got["Category"], ref["Category"])
    assert got["# Voxels"] == ref["# Voxels"], (got["# Voxels"], ref["# Voxels"])
    for k in ("Mean", "Std"):
        np.testing.assert_allclose(got[k], ref[k], rtol=rtol, equal_nan=True)
    np.testing.assert_array_equal(np.asarray(got["Median"]), np.asarray(ref["Median"]))  # exact selection


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_metrics_match_oracle(dtype):
    from dosma_b200.metrics import region_metrics
    from oracle import metrics_oracle as M

    rng = np.random.default_rng(0)
    shape = (96, 80, 40)
    vol = np.round(rng.uniform(-5, 120, shape), 1).astype(dtype)  # rounded map: many ties, like a real tc map
    vol[rng.random(shape) < 0.05] = np.nan
    vol[rng.random(shape) < 0.01] = np.inf
    vol[rng.random(shape) < 0.2] = 0.0
    lab = rng.integers(0, 5, shape).astype(np.uint8)
    # numpy reduces a float32 map in float32 (pairwise), the kernel always accumulates in float64: for float32
    # maps the oracle itself is only good to ~1e-6; fitted maps are float64 (fitting.py:870), the exact case
    rtol = 1e-11 if dtype == np.float64 else 2e-6

    def check(got, ref):
        _check(got, ref, rtol)

    for kw in (dict(), dict(bounds=(0, 100)), dict(bounds=(0, 100), closed="both"), dict(bounds=(10, 90), closed="neither")):
        check(region_metrics(vol, as_frame=False, **kw), M.to_metrics(vol, **kw))
        check(region_metrics(vol, mask=lab, as_frame=False, **kw), M.to_metrics(vol, mask=lab, **kw))
    sel = {2: "femoral", 4: "tibial", 9: "absent"}
    check(region_metrics(vol, mask=lab.astype(np.int32), labels=sel, as_frame=False, bounds=(0, 100)),
           M.to_metrics(vol, mask=lab, labels=sel, bounds=(0, 100)))
    frame = region_metrics(vol, mask=lab)
    assert list(frame.columns) == ["Category", "Mean", "Std", "Median", "# Voxels"]


def test_metrics_full_size_properties():
    """384^3 map: median equals the sorted middle element; odd/even counts; negative values order correctly."""
    import torch

    from dosma_b200.metrics import region_metrics

    g = torch.Generator().manual_seed(1)
    n = 384 * 384 * 96
    vol = (torch.randn(n, generator=g, dtype=torch.float64) * 30 + 40).numpy()
    got = region_metrics(vol, as_frame=False)
    assert got["# Voxels"][0] == n
    assert got["Median"][0] == float(np.median(vol))
    assert abs(got["Mean"][0] - vol.mean()) < 1e-9 and abs(got["Std"][0] - vol.std()) < 1e-9
    got = region_metrics(vol[:-1], as_frame=False)
    assert got["Median"][0] == float(np.median(vol[:-1]))


@pytest.mark.parametrize("name", G.names("metrics_"))
def test_metrics_golden_reference_outputs(name):
    """The CUDA reductions against the tables of the REAL `QuantitativeValue.to_metrics` (tests/golden/metrics_*.npz):
    categories, voxel counts and medians exactly; mean / std to float64 summation-order accuracy (a float32 map is
    reduced in float32 by numpy and in float64 by the kernel: 2e-6 there)."""
    from dosma_b200.metrics import region_metrics
    from tests.test_metrics_oracle import check_against_golden, run_case

    c = G.load(name)
    got = run_case(lambda vol, **kw: region_metrics(vol, as_frame=False, **kw), c)
    check_against_golden(got, c, rtol=1e-11 if c["volume"].dtype == np.float64 else 2e-6)
