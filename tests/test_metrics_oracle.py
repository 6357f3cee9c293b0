"""Pins oracle/metrics_oracle.py (numpy restatement of `QuantitativeValue.to_metrics`,
dosma/core/quant_vals.py:145-229) to the outputs of the REAL reference method: the golden fixtures
tests/golden/metrics_*.npz were written by tests/golden/make_golden_next.py, which loads quant_vals.py verbatim."""
import numpy as np
import pytest

from oracle import metrics_oracle as M
from tests import golden_util as G


def run_case(fn, c):
    """Call a to_metrics implementation (`fn(volume, mask=, labels=, bounds=, closed=)`) the way the fixture was made."""
    meta = c["meta"]
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in meta["kwargs"].items()}
    labels = {int(k): v for k, v in meta["labels"].items()} if meta["labels"] else None
    return fn(c["volume"], mask=c["mask"] if meta["with_mask"] else None, labels=labels, **kw)


def check_against_golden(got, c, rtol):
    assert list(got["Category"]) == c["meta"]["categories"]
    assert [int(v) for v in got["# Voxels"]] == c["count"].tolist()
    np.testing.assert_allclose(np.asarray(got["Mean"], dtype=np.float64), c["mean"], rtol=rtol, equal_nan=True)
    np.testing.assert_allclose(np.asarray(got["Std"], dtype=np.float64), c["std"], rtol=rtol, equal_nan=True)
    np.testing.assert_array_equal(np.asarray(got["Median"], dtype=np.float64), c["median"])


@pytest.mark.parametrize("name", G.names("metrics_"))
def test_metrics_oracle_matches_reference(name):
    c = G.load(name)
    check_against_golden(run_case(M.to_metrics, c), c, rtol=0)  # the same numpy calls in the same order: bit-identical


@pytest.mark.needs_reference
def test_metrics_oracle_matches_live_reference():
    """Fresh seeds against the reference method itself (build container only)."""
    from tests.golden import ref_loader as R

    F, MV = R.load_reference_fitting()
    Q = R.load_reference_quant_vals()
    rng = np.random.default_rng(99)
    vol = np.round(rng.uniform(-5, 120, (20, 18, 7)), 1)
    vol[rng.random(vol.shape) < 0.1] = np.nan
    lab = rng.integers(0, 4, vol.shape).astype(np.uint8)
    for kw in (dict(), dict(bounds=(0, 100)), dict(bounds=(0, 50), closed="both")):
        df = Q.T2(MV(vol, np.eye(4))).to_metrics(mask=MV(lab, np.eye(4)), **kw)
        got = M.to_metrics(vol, mask=lab, **kw)
        assert list(df["Category"]) == got["Category"] and df["# Voxels"].tolist() == got["# Voxels"]
        for k in ("Mean", "Std", "Median"):
            np.testing.assert_array_equal(df[k].to_numpy(), np.asarray(got[k]))
