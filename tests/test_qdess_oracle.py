"""Pins oracle/qdess_oracle.py (numpy restatement of `QDess.generate_t2_map`, dosma/scan_sequences/mri/qdess.py:193-255):
against the outputs of the REAL reference method -- the golden fixtures tests/golden/qdess_*.npz written by
tests/golden/make_golden_next.py, which loads qdess.py verbatim -- and against closed-form values."""
import numpy as np
import pytest

from oracle import qdess_oracle as Q
from tests import golden_util as G


def golden_kwargs(c):
    kw = dict(c["meta"]["params"])
    kw.update({k: (tuple(v) if isinstance(v, list) else v) for k, v in c["meta"]["kwargs"].items()})
    return kw


@pytest.mark.parametrize("name", G.names("qdess_"))
def test_qdess_oracle_matches_reference(name):
    """Same numpy calls in the same order as the reference: bit-identical, NaN for NaN."""
    c = G.load(name)
    got = Q.t2_map(c["echo1"], c["echo2"], **golden_kwargs(c))
    assert got.dtype == c["t2"].dtype and got.shape == c["t2"].shape
    assert np.array_equal(got, c["t2"], equal_nan=True)


@pytest.mark.needs_reference
def test_qdess_oracle_matches_live_reference():
    """Fresh seeds against the reference method itself (build container only)."""
    from tests.golden import ref_loader as R

    F, MV = R.load_reference_fitting()
    QD = R.load_reference_qdess()
    rng = np.random.default_rng(77)
    s1 = rng.uniform(100, 1000, (10, 9, 4))
    s2 = s1 * rng.uniform(0.02, 0.5, s1.shape)
    for kw in (dict(), dict(suppress_fat=True, suppress_fluid=True, beta=1.1), dict(nan_bounds=(1, 40), decimals=None)):
        ref = QD.QDess([MV(s1, np.eye(4)), MV(s2, np.eye(4))]).generate_t2_map(**PARAMS, **kw).volumetric_map.volume
        assert np.array_equal(Q.t2_map(s1, s2, **PARAMS, **kw), ref, equal_nan=True)


PARAMS = dict(tr=20.36, te=6.43, tg=3400.0, gl_area=3132.0, alpha=20.0, t1=1200.0)


def synth(shape, rng, t2_true):
    """Echo pair whose analytic map is exactly t2_true: invert qdess.py:232 for S2/S1."""
    k, c1, TR, TE = Q.constants(**PARAMS)
    ratio = k * np.exp(-2000 * (TR - TE) / t2_true - c1)
    s1 = rng.uniform(200, 1200, shape)
    return s1, s1 * ratio


def test_round_trip_and_postprocessing():
    rng = np.random.default_rng(0)
    t2 = rng.uniform(5, 95, (12, 10, 6))
    s1, s2 = synth(t2.shape, rng, t2)
    got = Q.t2_map(s1, s2, **PARAMS, decimals=None)
    assert np.allclose(got, t2, rtol=1e-10)
    got = Q.t2_map(s1, s2, **PARAMS, decimals=1)
    assert np.allclose(got, np.around(t2, 1), atol=1e-9)
    t2b = t2.copy()
    t2b[0] = 150.0  # out of (0, 100) -> NaN -> 0
    s1, s2 = synth(t2.shape, rng, t2b)
    got = Q.t2_map(s1, s2, **PARAMS)
    assert (got[0] == 0).all() and np.allclose(got[1:], np.around(t2[1:], 1), atol=1e-9)
    s1[1, 0, 0] = 0.0  # division by zero -> inf -> nan_to_num -> huge ratio -> negative t2 -> out of bounds -> 0
    s2[1, 0, 1] = 0.0  # ratio 0 -> log(0) = -inf -> t2 = -0.0 -> in bounds
    got = Q.t2_map(s1, s2, **PARAMS)
    assert got[1, 0, 0] == 0 and got[1, 0, 1] == 0
    sf = Q.t2_map(s1, s2, **PARAMS, suppress_fat=True, suppress_fluid=True)
    keep = (s1 > 0.15 * s1.max()) & ((s1 - 1.2 * s2) > 0.1 * (s1 - 1.2 * s2).max())
    assert np.array_equal(sf != 0, (got != 0) & keep)
