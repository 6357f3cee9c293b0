"""GPU parity of the qDESS analytic T2 map (C-ABI dfit_qdess_t2_host) against the numpy oracle."""
import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu

PARAMS = dict(tr=20.36, te=6.43, tg=3400.0, gl_area=3132.0, alpha=20.0, t1=1200.0)


def _echoes(rng, shape, dtype):
    from oracle import qdess_oracle as Q

    k, c1, TR, TE = Q.constants(**PARAMS)
    t2 = rng.uniform(2, 130, shape)  # some beyond the (0, 100) bounds
    s1 = rng.uniform(200, 1200, shape)
    s2 = s1 * k * np.exp(-2000 * (TR - TE) / t2 - c1) * (1 + 0.02 * rng.standard_normal(shape))
    s1.flat[:7] = [0, 0, 5, -3, 1e-30, 800, 800]
    s2.flat[:7] = [0, 4, 0, 2, 1e-30, -50, 800]
    if np.issubdtype(dtype, np.integer):
        s1, s2 = np.round(s1), np.round(s2)
    return s1.astype(dtype), s2.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16])
@pytest.mark.parametrize("suppress", [False, True])
def test_qdess_exact_matches_oracle(dtype, suppress):
    """Default path: the reference's float64 arithmetic -> bit-level agreement with numpy up to libm ulps."""
    from dosma_b200.qdess import qdess_t2_map
    from oracle import qdess_oracle as Q

    rng = np.random.default_rng(4)
    s1, s2 = _echoes(rng, (64, 48, 20), dtype)
    kw = dict(suppress_fat=suppress, suppress_fluid=suppress)
    for decimals in (1, None):
        got = qdess_t2_map(s1, s2, **PARAMS, decimals=decimals, **kw)
        ref = Q.t2_map(s1, s2, **PARAMS, decimals=decimals, **kw)
        assert got.shape == ref.shape and got.dtype == np.float64
        if decimals is None:
            assert np.allclose(got, ref, rtol=1e-13, atol=1e-13)
        else:  # a 1-ulp difference in log() can flip a value sitting exactly on a rounding boundary
            assert (got != ref).mean() < 1e-5 and np.abs(got - ref).max() < 0.1001


def test_qdess_fast_path_tolerance():
    """precision="fast": fp32 + MUFU; ~2e-6 relative before rounding, so after rounding to one decimal only
    values within that distance of a rounding boundary (or of the 100 ms bound) can differ, by one step."""
    from dosma_b200.qdess import qdess_t2_map
    from oracle import qdess_oracle as Q

    rng = np.random.default_rng(6)
    s1, s2 = _echoes(rng, (64, 48, 20), np.float32)
    got = qdess_t2_map(s1, s2, **PARAMS, decimals=None, precision="fast")
    ref = Q.t2_map(s1, s2, **PARAMS, decimals=None)
    assert got.dtype == np.float32
    near_bound = (np.abs(ref) > 99.99) | (np.abs(got) > 99.99)
    ok = np.isclose(got, ref, rtol=5e-6, atol=1e-6) | near_bound
    assert ok.mean() > 0.9999
    got = qdess_t2_map(s1, s2, **PARAMS, precision="fast", suppress_fat=True)
    ref = Q.t2_map(s1, s2, **PARAMS, suppress_fat=True)
    assert (np.abs(got - ref) > 0.1001).mean() < 1e-4 and (np.abs(got - ref) > 1e-5).mean() < 2e-3


def test_qdess_volume_wrapper_and_nan_options():
    import dosma_b200 as D
    from dosma_b200.qdess import qdess_t2_map
    from oracle import qdess_oracle as Q

    rng = np.random.default_rng(5)
    s1, s2 = _echoes(rng, (16, 16, 4), np.float64)
    v1, v2 = D.MedicalVolume(s1, np.eye(4)), D.MedicalVolume(s2, np.eye(4))
    out = qdess_t2_map(v1, v2, **PARAMS, nan_bounds=None, nan_to_num=None, decimals=3)
    ref = Q.t2_map(s1, s2, **PARAMS, nan_bounds=None, nan_to_num=None, decimals=3)
    assert isinstance(out, D.MedicalVolume) and np.allclose(out.volume, ref, rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("name", G.names("qdess_"))
def test_qdess_golden_reference_outputs(name):
    """The CUDA kernel (exact float64 path) against the outputs of the REAL `QDess.generate_t2_map`
    (tests/golden/qdess_*.npz): identical up to libm-ulp effects -- unrounded maps agree to 1e-13, rounded maps
    differ on at most a few voxels sitting on a rounding boundary, by one step; the NaN set is identical."""
    from dosma_b200.qdess import qdess_t2_map
    from tests.test_qdess_oracle import golden_kwargs

    c = G.load(name)
    kw = golden_kwargs(c)
    got = qdess_t2_map(c["echo1"], c["echo2"], **kw)
    ref = c["t2"]
    assert got.dtype == ref.dtype and got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    dec = kw.get("decimals", 1)
    fin = ~np.isnan(ref)
    if dec is None:
        assert np.allclose(got[fin], ref[fin], rtol=1e-13, atol=1e-13)
    else:
        step = 10.0 ** -dec
        assert (got[fin] != ref[fin]).mean() < 2e-3 and np.abs(got[fin] - ref[fin]).max() <= step * 1.0001
