"""CPU checks of the qDESS oracle restatement against closed-form values."""
import numpy as np

from oracle import qdess_oracle as Q

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
