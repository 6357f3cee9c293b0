"""CPU tests of the host-side mirror of the reference interface: everything that happens before a
launch (argument forms, validation errors, volume marshalling).  Mirrors the error/argument checks
of the reference's tests/core/test_fitting.py."""
import warnings

import numpy as np
import pytest

import dosma_b200 as D
from dosma_b200 import _cabi, fitting as F, models as M, sharding as S
from dosma_b200.med_volume import MedicalVolume


def test_model_registry_by_identity_and_fingerprint():
    assert M.resolve_model(D.monoexponential) == (_cabi.MODEL_MONOEXP, 2)
    assert M.resolve_model(D.biexponential) == (_cabi.MODEL_BIEXP, 4)

    def my_mono(t, amp, rate):  # README.md:113-114 style user re-definition
        return amp * np.exp(rate * t)

    def _linear(x, a):  # tests/core/test_fitting.py:52-53
        return a * x

    assert M.resolve_model(my_mono) == (_cabi.MODEL_MONOEXP, 2)
    assert M.resolve_model(_linear) == (_cabi.MODEL_LINEAR, 1)
    with pytest.raises(NotImplementedError):
        M.resolve_model(lambda x, a, b: a * x + b)


def test_ufunc_registry():
    assert M.resolve_ufunc(None) == _cabi.UFUNC_NONE
    assert M.resolve_ufunc(lambda v: 1 / np.abs(v)) == _cabi.UFUNC_INV_ABS
    assert M.resolve_ufunc(lambda v: -1 / v) == _cabi.UFUNC_NEG_INV
    assert M.resolve_ufunc(np.abs) == _cabi.UFUNC_ABS
    assert M.resolve_ufunc(lambda v: 2 * np.abs(v) + 5) is None


def test_split_p0_forms():
    names = ["a", "b"]
    assert F._split_p0(None, names, 5) == [1.0, 1.0]
    assert F._split_p0(3, names, 5) == [3.0, 3.0]
    assert F._split_p0((None, 2.0), names, 5) == [1.0, 2.0]
    assert F._split_p0({"b": 50.0}, names, 5) == [1.0, 50.0]
    cols = F._split_p0(np.ones((5, 2)), names, 5)
    assert all(isinstance(c, np.ndarray) and c.shape == (5,) for c in cols)
    cols = F._split_p0([np.ones(5), 50], names, 5)
    assert isinstance(cols[0], np.ndarray) and cols[1] == 50.0
    with pytest.raises(ValueError):
        F._split_p0((1.0, 2.0, 3.0), names, 5)
    with pytest.raises(ValueError):
        F._split_p0({"c": 1.0}, names, 5)
    with pytest.raises(ValueError):
        F._split_p0([np.ones(4), 1.0], names, 5)


def test_constructor_validation():
    with pytest.raises(ValueError):
        D.CurveFitter(D.monoexponential, out_bounds=[(0, 0.5, 1.0)])
    with pytest.raises(ValueError):
        D.CurveFitter(D.monoexponential, out_bounds=[(1.2, 0)])
    with pytest.raises(TypeError):
        D.CurveFitter(D.monoexponential, out_ufuncs=[None, 5])
    with pytest.warns(UserWarning):
        D.CurveFitter(D.monoexponential, out_ufuncs=[None, np.abs, np.abs])
    with pytest.raises(ValueError):
        D.CurveFitter(D.monoexponential, r2_threshold="bogus")
    assert D.CurveFitter(D.monoexponential).r2_threshold == 0.9
    assert "func=monoexponential" in str(D.CurveFitter(D.monoexponential, p0=(1.0, -1 / 30)))

    x = np.asarray([0.5, 1.0, 2.0, 4.0])
    y = [MedicalVolume(np.ones((4, 4, 2)), np.eye(4)) for _ in x]
    with pytest.warns(UserWarning):
        D.MonoExponentialFit(x, y)
    with pytest.warns(UserWarning):
        D.MonoExponentialFit(mask=y[0])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ValueError):
            D.MonoExponentialFit(list(x) + [5], y)
        with pytest.raises(TypeError):
            D.MonoExponentialFit(x, [_y.A for _y in y])
    with pytest.raises(ValueError):
        D.MonoExponentialFit(tc0="a value")
    with pytest.raises(ValueError):
        D.MonoExponentialFit(bounds=(0, 1, 2))


def test_fit_validation_before_launch():
    x = np.asarray([0.5, 1.0, 2.0, 4.0])
    y = [MedicalVolume(np.ones((4, 4, 2)), np.eye(4)) for _ in x]
    with pytest.raises(TypeError):
        D.CurveFitter(D.monoexponential).fit(x, [v.A for v in y])
    with pytest.raises(ValueError):
        D.CurveFitter(D.monoexponential).fit(x[:3], y)
    with pytest.raises(TypeError):
        D.CurveFitter(D.monoexponential).fit(x, y, mask="foo")
    with pytest.raises(RuntimeError):
        D.CurveFitter(D.monoexponential).fit(x, y, mask=np.ones((5, 5, 5)))
    with pytest.raises(ValueError):
        D.CurveFitter(D.monoexponential).fit(x, y, p0=(1.0, np.ones((3, 3, 3))))
    with pytest.raises(NotImplementedError):
        D.curve_fit(D.monoexponential, x, np.ones((4, 3)), bounds=(0, 1))
    with pytest.raises(NotImplementedError):
        D.curve_fit(lambda x, a, b, c: a + b * x + c * x * x, x, np.ones((4, 3)))


def test_medical_volume_reformat_roundtrip():
    rng = np.random.default_rng(0)
    aff = np.array([[0, 0, 1.5, -61.7], [-0.3125, 0, 0, 50.9], [0, -0.3125, 0, 88.6], [0, 0, 0, 1.0]])
    mv = MedicalVolume(rng.random((4, 5, 6)), aff, headers=np.array([{"i": i} for i in range(6)]).reshape(1, 1, 6))
    assert mv.orientation == ("AP", "SI", "LR")
    r = mv.reformat(("LR", "PA", "IS"))
    assert r.orientation == ("LR", "PA", "IS") and r.shape == (6, 4, 5)
    assert r.headers().shape == (6, 1, 1)
    # the same physical voxel keeps its world coordinate
    ijk = np.array([1, 2, 3, 1.0])
    world = aff @ ijk
    ijk_r = np.array([3, 4 - 1 - 1, 5 - 1 - 2, 1.0])
    assert np.allclose(r.affine @ ijk_r, world, atol=1e-3)
    assert r.volume[3, 2, 2] == mv.volume[1, 2, 3]
    back = r.reformat_as(mv)
    assert back.is_identical(mv)
    assert mv.is_same_dimensions(back, precision=4)
    s = mv[1:3, :, ::2]
    assert s.shape == (2, 5, 3) and np.allclose(s.affine[:3, 3], aff[:3, 3] + aff[:3, 0])
    assert s.headers().shape == (1, 1, 3)
    with pytest.raises(IndexError):
        mv[0]


def test_partition_helpers():
    assert S.slab_bounds(160, 8).tolist() == list(range(0, 161, 20))
    b = S.slab_bounds(10, 4)
    assert b.tolist() == [0, 3, 6, 8, 10]
    r = S.voxel_ranges(1000, 3, align=128)
    assert r[0] == 0 and r[-1] == 1000 and all(x % 128 == 0 for x in r[:-1]) and np.all(np.diff(r) >= 0)


def test_masked_spans_balance_a_clustered_mask():
    """`sharding.masked_spans`: consecutive spans covering the volume, each with (nearly) the same number of masked
    voxels however the tissue is clustered; cuts on aligned voxel indices; degenerate masks."""
    from dosma_b200 import sharding as S

    rng = np.random.default_rng(0)
    n = 200_000
    z = np.arange(n) / n
    mask = rng.random(n) < 0.3 * np.exp(-((z - 0.7) / 0.05) ** 2)  # all the tissue in a thin slab
    for world in (1, 2, 3, 8):
        b = S.masked_spans(mask, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1 and (np.diff(b) >= 0).all() and (b[1:-1] % 4 == 0).all()
        counts = [int(mask[b[r]:b[r + 1]].sum()) for r in range(world)]
        assert sum(counts) == int(mask.sum()) and max(counts) - min(counts) <= 4
    assert S.masked_spans(np.zeros(100, bool), 4).tolist() == [0, 24, 48, 72, 100]
    assert S.masked_spans(np.ones(10, bool), 2).tolist() == [0, 4, 10]
