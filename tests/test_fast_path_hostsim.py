"""CPU tests of the mono-exponential fast path (variable-projection Newton, dosma_b200/csrc/mono_fast.cuh
compiled by g++ through tests/hostsim): the two-voxels-per-lane solvers that the CUDA kernels run, against
the LM from p0 (fast=0) and against the C oracle (MINPACK restatement), over echo spacings, echo counts,
SNR and decay ranges.  High SNR: same minimiser to rounding.  Low SNR: the cost function has several
local minima on some voxels; the fast path may only differ there as rarely as the bounds below say, and
must not end in a worse minimum than the reference more often than the LM-from-p0 does by more than 0.1 %."""
import numpy as np
import pytest

from oracle import c_oracle
from tests import hostsim as H

SPACINGS = {
    "uniform8": np.arange(1, 9) * 10.0,
    "uniform3": np.array([10.0, 20.0, 30.0]),
    "uniform5_from0": np.arange(0, 5) * 8.0,
    "uniform16": np.arange(1, 17) * 5.0,
    "descending8": np.arange(8, 0, -1) * 10.0,
    "t1rho7": np.array([0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]),
    "mono4": np.array([10.0, 20.0, 40.0, 80.0]),
}


def _synth(x, n, sigma, t_rng, seed, sign=1.0):
    rng = np.random.default_rng(seed)
    a = rng.uniform(500, 1500, n)
    t = rng.uniform(*t_rng, n)
    return (sign * a * np.exp(-x[:, None] / t) + rng.normal(0, sigma, (len(x), n))).astype(np.float32)


def _relb(p, q):
    return np.maximum(np.abs(p[:, 1] - q[:, 1]) - 2e-6, 0) / np.abs(q[:, 1])


@pytest.mark.parametrize("name", list(SPACINGS))
@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_high_snr_same_minimiser(name, dtype):
    x = SPACINGS[name]
    for sigma, t_rng, sign in ((10.0, (10, 80), 1.0), (10.0, (200, 2000), 1.0), (0.0, (5, 300), 1.0), (10.0, (10, 80), -1.0)):
        y = _synth(x, 4001, sigma, t_rng, 3, sign)
        p0, r0, s0, i0 = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), dtype=dtype, fast=0)
        for fast in (1, 2):
            p, r, s, it = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), dtype=dtype, fast=fast)
            ok = (s0 >= 1) & (s0 <= 4)
            assert ((s >= 1) & (s <= 4))[ok].all()
            # on nearly flat signals b is determined to ~50 %: the LM's own ftol slack there is ~4e-4 of b
            tol = (1e-4 if t_rng[1] <= 300 else 5e-4) if dtype == "f32" else 1e-6
            assert _relb(p[ok], p0[ok]).max() < tol, (name, sigma, t_rng, fast)
            # a is the amplitude extrapolated to x = 0: on fast decays its error is |db| x_min, a few times b's
            assert (np.abs(p[ok, 0] - p0[ok, 0]) / np.abs(p0[ok, 0])).max() < (3e-4 if dtype == "f32" else 1e-6)
            # fp32 residuals of ~1000-valued samples carry ~1e-4 absolute rounding: r2 agrees to ~1e-4 on flat signals
            assert np.abs(r[ok] - r0[ok]).max() < (2e-4 if dtype == "f32" else 1e-7)
            if len(x) >= 4:  # the fast path is what ran: two passes instead of the LM's five or more
                assert it[ok].mean() < 0.6 * i0[ok].mean()


def test_two_voxel_and_one_voxel_solvers_agree():
    x = SPACINGS["uniform8"]
    y = _synth(x, 20001, 10.0, (10, 80), 11)
    p1, r1, s1, i1 = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=1)
    p2, r2, s2, i2 = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=2)
    # the two-voxel attempt is straight-line (exactly two passes); what it turns down (here: the ~0.4 % of the voxels
    # that want a third pass) goes through the one-voxel solver, so those voxels are bit-identical
    assert ((s1 >= 1) & (s1 <= 4)).all() and ((s2 >= 1) & (s2 <= 4)).all()
    two = i2 == 2
    assert two.mean() > 0.99 and (i1[two] <= 2).all()
    assert np.array_equal(p1[~two], p2[~two]) and np.array_equal(r1[~two], r2[~two]) and (i1[~two] == i2[~two]).all()
    assert (np.abs(p1 - p2) / np.abs(p2)).max() < 5e-6 and np.abs(r1 - r2).max() < 1e-6


@pytest.mark.parametrize("name", ["uniform8", "descending8", "t1rho7", "mono4"])
@pytest.mark.parametrize("sigma,t_rng,lim", [(100.0, (10, 80), 5e-3), (200.0, (10, 80), 1.5e-2), (10.0, (2, 10), 1.5e-2)])
def test_low_snr_against_lm_and_reference(name, sigma, t_rng, lim):
    x = SPACINGS[name]
    n = 12000
    y = _synth(x, n, sigma, t_rng, 17)
    pr, rr = c_oracle.curve_fit("monoexponential", x, y.astype(np.float64), p0=(1.0, -1 / 30))
    okr = ~np.isnan(pr[:, 0])
    sse = lambda pp, sel: ((pp[sel, 0] * np.exp(pp[sel, 1] * x[:, None]) - y[:, sel]) ** 2).sum(0)  # noqa: E731
    worse = {}
    for fast in (0, 2):
        p, r, s, it = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=fast)
        ok = (s >= 1) & (s <= 4)
        assert abs(ok.mean() - okr.mean()) < 0.02  # comparable failure rates (the sets cannot coincide)
        both = ok & okr
        worse[fast] = (sse(p, both) > sse(pr, both) * (1 + 1e-4)).mean()
        if fast == 0:
            p_lm, ok_lm = p, ok
        else:
            sel = ok & ok_lm
            assert (_relb(p[sel], p_lm[sel]) > 1e-3).mean() < lim, (name, sigma)
    # (signal gone after the first echo, T in (2, 10) ms: ~12 % of the voxels fail either way; the bound is looser)
    assert worse[2] < worse[0] + (1e-3 if t_rng[0] >= 10 else 3e-3), worse


def test_declines_to_lm_where_it_must():
    """Voxels the fast path must hand over: all-zero (skipped), NaN (flagged), y_bounds (skip rules),
    too few passes allowed; results then equal the LM's exactly."""
    x = SPACINGS["uniform8"]
    y = _synth(x, 64, 10.0, (10, 80), 5).astype(np.float64)
    y[:, 3] = 0.0
    y[:, 7] = 7.0
    y[2, 9] = np.nan
    a = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=0)
    b = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=2)
    assert b[2][3] == 0 and np.isnan(b[0][3]).all() and b[1][3] == 0.0
    assert b[2][9] == a[2][9] == 6 and np.isnan(b[0][9]).all()
    assert abs(b[0][7, 1]) < 1e-6 and abs(b[0][7, 0] - 7) < 1e-4
    a = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=0, y_bounds=(0, 1400))
    b = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=2, y_bounds=(0, 1400))
    for u, v in zip(a, b):
        assert np.array_equal(u, v, equal_nan=True)


@pytest.mark.parametrize("name", ["uniform8", "t1rho7"])
def test_invariances(name):
    """Properties of the least-squares minimiser that do not depend on the data: scaling the samples scales a
    and leaves b; stretching the echo times by c divides b by c; the result of a voxel does not depend on its
    neighbour in the lane pair (two-voxel packing) nor on its position."""
    x = SPACINGS[name]
    y = _synth(x, 6001, 10.0, (10, 80), 23)
    p, r, s, it = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), fast=2)
    assert ((s >= 1) & (s <= 4)).all()
    for c in (1e-6, 0.125, 8.0, 1000.0, 1e6):
        pc, rc, sc, ic = H.fit("monoexponential", x, y * np.float32(c), p0=(1.0, -1 / 30), fast=2)
        assert ic.mean() < 2.5  # still the fast path (its moments are scale-free), not the LM
        tol = 2e-5  # the solver's own tolerance: the start and the rounding differ with the scale
        assert (np.abs(pc[:, 0] / c - p[:, 0]) / np.abs(p[:, 0])).max() < tol
        assert (np.abs(pc[:, 1] - p[:, 1]) / np.abs(p[:, 1])).max() < tol
        if c >= 0.1:  # (r2 carries the reference's absolute eps = 1e-8 in its denominator: not scale-free for tiny signals)
            assert np.abs(rc - r).max() < 1e-5
    for c in (0.5, 4.0):
        pc, rc, sc, _ = H.fit("monoexponential", x * c, y, p0=(1.0, -1 / (30 * c)), fast=2)
        assert (np.abs(pc[:, 1] * c - p[:, 1]) / np.abs(p[:, 1])).max() < 1e-5
        assert (np.abs(pc[:, 0] - p[:, 0]) / np.abs(p[:, 0])).max() < 1e-5
    perm = np.random.default_rng(1).permutation(y.shape[1])
    pp, rp, sp, ip = H.fit("monoexponential", x, y[:, perm], p0=(1.0, -1 / 30), fast=2)
    assert np.array_equal(pp, p[perm]) and np.array_equal(rp, r[perm]) and np.array_equal(ip, it[perm])


def test_recovers_truth_over_decades_of_decay():
    """Noise-free signals from T = 0.3 dx to 3000 dx (q from 0.04 to 0.9997) and growing ones: the fast path
    (or the LM it hands over to) must return the generating parameters."""
    x = SPACINGS["uniform8"]
    rng = np.random.default_rng(9)
    n = 4000
    T = 10.0 ** rng.uniform(np.log10(3.0), np.log10(3e4), n)
    sign = np.where(rng.random(n) < 0.15, -1.0, 1.0)  # 15 % growing exponentials (b > 0), capped below
    b = -sign / np.maximum(T, np.where(sign < 0, 40.0, 0.0))
    a = rng.uniform(1, 3000, n)
    y = a * np.exp(b * x[:, None])
    for dtype, tol in (("f64", 1e-9), ("f32", 2e-4)):
        p, r, s, it = H.fit("monoexponential", x, y, p0=(1.0, -1 / 30), dtype=dtype, fast=2)
        ok = (s >= 1) & (s <= 4)
        assert ok.mean() > 0.995  # the rest: T << dx, nothing but the first echo left in fp32
        relb = np.abs(p[ok, 1] - b[ok]) / np.abs(b[ok])
        # fp32 resolves b dx to ~1e-7 absolute: slow decays lose relative precision in b accordingly
        lim = tol + (2e-7 / np.abs(b[ok] * 10.0) if dtype == "f32" else 0)
        assert (relb < lim).all(), (dtype, float((relb / lim).max()))
        assert (np.abs(p[ok, 0] - a[ok]) / a[ok] < 50 * lim).all()
        assert (r[ok] > 1 - 1e-4).all()


def test_echo_table_classification():
    """Which solver a launch gets: uniform spacing is recognised to within the arithmetic's resolution of the
    largest echo time (so that treating the spacing as uniform cannot change the model), descending order selects
    the backward start."""
    import ctypes

    lib = H._load()

    def flags(x, dtype="f32"):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib.hostsim_xtab_flags(0 if dtype == "f32" else 1, len(x), x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))

    assert flags(np.arange(1, 9) * 10.0) == 1
    assert flags(np.linspace(4.6, 32.6, 8)) == 1 and flags(np.linspace(4.6, 32.6, 8), "f64") == 1
    assert flags(np.arange(8, 0, -1) * 10.0) == 3
    assert flags([10.0, 20.0, 40.0, 80.0]) == 0
    assert flags([0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]) == 0
    assert flags([10.0, 20.0, 30.0, 40.0 + 1e-3]) == 0               # 2.5e-5 off the grid: not uniform in fp32
    assert flags([10.0, 20.0, 30.0, 40.0 + 1e-6]) == 1               # below fp32's resolution of 40
    assert flags([10.0, 20.0, 30.0, 40.0 + 1e-6], "f64") == 0        # but visible in fp64
    assert flags([10.0, 20.0]) == 0 and flags([5.0, 5.0, 5.0]) == 0  # too few echoes / zero spacing


def test_random_echo_times_fuzz():
    """Seeded fuzz over echo counts, random (non-uniform, unsorted-free) echo times and decay ranges: the general
    fast path must land on the LM's minimiser (high SNR) for every spacing, or hand the voxel over."""
    rng = np.random.default_rng(2024)
    worst = 0.0
    for trial in range(24):
        E = int(rng.choice([3, 4, 5, 7, 8, 16]))
        span = 10.0 ** rng.uniform(0.5, 2.5)                       # 3 .. 300 time units
        x = np.sort(rng.uniform(0.0, span, E))
        x += np.arange(E) * span * 0.02                             # no (near-)coincident echo times
        t_lo, t_hi = 0.15 * span, 1.5 * span
        y = _synth(x, 1500, 5.0, (t_lo, t_hi), 100 + trial)
        p0 = (1.0, -1.0 / (0.5 * span))
        pl, rl, sl, il = H.fit("monoexponential", x, y, p0=p0, fast=0)
        pf, rf, sf, itf = H.fit("monoexponential", x, y, p0=p0, fast=2)
        ok = (sl >= 1) & (sl <= 4)
        assert ok.mean() > 0.985 and ((sf >= 1) & (sf <= 4))[ok].all(), (trial, E, x)  # (SNR 5, 3 echoes: the LM fails ~1 %)
        rel = _relb(pf[ok], pl[ok])
        worst = max(worst, float(rel.max()))
        assert rel.max() < 3e-4 and np.abs(rf[ok] - rl[ok]).max() < 2e-4, (trial, E, x, float(rel.max()))
        assert itf[ok].mean() < il[ok].mean()                       # and it is the cheaper path
    assert worst < 3e-4


def test_broad_protocol_fuzz():
    """Seeded fuzz over whole protocols: 3-16 echoes, uniform / random / descending echo times with and without an
    offset, signal scales 1e-4 .. 1e5, both signs of a and b, time constants from 0.2 to 10 spans, SNR 30 .. inf and
    an optional baseline.  Wherever the fast path, the LM and a tight fp64 LM all converge, the fast path's cost must
    not exceed the tight solution's by more than 0.1 % (nor may it return anything non-finite)."""
    rng = np.random.default_rng(99)
    total = worse = 0
    for trial in range(300):
        E = int(rng.choice([3, 4, 5, 7, 8, 16]))
        span = 10.0 ** rng.uniform(0.0, 3.0)
        x0 = span * rng.uniform(0, 2) if rng.random() < 0.5 else 0.0
        if rng.random() < 0.4:
            x = x0 + np.arange(E) * span / E
        else:
            x = x0 + np.sort(rng.uniform(0.0, span, E)) + np.arange(E) * span * 0.02
        if rng.random() < 0.2:
            x = x[::-1].copy()
        n = 300
        scale = 10.0 ** rng.uniform(-4, 5)
        a = rng.uniform(0.3, 3, n) * scale * np.where(rng.random(n) < 0.2, -1, 1)
        T = 10.0 ** rng.uniform(np.log10(0.2 * span), np.log10(10 * span), n)
        sgn = np.where(rng.random(n) < 0.1, 1.0, -1.0)
        T = np.where(sgn > 0, np.maximum(T, span), T)
        snr = float(rng.choice([np.inf, 300, 100, 30]))
        clean = a * np.exp(sgn * (x[:, None] - x.min()) / T)
        sigma = 0.0 if np.isinf(snr) else scale / snr
        y = (clean + rng.normal(0, 1, clean.shape) * sigma + (rng.random() < 0.2) * 0.05 * scale).astype(np.float32)
        p0 = (1.0, -1.0 / (0.5 * span))
        pl, rl, sl, il = H.fit("monoexponential", x, y, p0=p0, fast=0)
        pf, rf, sf, itf = H.fit("monoexponential", x, y, p0=p0, fast=2)
        pd, rd, sd, itd = H.fit("monoexponential", x, y, p0=p0, fast=0, dtype="f64", ftol=1e-13)
        okf = (sf >= 1) & (sf <= 4)
        assert np.isfinite(pf[okf]).all() and np.isfinite(rf[okf]).all(), trial
        ok = okf & (sl >= 1) & (sl <= 4) & (sd >= 1) & (sd <= 4)

        def sse(p):
            with np.errstate(all="ignore"):
                return ((p[:, 0].astype(np.float64) * np.exp(p[:, 1].astype(np.float64) * x[:, None]) - y.astype(np.float64)) ** 2).sum(0)

        ys = (y.astype(np.float64) ** 2).sum(0)
        total += int(ok.sum())
        worse += int((ok & (sse(pf) > sse(pd) * (1 + 1e-3) + 1e-6 * ys)).sum())
    assert total > 70000 and worse == 0, (total, worse)
