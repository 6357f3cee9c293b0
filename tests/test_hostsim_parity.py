"""CPU tests of the *device* arithmetic: dosma_b200/csrc/lm_core.cuh and mono_fast.cuh compiled by g++ (tests/hostsim)
against the golden fixtures and the oracle.  The GPU tests repeat these through the real kernels;
this file is what keeps the solver honest in the CPU-only build container."""
import ctypes

import numpy as np
import pytest

from oracle import dosma_oracle as O
from tests import golden_util as G, hostsim as H


def _rel(p, ref, atol=0.0):
    """Relative error; `atol` absorbs parameters whose true value is ~0 (fp32 resolves b*x, not b)."""
    return np.maximum(np.abs(p - ref) - atol, 0) / np.maximum(np.abs(ref), 1e-300)


def _p0(c, P):
    p0 = G.p0_of(c)
    if p0 is None:
        return None
    if isinstance(p0, dict):
        names = ["a", "b"]
        return [p0.get(k, 1.0) for k in names]
    return p0


@pytest.mark.parametrize("name", ["curvefit_mono4_clean_f32", "curvefit_mono8_clean_f32",
                                  "curvefit_mono4_growing_f64_p0ones", "curvefit_mono4_growing_f64_tc30",
                                  "curvefit_mono4_unit_f64_p0ones", "curvefit_mono4_unit_f64_p0dict",
                                  "curvefit_linear4_f64"])
@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("fast", [0, 1, 2])  # LM from p0 / one-voxel fast path / two-voxel fast path (what the GPU runs)
def test_noise_free_rtol(name, dtype, fast):
    c = G.load(name)
    model = {"_linear": "linear"}.get(c["meta"]["func"], c["meta"]["func"])
    popt, r2, st, it = H.fit(model, c["x"], c["y"], p0=_p0(c, 2), dtype=dtype, fast=fast)
    assert ((st >= 1) & (st <= 4)).all()
    assert _rel(popt, c["popt"], atol=2e-6 if dtype == "f32" else 1e-12).max() < (1e-4 if dtype == "f32" else 1e-8)
    assert np.abs(r2 - c["r2"]).max() < (1e-5 if dtype == "f32" else 1e-9)


@pytest.mark.parametrize("name,frac", [("curvefit_mono8_snr100_f32", 2e-3), ("curvefit_mono7_t1rho_snr100_f32", 5e-3),
                                       ("curvefit_mono8_snr30_f32", 5e-2)])
@pytest.mark.parametrize("fast", [0, 1, 2])
def test_noisy_percentiles(name, frac, fast):
    c = G.load(name)
    popt, r2, st, it = H.fit("monoexponential", c["x"], c["y"], p0=(1.0, -1 / 30), dtype="f32", fast=fast)
    ok = ~np.isnan(c["popt"][:, 0]) & (st >= 1) & (st <= 4)
    assert ok.mean() > 0.999
    rel = _rel(popt[ok], c["popt"][ok]).max(axis=1)
    assert (rel > 1e-4).mean() < frac and np.median(rel) < 2e-5
    assert it[ok].mean() < 8


@pytest.mark.parametrize("name", sorted(G.BIEXP_F32_TOL))
def test_biexp_fp32_golden(name):
    """Bi-exponential LM in fp32 (the arithmetic of BASELINE config 4) against the reference's outputs on both
    bi-exponential fixtures -- tolerances stated in tests/golden_util.py."""
    c = G.load(name)
    popt, r2, st, it = H.fit("biexponential", c["x"], c["y"], p0=G.p0_of(c), dtype="f32", fast=0, init_linear=0)
    G.check_biexp_f32(name, popt, r2)


@pytest.mark.parametrize("rounds", [(1, 1), (4, 3), (2, 7)])
def test_lm_in_rounds_is_lm_solve_bit_for_bit(rounds):
    """fit_kernel_lmq runs the LM in rounds -- lm_begin, then lm_iterate with a budget of evaluations, the solver state
    parked between rounds (csrc/lmq_kernel.cuh).  However the trips are cut, every output must equal the uncut
    iteration's bit for bit (reference: the same code with a budget that never suspends -- what the GPU test compares
    too), and lm_solve's -- another instantiation of the same source, in which the compiler may contract other
    multiply-adds -- to rounding: bi-exponential (noisy and clean fixture, incl. voxels that run into maxfev),
    mono-exponential LM with the projection start, the linear model, a tiny maxfev."""
    cases = []
    for name in sorted(G.BIEXP_F32_TOL):
        c = G.load(name)
        cases.append(("biexponential", c["x"], c["y"], dict(p0=G.p0_of(c), fast=0, init_linear=0)))
        cases.append(("biexponential", c["x"], c["y"], dict(p0=G.p0_of(c), fast=0, init_linear=1, maxfev=30)))
    c = G.load("curvefit_mono8_snr5_f32")
    cases.append(("monoexponential", c["x"], c["y"], dict(p0=(1.0, -1 / 30), fast=0)))
    cases.append(("monoexponential", c["x"], c["y"], dict(p0=(1.0, -1 / 30), fast=0, maxfev=7)))
    c = G.load("curvefit_linear4_f64")
    cases.append(("linear", c["x"], c["y"], dict(p0=G.p0_of(c), fast=0)))
    cut = 0
    try:
        for model, x, y, kw in cases:
            for dtype in ("f32", "f64"):
                H.set_rounds(0)
                solve = H.fit(model, x, y, dtype=dtype, **kw)
                H.set_rounds(1 << 30)
                ref = H.fit(model, x, y, dtype=dtype, **kw)
                H.set_rounds(*rounds)
                out = H.fit(model, x, y, dtype=dtype, **kw)
                for a, b in zip(ref, out):
                    assert np.array_equal(a, b, equal_nan=True), (model, dtype, kw)
                cut = max(cut, int(ref[3].max()) - rounds[0] - 1)
                # lm_solve: same statuses and pass counts on (nearly) every voxel, same minimiser
                assert (solve[2] != ref[2]).mean() < 5e-3 and (solve[3] != ref[3]).mean() < 2e-2, (model, dtype, kw)
                ok = (solve[2] >= 1) & (solve[2] <= 4) & (ref[2] >= 1) & (ref[2] <= 4)
                if not ok.any():
                    continue
                tol = 1e-3 if dtype == "f32" else 1e-9
                assert np.percentile(_rel(ref[0][ok], solve[0][ok]).max(axis=1), 99) < tol
                assert np.percentile(np.abs(ref[1][ok] - solve[1][ok]), 99) < (1e-5 if dtype == "f32" else 1e-10)
    finally:
        H.set_rounds(0)
    assert cut > 10  # (the budgets did cut fits into several rounds)


def test_uniform_recurrence_form_of_the_lm():
    """On uniformly spaced echoes fit_kernel_lmq takes the model's exponentials from a two-echo recurrence
    (exp(b x_k+2) = exp(b x_k) * exp(2 b dx): MonoExp / BiExp ::Rec in lm_core.cuh) instead of one MUFU.EX2 per echo.
    That perturbs the model like a relative change of b by a few 1e-7: the bi-exponential fixtures must hold their
    reference tolerances, and the result must agree with the direct form far inside them."""
    try:
        for name in sorted(G.BIEXP_F32_TOL):
            c = G.load(name)
            kw = dict(p0=G.p0_of(c), dtype="f32", fast=0, init_linear=0)
            H.set_rounds(0)
            ref = H.fit("biexponential", c["x"], c["y"], **kw)
            H.set_rounds(5, 2)
            H.set_uniform_recurrence(1)
            out = H.fit("biexponential", c["x"], c["y"], **kw)
            H.set_uniform_recurrence(0)
            G.check_biexp_f32(name, out[0], out[1])
            ok = ~np.isnan(ref[0][:, 0]) & ~np.isnan(out[0][:, 0])
            assert (np.isnan(ref[0][:, 0]) ^ np.isnan(out[0][:, 0])).mean() < 5e-3
            rel = _rel(out[0][ok], ref[0][ok]).max(axis=1)
            assert np.median(rel) < 5e-6 and np.percentile(rel, 99) < 5e-4, (np.median(rel), np.percentile(rel, 99))
            assert np.abs(out[1][ok] - ref[1][ok]).max() < 1e-6
        for name, odd in (("curvefit_mono8_snr100_f32", False), ("curvefit_mono7_t1rho_snr100_f32", True)):
            c = G.load(name)  # (the 7-echo T1rho protocol is not uniformly spaced: the recurrence must not be taken)
            H.set_rounds(6, 6)
            ref = H.fit("monoexponential", c["x"], c["y"], p0=(1.0, -1 / 30), fast=0)
            H.set_rounds(6, 6)
            H.set_uniform_recurrence(1)
            out = H.fit("monoexponential", c["x"], c["y"], p0=(1.0, -1 / 30), fast=0)
            H.set_uniform_recurrence(0)
            if odd:
                assert np.array_equal(ref[0], out[0], equal_nan=True)
            else:
                assert np.nanmax(_rel(out[0], ref[0])) < 2e-6 and not np.array_equal(ref[0], out[0])
    finally:
        H.set_uniform_recurrence(0)
        H.set_rounds(0)


def test_degenerate_and_bounds():
    c = G.load("curvefit_mono8_degenerate_f32")
    popt, r2, st, it = H.fit("monoexponential", c["x"], c["y"], p0=(1.0, -1 / 30))
    assert (st[:16] == 0).all() and np.isnan(popt[:16]).all() and (r2[:16] == 0).all()
    c = G.load("curvefit_mono8_ybounds")
    popt, r2, st, it = H.fit("monoexponential", c["x"], c["y"], p0=(1.0, -1 / 30), y_bounds=(0, 1400))
    assert G.same_nan(popt, c["popt"])


def test_loglinear_init_matches_reference_semantics():
    """In-kernel log-linear p0 equals the oracle's polyfit p0 (fitting.py:701-718), including the
    zero -> 1e-10 and negative -> (1, 0) rules; the converged fit is then the same."""
    c = G.load("monoexpfit_polyfit_zeros_negatives")
    x, y = c["x"], c["y"].astype(np.float64)
    popt, r2, st, it = H.fit("monoexponential", x, y, dtype="f64", init_mode=1, maxfev=2)  # no LM budget
    p0 = O.loglinear_p0(x, y)
    # with no iterations left the solver reports failure; re-run with budget and compare the result
    popt, r2, st, it = H.fit("monoexponential", x, y, dtype="f64", init_mode=1)
    tc = O.process_params(popt.copy(), r2, out_ufuncs=(None, lambda v: 1 / np.abs(v)),
                          out_bounds=((-np.inf, np.inf), (0, 100)), r2_threshold=0.9, nan_to_num=0.0)[:, 1]
    ref = c["tc"].reshape(-1)
    agree = np.abs(np.around(tc, 3) - ref) <= 1.001e-3
    assert agree.mean() > 0.99, agree.mean()
    assert p0.shape == (y.shape[1], 2)


def test_post_param_matches_process_params():
    rng = np.random.default_rng(0)
    lib = H._load()
    v = np.concatenate([rng.normal(0, 0.05, 500), [0.0, np.nan, np.inf, -np.inf, 1e-3, -1e-3]])
    r2 = rng.uniform(0.5, 1.0, v.size)
    ref = O.process_params(np.stack([v, v], axis=1), r2, out_ufuncs=(None, lambda t: 1 / np.abs(t)),
                           out_bounds=((-np.inf, np.inf), (0, 100)), r2_threshold=0.9, nan_to_num=0.0)
    ref[:, 1] = np.around(ref[:, 1], 3)
    I4, D4 = ctypes.c_int * 4, ctypes.c_double * 4
    uf, dec = I4(0, 1, 0, 0), I4(-1, 3, -1, -1)
    lb, ub = D4(-np.inf, 0, -np.inf, -np.inf), D4(np.inf, 100, np.inf, np.inf)
    for i in (0, 1):
        got = np.array([lib.hostsim_post_param(1, uf, lb, ub, 1, ctypes.c_double(0.9), 1, ctypes.c_double(0.0), dec, i,
                                               ctypes.c_double(a), ctypes.c_double(b)) for a, b in zip(v, r2)])
        assert np.array_equal(got, ref[:, i])


def test_fp32_epilogue_takes_the_float64_decisions():
    """post_param_f32 (fp32 where the epilogue is comparisons only) must return exactly what the float64
    epilogue returns after conversion to float32, including at bounds / thresholds that are not representable
    in float32 and on NaN / inf inputs."""
    lib = H._load()
    rng = np.random.default_rng(0)
    n = 200000
    v = rng.uniform(-5, 130, n).astype(np.float32)
    r2 = rng.uniform(0.5, 1.0, n).astype(np.float32)
    # values sitting exactly on / next to the float32 neighbours of the bounds and of the threshold
    lb, ub, thr = 0.1, 100.00000123, 0.9
    for k, x0 in enumerate((lb, ub)):
        f = np.float32(x0)
        v[k * 3: k * 3 + 3] = [np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))]
    t = np.float32(thr)
    r2[10:13] = [np.nextafter(t, np.float32(0)), t, np.nextafter(t, np.float32(2))]
    v[20:24] = [np.nan, np.inf, -np.inf, 0.0]
    ip = ctypes.POINTER(ctypes.c_int)
    dp = ctypes.POINTER(ctypes.c_double)
    fp = ctypes.POINTER(ctypes.c_float)
    for ufunc, decimals, fill in ((0, -1, 0.0), (0, -1, None), (0, -1, 0.1), (1, 3, 0.0), (0, 2, 0.0)):
        uf = (ctypes.c_int * 4)(ufunc, ufunc, 0, 0)
        dec = (ctypes.c_int * 4)(decimals, decimals, -1, -1)
        lbs = (ctypes.c_double * 4)(lb, lb, -np.inf, -np.inf)
        ubs = (ctypes.c_double * 4)(ub, ub, np.inf, np.inf)
        a = np.empty(n, np.float32)
        b = np.empty(n, np.float32)
        lib.hostsim_post_param_f32(1, uf, lbs, ubs, 1, ctypes.c_double(thr), int(fill is not None),
                                   ctypes.c_double(fill or 0.0), dec, 0, ctypes.c_int64(n), v.ctypes.data_as(fp),
                                   r2.ctypes.data_as(fp), a.ctypes.data_as(fp), b.ctypes.data_as(fp))
        assert np.array_equal(a, b, equal_nan=True), (ufunc, decimals, fill)


@pytest.mark.parametrize("decimals", [-1, 0, 1, 3, 6])
@pytest.mark.parametrize("fill", [None, 0.0, 0.05])
@pytest.mark.parametrize("pairs", [0, 1])  # one-voxel form (post_param_f32) / packed two-voxel form (post_pair_f32)
def test_fp32_inverse_epilogue_is_the_float64_one(decimals, fill, pairs):
    """The MonoExponentialFit column in fp32 (1 / |v|, bounds, r2 threshold, fill, rounding: post_param_f32's
    `fastinv` plan) against the float64 epilogue converted to float, bit for bit -- on random values and on
    adversarial ones: |v| whose reciprocal sits next to a rounding tie (10^d / (k + 1/2) and its float
    neighbours), next to the bounds, exact ties, zeros, infinities, NaN, tiny and huge magnitudes."""
    lib = H._load()
    lib.hostsim_set_pairs(pairs)
    rng = np.random.default_rng(1)
    S = 10.0 ** max(decimals, 0)
    lb, ub, thr = 0.0, 100.0, 0.9
    parts = [rng.uniform(-0.3, 0.3, 100000), rng.uniform(0.005, 0.02, 50000) * rng.choice([-1, 1], 50000),
             10.0 ** rng.uniform(-38, 38, 20000)]
    k = np.arange(0, 20000, dtype=np.float64)
    ties = (S / (k + 0.5)).astype(np.float32)
    for d in (-2, -1, 0, 1, 2):
        t = ties.copy()
        for _ in range(abs(d)):
            t = np.nextafter(t, np.float32(np.inf if d > 0 else -np.inf))
        parts.append(t.astype(np.float64))
    for edge in (1.0 / ub, 20.0, 4.0, 2000.0, 16.0, 0.25):
        f = np.float32(edge)
        parts.append(np.array([np.nextafter(f, np.float32(0)), f, np.nextafter(f, np.float32(np.inf))], dtype=np.float64))
    parts.append(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3e38, -3e38]))
    v = np.concatenate(parts).astype(np.float32)
    n = v.size
    r2 = rng.uniform(0.8, 1.0, n).astype(np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    uf = (ctypes.c_int * 4)(1, 1, 0, 0)
    dec = (ctypes.c_int * 4)(decimals, decimals, -1, -1)
    for lbs_, ubs_ in (((lb,) * 2, (ub,) * 2), ((-np.inf,) * 2, (np.inf,) * 2), ((0.1,) * 2, (37.123456789,) * 2)):
        lbs = (ctypes.c_double * 4)(*lbs_, -np.inf, -np.inf)
        ubs = (ctypes.c_double * 4)(*ubs_, np.inf, np.inf)
        a = np.empty(n, np.float32)
        b = np.empty(n, np.float32)
        lib.hostsim_post_param_f32(1, uf, lbs, ubs, 1, ctypes.c_double(thr), int(fill is not None),
                                   ctypes.c_double(fill or 0.0), dec, 1, ctypes.c_int64(n), v.ctypes.data_as(fp),
                                   r2.ctypes.data_as(fp), a.ctypes.data_as(fp), b.ctypes.data_as(fp))
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        assert not bad.any(), (decimals, fill, lbs_, v[bad][:5], a[bad][:5], b[bad][:5])
    lib.hostsim_set_pairs(0)


def test_monoexpfit_epilogue_gets_the_fp32_plan():
    """The epilogue MonoExponentialFit asks for (fitting.py:722-737: 1/|b| within (0, 100), fill 0, one decimal) must
    be served by the fp32 forms, not by the float64 fallback."""
    lib = H._load()
    I4, D4 = ctypes.c_int * 4, ctypes.c_double * 4
    uf, dec = I4(0, 1, 0, 0), I4(-1, 1, -1, -1)
    lb, ub = D4(-np.inf, 0.0, -np.inf, -np.inf), D4(np.inf, 100.0, np.inf, np.inf)
    assert lib.hostsim_post_plan(uf, lb, ub, 1, ctypes.c_double(0.0), dec, 0) == 1  # a: comparisons only
    assert lib.hostsim_post_plan(uf, lb, ub, 1, ctypes.c_double(0.0), dec, 1) == 2  # tc: fp32 two-float form
    ub2 = D4(np.inf, np.inf, np.inf, np.inf)  # unbounded tc: 10^d / x is unbounded too -> float64 form
    assert lib.hostsim_post_plan(uf, lb, ub2, 1, ctypes.c_double(0.0), dec, 1) == 0


@pytest.mark.parametrize("name", G.names("monoexpfit_"))
@pytest.mark.parametrize("fast", [0, 2])
def test_monoexpfit_chain_on_the_host_build(name, fast):
    """The whole device chain of `MonoExponentialFit.fit` -- fit (LM or two-voxel fast path), r2, fused epilogue
    (1/|b|, bounds, r2 threshold, fill 0, rounding), mask fill -- compiled by g++, against the golden maps
    generated by the real reference (same acceptance as the GPU test `test_monoexpfit_golden`)."""
    c = G.load(name)
    m = c["meta"]
    y = c["y"].astype(np.float32) if c["y"].dtype != np.float64 else c["y"]
    init_mode = 1 if m["tc0"] == "polyfit" else 0
    p0 = (1.0, -1.0 / float(m["tc0"])) if init_mode == 0 else (1.0, 1.0)
    popt, r2, st, it = H.fit("monoexponential", c["x"], y, p0=p0, dtype="f32", init_mode=init_mode, fast=fast)
    lib = H._load()
    ip = lambda a: (ctypes.c_int * 4)(*a)  # noqa: E731
    dp = lambda a: (ctypes.c_double * 4)(*a)  # noqa: E731
    lb, ub = float(m["bounds"][0]), float(m["bounds"][1])
    tc = np.array([lib.hostsim_post_param(1, ip([0, 1, 0, 0]), dp([-np.inf, lb, -np.inf, -np.inf]),
                                          dp([np.inf, ub, np.inf, np.inf]), 1, ctypes.c_double(m["r2_threshold"]), 1,
                                          ctypes.c_double(0.0), ip([-1, m["decimal_precision"], -1, -1]), 1,
                                          ctypes.c_double(float(np.float32(b))), ctypes.c_double(float(np.float32(r))))
                   for b, r in zip(popt[:, 1], r2)])
    r2 = r2.copy()
    ref_tc, ref_r2 = c["tc"].reshape(-1), c["r2"].reshape(-1)
    if m["use_mask"]:
        mask = c["mask"].reshape(-1)
        tc[~mask] = 0.0
        r2[~mask] = 0.0
    step = 10.0 ** (-m["decimal_precision"])
    zeroed = (tc == 0) != (ref_tc == 0)
    assert zeroed.mean() < (0.03 if "snr30" in name else 0.01), zeroed.mean()
    close = np.abs(tc - ref_tc)[~zeroed] <= 1.001 * step
    need = 0.95 if "snr30" in name else (0.99 if "zeros_negatives" in name else 0.995)
    assert close.mean() > need, close.mean()
    kept = (tc != 0) & (ref_tc != 0)
    assert np.quantile(np.abs(r2 - ref_r2)[kept], 0.99) < 1e-4
