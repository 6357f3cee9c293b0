"""GPU parity tests: the CUDA path (through the C-ABI) against (a) the golden fixtures generated
from the real reference code and (b) the live CPU oracle on fresh seeded inputs.

Parity definition: the reference stops MINPACK at ftol = 1e-5, which on noise-free /
high-SNR data is the least-squares minimiser to ~1e-16 / ~1e-5.  The engine converges to the same
minimiser (tighter), so
  * "at convergence" sets (noise-free): popt within rtol 1e-4 (north-star tolerance) -- in fact 1e-5;
  * SNR 100: >= 99.9 % of voxels within 1e-4, median < 1e-6;
  * lower SNR: the disagreement is the reference's own early-stop slack; checked by percentiles and
    against the oracle re-run to tight tolerance.
r2 is compared with atol 1e-5 (fp32 arithmetic) / 1e-9 (fp64).
"""
import warnings

import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import dosma_b200

    return dosma_b200


def _func(D, name):
    return {"monoexponential": D.monoexponential, "biexponential": D.biexponential, "_linear": D.linear,
            "linear": D.linear}[name]


def _rel(p, ref, atol=0.0):
    """Relative error; `atol` absorbs parameters whose true value is ~0 (fp32 resolves b*x, not b)."""
    return np.maximum(np.abs(p - ref) - atol, 0) / np.maximum(np.abs(ref), 1e-300)


CLEAN = ["curvefit_mono4_clean_f32", "curvefit_mono8_clean_f32", "curvefit_mono4_growing_f64_p0ones",
         "curvefit_mono4_growing_f64_tc30", "curvefit_mono4_unit_f64_p0ones", "curvefit_mono4_unit_f64_p0dict",
         "curvefit_linear4_f64"]


@pytest.mark.parametrize("name", CLEAN)
@pytest.mark.parametrize("cd", ["f32", "f64"])
def test_noise_free_golden_rtol_1e4(D, name, cd):
    c = G.load(name)
    popt, r2, st = D.curve_fit(_func(D, c["meta"]["func"]), c["x"], c["y"], p0=G.p0_of(c), compute_dtype=cd,
                               return_stats=True)
    assert popt.dtype == np.float64 and popt.shape == c["popt"].shape
    assert st["n_failed"] == 0 and not np.isnan(popt).any()
    rel = _rel(popt, c["popt"], atol=2e-6 if cd == "f32" else 1e-12)
    assert rel.max() < (1e-4 if cd == "f32" else 1e-8), rel.max()
    assert np.abs(r2 - c["r2"]).max() < (1e-5 if cd == "f32" else 1e-9)


@pytest.mark.parametrize("name,frac_limit,median_limit", [
    ("curvefit_mono8_snr100_f32", 2e-3, 2e-6),
    ("curvefit_mono7_t1rho_snr100_f32", 5e-3, 2e-6),
    ("curvefit_mono8_snr100_p0voxel", 5e-3, 2e-6),
    ("curvefit_mono8_snr30_f32", 5e-2, 2e-5),
])
def test_noisy_golden_percentiles(D, name, frac_limit, median_limit):
    c = G.load(name)
    popt, r2 = D.curve_fit(D.monoexponential, c["x"], c["y"], p0=G.p0_of(c))
    ok = ~np.isnan(c["popt"][:, 0]) & ~np.isnan(popt[:, 0])
    assert ok.mean() > 0.999
    rel = _rel(popt[ok], c["popt"][ok]).max(axis=1)
    assert np.median(rel) < median_limit, np.median(rel)
    assert (rel > 1e-4).mean() < frac_limit, (rel > 1e-4).mean()
    assert np.abs(r2[ok] - c["r2"][ok]).max() < 1e-4


def test_noisy_against_tight_oracle(D):
    """Separate the reference's early-stop slack from engine error: compare with SciPy re-run to
    ftol = xtol = 1e-15 from the reference's own solution."""
    from scipy.optimize import curve_fit as scf

    c = G.load("curvefit_mono8_snr30_f32")
    x, y = c["x"], c["y"].astype(np.float64)
    popt, _ = D.curve_fit(D.monoexponential, x, c["y"], p0=(1.0, -1 / 30))
    n = 400
    tight = np.empty((n, 2))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            tight[i] = scf(lambda t, a, b: a * np.exp(b * t), x, y[:, i], p0=c["popt"][i], ftol=1e-15, xtol=1e-15,
                           maxfev=2000)[0]
    rel = _rel(popt[:n], tight).max(axis=1)
    assert np.quantile(rel, 0.99) < 5e-5 and rel.max() < 1e-3, (np.quantile(rel, 0.99), rel.max())
    ref_rel = _rel(c["popt"][:n], tight).max(axis=1)
    assert np.quantile(rel, 0.99) < np.quantile(ref_rel, 0.99)  # engine is closer to the minimiser than MINPACK


def test_degenerate_voxels(D):
    """fitting.py:1065-1067: all-zero -> NaN, r2 = 0.  Other rows match the reference where it converged."""
    c = G.load("curvefit_mono8_degenerate_f32")
    popt, r2 = D.curve_fit(D.monoexponential, c["x"], c["y"], p0=(1.0, -1 / 30))
    assert np.isnan(popt[:16]).all() and (r2[:16] == 0).all()
    assert np.isnan(c["popt"][:16]).all()
    # constant rows: a = 100, b = 0 (absolute tolerance: b's true value is 0)
    assert np.abs(popt[48:64, 0] - 100).max() < 1e-2 and np.abs(popt[48:64, 1]).max() < 1e-5
    both = ~np.isnan(popt[:, 0]) & ~np.isnan(c["popt"][:, 0])
    both[:64] = False
    rel = _rel(popt[both], c["popt"][both]).max(axis=1)
    assert np.median(rel) < 2e-5


def test_y_bounds_skip(D):
    c = G.load("curvefit_mono8_ybounds")
    with pytest.warns(UserWarning):
        popt, r2 = D.curve_fit(D.monoexponential, c["x"], c["y"], p0=(1.0, -1 / 30), y_bounds=(0, 1400))
    assert G.same_nan(popt, c["popt"])
    assert np.array_equal(r2 == 0, c["r2"] == 0)
    ok = ~np.isnan(popt[:, 0])
    assert (_rel(popt[ok], c["popt"][ok]).max(axis=1) > 1e-4).mean() < 5e-3


def test_low_snr_failure_rate(D):
    """Failure *sets* cannot coincide (different algorithms); the rates must be comparable and every
    failed voxel must read NaN / r2 = 0 (fitting.py:1069-1073)."""
    c = G.load("curvefit_mono8_snr5_f32")
    popt, r2 = D.curve_fit(D.monoexponential, c["x"], c["y"], p0=(1.0, -1 / 30))
    fail = np.isnan(popt[:, 0])
    ref_fail = np.isnan(c["popt"][:, 0])
    assert np.isnan(popt[fail]).all() and (r2[fail] == 0).all()
    # measured on this fixture (SNR 5, 2048 voxels): reference 1.4 % failures, engine 2.0 % (2.2 % with the LM only),
    # symmetric difference of the two sets 1.0 % -- bounded at a few voxels above that, not at percentage points
    assert abs(fail.mean() - ref_fail.mean()) < 0.012, (fail.mean(), ref_fail.mean())
    assert (fail ^ ref_fail).mean() < 0.02, (fail ^ ref_fail).mean()


@pytest.mark.parametrize("post", [False, True])
def test_lm_tail_of_the_dense_kernel(D, post, monkeypatch):
    """Dense mono-exponential fit of a volume that is half noise (air): the two-voxel TMA kernel hands the voxels that
    neither its straight-line fit nor the Newton loop settle to the LM-in-rounds kernel through a device list (the LM
    tail, csrc/mono2_kernels.cuh::fit_deferred), launched right behind it.  Same results as with the LM run inside
    the kernel (DFIT_LM_TAIL=0) up to the two kernels' rounding, same counts; repeated launches alternate the list's
    two counters; every voxel is written exactly once; and the LM voxels equal what the LM alone (fast_path=0) finds."""
    import torch

    from dosma_b200 import _cabi, device_api as A

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(21)
    n = 288 * 4096 + 128  # TMA-aligned pitch; above the 2^20 voxels from which a launch uses the LM tail
    x = [10.0 * i for i in range(1, 9)]
    xt = torch.tensor(x, device=dev)[:, None]
    air = (torch.rand(n // 1024 + 1, device=dev, generator=g) < 0.5).repeat_interleave(1024)[:n]
    sig = (500 + 1000 * torch.rand(n, device=dev, generator=g)) * torch.exp(-xt / (10 + 70 * torch.rand(n, device=dev, generator=g)))
    y = torch.where(air, torch.zeros((), device=dev), sig) + 10 * torch.randn(8, n, device=dev, generator=g)
    kw = dict(post=dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, 3], r2_threshold=0.9,
                        nan_to_num=0.0)) if post else {}
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), **kw)

    def run(opts=o):
        popt = torch.full((n, 2), -7.0, device=dev)
        r2 = torch.full((n,), -7.0, device=dev)
        A.fit_device(opts, P, x, y, popt=popt, r2=r2)
        torch.cuda.synchronize()
        return popt, r2, _cabi.get_handle(0).stats()

    monkeypatch.setenv("DFIT_LM_TAIL", "0")
    ref = run()
    assert ref[2]["n_launches"] == 1 and ref[2]["n_deferred"] > 0.4 * n
    monkeypatch.setenv("DFIT_LM_TAIL", "1")
    outs = [run() for _ in range(3)]  # (three launches: both counters of the list are used, each starts from zero)
    for out in outs:
        assert out[2]["n_launches"] == 2
        assert not (out[0] == -7.0).any() and not (out[1] == -7.0).any()
        for k in ("n_fitted", "n_deferred"):
            assert out[2][k] == ref[2][k], (k, out[2], ref[2])
        assert abs(out[2]["n_failed"] - ref[2]["n_failed"]) <= 2e-3 * n
        assert torch.equal(out[0].view(torch.int32), outs[0][0].view(torch.int32)) and torch.equal(out[1], outs[0][1])
        same = (out[0] == ref[0]).all(dim=1) | (torch.isnan(out[0]).all(dim=1) & torch.isnan(ref[0]).all(dim=1))
        assert same[~air].float().mean() > 0.995  # tissue: the same kernel code fitted it
        # air: the LM from p0 on pure noise is chaotic -- the two kernels are different compilations of the solver, and
        # a last-bit difference can end in another local minimum or in maxfev -- so the bulk must agree, not every voxel
        nan_o, nan_r = torch.isnan(out[0][:, 0]), torch.isnan(ref[0][:, 0])
        assert (nan_o ^ nan_r).float().mean() < 0.02
        both = ~nan_o & ~nan_r
        if not post:
            assert ((out[1][both] - ref[1][both]).abs() < 1e-4).float().mean() > 0.97
            rel = ((out[0][both] - ref[0][both]).abs() / ref[0][both].abs()).max(dim=1).values
            assert (rel < 1e-3).float().mean() > 0.95
        else:
            assert ((out[0][both, 1] - ref[0][both, 1]).abs() <= 1.001e-3).float().mean() > 0.99  # one rounding step
    if not post:
        o_lm, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), fast_path=0)
        lm = run(o_lm)  # the LM alone, through the same rounds kernel: the air voxels' results bit for bit
        both = air & ~torch.isnan(outs[0][0][:, 0]) & ~torch.isnan(lm[0][:, 0])
        assert both.float().sum() > 0.3 * n
        same_lm = (outs[0][0][both] == lm[0][both]).all(dim=1).float().mean()
        assert same_lm > 0.6, same_lm  # (the rest: air voxels the Newton loop settled before the list)


@pytest.mark.parametrize("name", ["curvefit_biexp16_clean_f32"])
def test_biexp_noise_free(D, name):
    c = G.load(name)
    popt, r2 = D.curve_fit(D.biexponential, c["x"], c["y"], p0=G.p0_of(c), compute_dtype="f64")
    ok = ~np.isnan(popt[:, 0]) & ~np.isnan(c["popt"][:, 0])
    assert ok.mean() > 0.98
    rel = _rel(popt[ok], c["popt"][ok]).max(axis=1)
    # the data are float32-rounded, so the "noise-free" problem has a 6e-8 relative perturbation that the
    # ill-conditioned 4-parameter fit amplifies; both solvers sit at the same minimiser to ~1e-6
    assert (rel < 1e-4).mean() > 0.99, (rel < 1e-4).mean()
    assert np.abs(r2[ok] - c["r2"][ok]).max() < 1e-7


@pytest.mark.parametrize("name", sorted(G.BIEXP_F32_TOL))
def test_biexp_fp32_golden(D, name):
    """Bi-exponential fit in fp32 arithmetic (what BASELINE config 4 runs) against the reference's outputs on the
    noise-free and the SNR-100 fixture: percentile tolerances on popt, r2 and the NaN set as stated in
    tests/golden_util.py (the same bounds hold the host build of the solver in the CPU suite)."""
    c = G.load(name)
    popt, r2 = D.curve_fit(D.biexponential, c["x"], c["y"], p0=G.p0_of(c), compute_dtype="f32")
    G.check_biexp_f32(name, popt, r2)


@pytest.mark.parametrize("case", ["biexp16", "biexp16_mask", "biexp9_ragged", "biexp7_nonuniform", "mono8_lm", "mono8_ybounds",
                                  "linear4"])
def test_lm_in_rounds_kernel(D, case, monkeypatch):
    """fit_kernel_lmq (LM in rounds, suspended fits parked on a per-warp stack in shared memory; csrc/lmq_kernel.cuh):
      * however the rounds are cut -- budgets of 1, 2, 5, 9 or 'never suspend' -- popt, r2, status and pass counts are
        bit-identical: a round boundary only decides which lane evaluates the next trial point;
      * against the plain one-voxel-per-lane kernel (DFIT_LMQ=0; a different compilation of the same source, so ptxas
        contracts different multiply-adds and the last bits differ) and against the exponentials-by-MUFU form of itself
        (DFIT_LMQ_UNI=0): same minimiser to fp32 resolution on (nearly) every voxel, same failure statistics.
    Dense and masked launches, ragged sizes, skipped voxels, y_bounds."""
    import torch

    from dosma_b200 import _cabi, device_api as A

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11)
    kw, mask = {}, None
    if case.startswith("biexp"):
        E = int("".join(ch for ch in case.split("_")[0] if ch.isdigit()))
        n = 40000 if "ragged" not in case else 32 * 77 + 13
        x = [5.0 * i for i in range(1, E + 1)]
        if "nonuniform" in case:
            x = [4.0, 9.0, 15.0, 24.0, 36.0, 55.0, 80.0]
        xt = torch.tensor(x, device=dev)[:, None]
        amp = 500 + 1000 * torch.rand(n, device=dev, generator=g)
        fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
        y = amp * fs * torch.exp(-xt / (8 + 12 * torch.rand(n, device=dev, generator=g))) + \
            amp * (1 - fs) * torch.exp(-xt / (50 + 50 * torch.rand(n, device=dev, generator=g))) + \
            10 * torch.randn(E, n, device=dev, generator=g)
        model, p0 = D.biexponential, (500.0, -1 / 10, 500.0, -1 / 60)
        if "mask" in case:
            mask = (torch.rand(n, device=dev, generator=g) < 0.3).to(torch.uint8)
    elif case.startswith("mono8"):
        n = 50001
        x = [10.0 * i for i in range(1, 9)]
        xt = torch.tensor(x, device=dev)[:, None]
        y = 1000 * torch.exp(-xt / (10 + 70 * torch.rand(n, device=dev, generator=g))) + 100 * torch.randn(8, n, device=dev, generator=g)
        y[:, ::97] = 0  # skipped voxels
        model, p0 = D.monoexponential, (1.0, -1 / 30)
        kw = dict(fast_path=0) if case.startswith("mono8_lm") else dict(y_bounds=(-150.0, 1100.0))
    else:
        n = 5000
        x = [1.0, 2.0, 3.0, 4.0]
        xt = torch.tensor(x, device=dev)[:, None]
        y = xt * (1 + torch.rand(n, device=dev, generator=g)) + 0.1 * torch.randn(4, n, device=dev, generator=g)
        model, p0 = D.linear, (1.0,)
    o, P = A.make_opts(model, p0=p0, compute_dtype="f32", **kw)

    def run():
        popt = torch.full((n, P), -7.0, device=dev)
        r2 = torch.full((n,), -7.0, device=dev)
        st = torch.full((n,), 99, device=dev, dtype=torch.uint8)
        it = torch.full((n,), 99, device=dev, dtype=torch.uint8)
        A.fit_device(o, P, x, y, mask=mask, popt=popt, r2=r2, status=st, niter=it)
        torch.cuda.synchronize()
        return popt, r2, st, it, _cabi.get_handle(0).stats()

    def bits(t):
        return t if t.dtype == torch.uint8 else t.view(torch.int32)

    def close(ref, out, what):
        fitted = (ref[2] >= 1) & (ref[2] <= 4) & (out[2] >= 1) & (out[2] <= 4)
        assert ((ref[2] >= 1) & (ref[2] <= 4)).sum() > 0.5 * (n if mask is None else int(mask.sum()))
        assert (ref[2] != out[2])[(ref[2] == 0) | (out[2] == 0)].sum() == 0, what  # skipped voxels are the same voxels
        assert (fitted != ((ref[2] >= 1) & (ref[2] <= 4))).float().mean() < 2e-3, what   # success / failure flips: rare
        assert (ref[1][fitted] - out[1][fitted]).abs().max() < 2e-5, what               # the same minimum
        rel = ((ref[0][fitted] - out[0][fitted]).abs() / ref[0][fitted].abs()).max(dim=1).values
        assert float(torch.quantile(rel, 0.5)) < 2e-5 and float(torch.quantile(rel, 0.99)) < 5e-3, (what, torch.quantile(rel, 0.99))
        assert abs(ref[4]["sum_iters"] - out[4]["sum_iters"]) < 0.01 * ref[4]["sum_iters"], what

    monkeypatch.setenv("DFIT_LMQ", "5,2")
    ref = run()
    assert ref[4]["n_fitted"] == (n if mask is None else int(mask.sum())) - (0 if not case.startswith("mono8") else int((y == 0).all(dim=0).sum()) + ref[4]["n_oob"])
    for budgets in ("1,1", "2,9", "1000000,1000000"):
        monkeypatch.setenv("DFIT_LMQ", budgets)
        out = run()
        for a, b in zip(ref[:4], out[:4]):
            assert torch.equal(bits(a), bits(b)), (case, budgets)
        for k in ("n_fitted", "n_failed", "sum_iters", "max_iters", "n_oob"):
            assert ref[4][k] == out[4][k], (k, ref[4], out[4])
    assert case == "linear4" or ref[4]["max_iters"] > 5
    monkeypatch.setenv("DFIT_LMQ", "0")
    close(ref, run(), "plain kernel")
    monkeypatch.setenv("DFIT_LMQ", "5,2")
    monkeypatch.setenv("DFIT_LMQ_UNI", "0")
    close(ref, run(), "exponentials by MUFU")


@pytest.mark.parametrize("variant", ["echo_fastest", "int16", "p0_voxel", "f64_maps_status"])
def test_lm_in_rounds_kernel_input_forms(D, variant, monkeypatch):
    """The rounds kernel behind every way samples and initial guesses reach it -- echo-fastest layout, int16 samples,
    a per-voxel initial guess, float64 result maps with status / pass-count bytes -- against the plain kernel
    (DFIT_LMQ=0): same fits to rounding, same skipped voxels."""
    import torch

    from dosma_b200 import device_api as A

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(17)
    n, E = 30000 + 7, 12
    x = [6.0 * i for i in range(1, E + 1)]
    xt = torch.tensor(x, device=dev)[:, None]
    amp = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
    ts, tl = 8 + 12 * torch.rand(n, device=dev, generator=g), 50 + 50 * torch.rand(n, device=dev, generator=g)
    y = amp * fs * torch.exp(-xt / ts) + amp * (1 - fs) * torch.exp(-xt / tl) + 5 * torch.randn(E, n, device=dev, generator=g)
    y[:, ::211] = 0  # skipped voxels
    kw_fit, p0 = {}, (500.0, -1 / 10, 500.0, -1 / 60)
    if variant == "echo_fastest":
        y = y.t().contiguous()
        kw_fit["layout"] = "echo_fastest"
    elif variant == "int16":
        y = y.round().to(torch.int16)
    elif variant == "p0_voxel":
        p0 = (None, -1 / 10, None, -1 / 60)
        pv = torch.stack([amp * 0.5, torch.zeros_like(amp), amp * 0.5, torch.zeros_like(amp)], dim=1).contiguous()
        kw_fit["p0_voxel"] = pv
    o, P = A.make_opts(D.biexponential, p0=p0, compute_dtype="f32")

    def run():
        od = torch.float64 if variant == "f64_maps_status" else torch.float32
        popt = torch.full((n, P), -7.0, device=dev, dtype=od)
        r2 = torch.full((n,), -7.0, device=dev, dtype=od)
        st = torch.full((n,), 99, device=dev, dtype=torch.uint8)
        it = torch.full((n,), 99, device=dev, dtype=torch.uint8)
        A.fit_device(o, P, x, y, popt=popt, r2=r2, status=st, niter=it, **kw_fit)
        torch.cuda.synchronize()
        return popt, r2, st, it

    monkeypatch.setenv("DFIT_LMQ", "0")
    ref = run()
    monkeypatch.delenv("DFIT_LMQ")
    out = run()
    assert not (out[2] == 99).any() and not (out[0] == -7.0).any()
    assert torch.equal(out[2] == 0, ref[2] == 0) and int((out[2] == 0).sum()) == len(range(0, n, 211))
    ok = (ref[2] >= 1) & (ref[2] <= 4) & (out[2] >= 1) & (out[2] <= 4)
    assert ok.float().mean() > 0.95 and ((ref[2] >= 5) != (out[2] >= 5)).float().mean() < 5e-3
    assert (ref[1][ok] - out[1][ok]).abs().max() < 2e-5
    rel = ((ref[0][ok] - out[0][ok]).abs() / ref[0][ok].abs()).max(dim=1).values
    assert float(torch.quantile(rel.float(), 0.5)) < 2e-5 and float(torch.quantile(rel.float(), 0.99)) < 5e-3
    assert (ref[3][ok].float() - out[3][ok].float()).abs().mean() < 0.05


@pytest.mark.parametrize("n,offset", [(8 * 5000, 0), (8 * 5000 + 3, 0), (4099, 0), (40000, 3), (7, 0), (300000, 8)])
@pytest.mark.parametrize("form", ["f32", "f64_status", "out_param", "biexp"])
def test_mask_compaction_and_fill(D, n, offset, form):
    """The mask pass (mask_compact_kernel: 8 consecutive voxels per thread, 16-byte fill stores, shuffle-scan compaction):
    inside the mask the masked fit equals the dense fit (to rounding: the two may take different kernels), outside it every output holds the fill value --
    ragged sizes, a mask that does not start on an 8-byte boundary, float64 maps with status / pass-count bytes, one
    selected parameter with a nan_to_num fill, a four-parameter model."""
    import torch

    from dosma_b200 import device_api as A

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(n + offset)
    E = 8 if form != "biexp" else 10
    x = [10.0 * i for i in range(1, E + 1)]
    xt = torch.tensor(x, device=dev)[:, None]
    y = (500 + 1000 * torch.rand(n, device=dev, generator=g)) * torch.exp(-xt / (10 + 70 * torch.rand(n, device=dev, generator=g))) \
        + 10 * torch.randn(E, n, device=dev, generator=g)
    big = torch.rand(n + offset, device=dev, generator=g) < 0.2
    big[offset: offset + min(n, 64)] = True   # a solid run and, below, a solid hole
    if n > 200:
        big[offset + 100: offset + 164] = False
    mask = big[offset:]                        # (offset != 0: the mask bytes do not start on an 8-byte boundary)
    post, kw, P_model = None, {}, D.monoexponential
    if form == "out_param":
        post = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, 2], r2_threshold=0.5, nan_to_num=0.0)
        kw = dict(out_param=1)
    p0 = (1.0, -1 / 30)
    if form == "biexp":
        P_model, p0 = D.biexponential, (500.0, -1 / 10, 500.0, -1 / 60)
    o, P = A.make_opts(P_model, p0=p0, post=post, **kw)
    od = torch.float64 if form == "f64_status" else torch.float32
    shape = (n,) if form == "out_param" else (n, P)

    def run(m):
        popt = torch.full(shape, -7.0, device=dev, dtype=od)
        r2 = torch.full((n,), -7.0, device=dev, dtype=od)
        st = torch.full((n,), 99, device=dev, dtype=torch.uint8) if form == "f64_status" else None
        it = torch.full((n,), 99, device=dev, dtype=torch.uint8) if form == "f64_status" else None
        A.fit_device(o, P, x, y, mask=m, popt=popt, r2=r2, status=st, niter=it)
        torch.cuda.synchronize()
        return popt, r2, st, it

    dense, masked = run(None), run(mask)
    fill = 0.0 if form == "out_param" else float("nan")
    for d, m in zip(dense[:2], masked[:2]):
        inside, ref = m[mask].double(), d[mask].double()  # (dense and masked launches may take different kernels: to rounding)
        assert torch.equal(torch.isnan(inside), torch.isnan(ref))
        tol = 1e-5 if form != "out_param" else 1.001e-2  # (out_param: rounded to 2 decimals -- one rounding step)
        err = (inside - ref).abs() / (ref.abs() if form != "out_param" else 1.0)
        assert float(err.nan_to_num(0.0).max()) <= tol
        outside = m[~mask]
        assert bool(torch.isnan(outside).all()) if fill != fill else bool((outside == fill).all())
    if form == "f64_status":
        assert bool((masked[2][~mask] == 0).all()) and bool((masked[3][~mask] == 0).all())
        assert (masked[2][mask] != dense[2][mask]).float().mean() < 1e-3 and (masked[3][mask] != dense[3][mask]).float().mean() < 1e-2


def test_config4_sample_against_c_oracle(D):
    """A config-4-shaped volume (16 echoes x 5 ms, bi-exponential, SNR 100, fp32) fitted on the GPU in fp32; a seeded
    sample of voxels is compared with the MINPACK restatement (oracle/minpack_lmdif.c, pinned to SciPy on the
    bi-exponential fixtures): same minimum (r2), parameters within the ftol slack of the reference."""
    import torch

    from dosma_b200 import device_api as A
    from oracle import c_oracle

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    x16 = [5.0 * i for i in range(1, 17)]
    n = 256 * 256 * 128
    xt = torch.tensor(x16, device=dev, dtype=torch.float32)[:, None]
    amp = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
    ts = 8 + 12 * torch.rand(n, device=dev, generator=g)
    tl = 50 + 50 * torch.rand(n, device=dev, generator=g)
    y = amp * fs * torch.exp(-xt / ts) + amp * (1 - fs) * torch.exp(-xt / tl) + 10 * torch.randn(16, n, device=dev, generator=g)
    p0 = (500.0, -1 / 10, 500.0, -1 / 60)
    o, P = A.make_opts(D.biexponential, p0=p0, compute_dtype="f32")
    p, r = A.fit_device(o, P, x16, y)
    torch.cuda.synchronize()
    sel = torch.from_numpy(np.random.default_rng(4).choice(n, 4096, replace=False)).to(dev)
    ys = y[:, sel].double().cpu().numpy()
    ps, rs = p[sel].double().cpu().numpy(), r[sel].double().cpu().numpy()
    pr, rr = c_oracle.curve_fit("biexponential", x16, ys, p0=p0)
    nan, ref_nan = np.isnan(ps[:, 0]), np.isnan(pr[:, 0])
    assert (nan ^ ref_nan).mean() < 0.03, ((nan ^ ref_nan).sum(), nan.sum(), ref_nan.sum())
    ok = ~nan & ~ref_nan
    assert np.abs(rs[ok] - rr[ok]).max() < 2e-6
    rel = (np.abs(ps[ok] - pr[ok]) / np.abs(pr[ok])).max(axis=1)
    q = np.percentile(rel, [50, 90])
    assert q[0] < 3e-4 and q[1] < 5e-3, q
    assert float(torch.isnan(p[:, 0]).float().mean()) < 0.05


@pytest.mark.parametrize("name", G.names("monoexpfit_"))
def test_monoexpfit_golden(D, name):
    c = G.load(name)
    m = c["meta"]
    shape = tuple(m["shape"])
    vols = [D.MedicalVolume(c["y"][e].reshape(shape), c["affine"]) for e in range(c["y"].shape[0])]
    mask = D.MedicalVolume(c["mask"], c["affine"]) if m["use_mask"] else None
    tc, r2 = D.MonoExponentialFit(bounds=tuple(m["bounds"]), tc0=m["tc0"], decimal_precision=m["decimal_precision"]).fit(
        c["x"], vols, mask=mask)
    assert tc.volume.shape == shape and tc.volume.dtype == np.float64
    assert np.array_equal(tc.affine, c["affine"])
    if m["use_mask"]:
        assert (tc.volume[~c["mask"]] == 0).all() and (r2.volume[~c["mask"]] == 0).all()
    # rounded maps: identical except where the un-rounded values straddle a rounding boundary or the r2
    # threshold (both solvers agree to ~1e-5 relative there, the rounding step is 1e-3)
    diff = np.abs(tc.volume - c["tc"])
    step = 10.0 ** (-m["decimal_precision"])
    zeroed = (tc.volume == 0) != (c["tc"] == 0)
    lim = 0.03 if "snr30" in name else 0.01
    assert zeroed.mean() < lim, zeroed.mean()
    close = diff[~zeroed] <= 1.001 * step
    # large-residual voxels (r2 0.90-0.97: snr30, flipped-sign samples) are where MINPACK's ftol=1e-5 stop
    # leaves ~3e-4 relative slack, i.e. more than one rounding step
    need = 0.95 if "snr30" in name else (0.99 if "zeros_negatives" in name else 0.995)
    assert close.mean() > need, close.mean()
    # r2 is compared on accepted fits; below the 0.9 threshold both maps read 0 and the r2 of such a
    # rejected (degenerate / noise-only) fit depends on where each solver happened to stop
    kept = (tc.volume != 0) & (c["tc"] != 0)
    assert np.quantile(np.abs(r2.volume - c["r2"])[kept], 0.99) < 1e-4


def test_curvefitter_golden(D):
    c = G.load("curvefitter_mask_nan")
    shape = tuple(c["meta"]["shape"])
    vols = [D.MedicalVolume(c["y"][e].reshape(shape), c["affine"]) for e in range(8)]
    popt, r2 = D.CurveFitter(D.monoexponential, p0=(1.0, -1 / 30)).fit(c["x"], vols, mask=c["mask"])
    assert popt.volume.shape == shape + (2,)
    assert np.isnan(popt.volume[~c["mask"]]).all() and np.isnan(r2.volume[~c["mask"]]).all()
    same = np.isnan(popt.volume) == np.isnan(c["popt"])
    assert same.mean() > 0.99
    ok = ~np.isnan(popt.volume) & ~np.isnan(c["popt"])
    assert (_rel(popt.volume[ok], c["popt"][ok]) > 1e-4).mean() < 0.05

    c = G.load("curvefitter_post")
    m = c["meta"]
    vols = [D.MedicalVolume(c["y"][e].reshape(shape), c["affine"]) for e in range(8)]
    popt, r2 = D.CurveFitter(D.monoexponential, p0=(1.0, -1 / 30), out_ufuncs=[None, lambda v: 1 / np.abs(v)],
                             out_bounds=m["out_bounds"], r2_threshold=m["r2_threshold"],
                             nan_to_num=m["nan_to_num"]).fit(c["x"], vols)
    filled = (popt.volume == m["nan_to_num"]) == (c["popt"] == m["nan_to_num"])
    assert filled.mean() > 0.98
    ok = (popt.volume != m["nan_to_num"]) & (c["popt"] != m["nan_to_num"])
    assert (_rel(popt.volume[ok], c["popt"][ok]) > 1e-4).mean() < 0.05


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint16, np.int32, np.uint8])
def test_input_dtypes(D, dtype):
    from oracle import c_oracle

    rng = np.random.default_rng(11)
    x = np.arange(1, 9) * 10.0
    n = 3000
    amp = 200 if dtype == np.uint8 else 1200
    y = np.round(rng.uniform(0.5 * amp, amp, n) * np.exp(-x[:, None] / rng.uniform(20, 80, n))).astype(dtype)
    popt, r2 = D.curve_fit(D.monoexponential, x, y, p0=(1.0, -1 / 30), compute_dtype="f64")
    ref, ref_r2 = c_oracle.curve_fit("monoexponential", x, y.astype(np.float64), p0=(1.0, -1 / 30))
    ok = ~np.isnan(ref[:, 0]) & ~np.isnan(popt[:, 0])
    assert ok.mean() > 0.999
    rel = _rel(popt[ok], ref[ok]).max(axis=1)
    assert (rel > 1e-4).mean() < 0.02 and np.median(rel) < 1e-5


def test_device_layouts_and_determinism(D):
    """Planar (E, N) and echo-fastest (N, E) device layouts give bit-identical results; repeated runs
    are bit-identical (the caller-level reference tests require `is_identical` maps)."""
    import torch

    from dosma_b200 import device_api as A

    rng = np.random.default_rng(5)
    x = np.arange(1, 9) * 10.0
    n = 100_003  # not a multiple of anything
    y = (rng.uniform(500, 1500, n) * np.exp(-x[:, None] / rng.uniform(10, 80, n)) + rng.normal(0, 10, (8, n))).astype(
        np.float32)
    yp = torch.from_numpy(y).cuda()
    ye = yp.t().contiguous()
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
    p1, r1 = A.fit_device(o, P, x, yp, layout="planar")
    p2, r2 = A.fit_device(o, P, x, ye, layout="echo_fastest")
    p3, r3 = A.fit_device(o, P, x, yp, layout="planar")
    torch.cuda.synchronize()
    assert torch.equal(p1, p3) and torch.equal(r1, r3)
    assert torch.equal(p1.nan_to_num(-1), p2.nan_to_num(-1)) and torch.equal(r1, r2)


def test_tma_staged_variant_is_bitwise_identical(D, monkeypatch):
    """The TMA-staged persistent kernel (cp.async.bulk.tensor + mbarrier ring, use_tma=1), the plain
    coalesced-load kernel and the LM-in-rounds kernel run the same per-voxel solver: the same voxels fitted / failed
    and the same minimiser to fp32 resolution, including a ragged last tile.  (Bit-identical they were only as long as
    ptxas happened to contract the same multiply-adds in each kernel.)"""
    import torch

    from dosma_b200 import device_api as A

    g = torch.Generator(device="cuda").manual_seed(9)
    x = np.arange(1, 9) * 10.0
    xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
    for n in (40, 100_004, 1_000_000):
        y = (500 + 1000 * torch.rand(n, device="cuda", generator=g)) * torch.exp(
            -xt / (10 + 70 * torch.rand(n, device="cuda", generator=g))) + 10 * torch.randn(8, n, device="cuda", generator=g)
        mask = torch.rand(n, device="cuda", generator=g) > 0.3
        for init in ("given", "loglinear"):
            # the LM (fast_path=0) is the same code in both kernels
            o0, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), init=init, use_tma=0, fast_path=0)
            o1, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), init=init, use_tma=1, fast_path=0)
            monkeypatch.setenv("DFIT_LMQ", "0")  # (the plain one-voxel-per-lane kernel; the default is the LM in rounds)
            p0_, r0_ = A.fit_device(o0, P, x, y, mask=mask)
            p1_, r1_ = A.fit_device(o1, P, x, y, mask=mask)
            torch.cuda.synchronize()
            # (the same solver source in two kernels: ptxas is free to contract different multiply-adds in each, so the
            # last bits may differ -- the minimiser may not)
            assert torch.equal(torch.isnan(p0_), torch.isnan(p1_))
            assert ((p0_ - p1_).abs() / p1_.abs()).nan_to_num(0).max() < 1e-5
            assert (r0_ - r1_).abs().nan_to_num(0).max() < 1e-5
            monkeypatch.delenv("DFIT_LMQ")
            # the LM-in-rounds kernel (what use_tma=0 runs by default): the same iteration, exponentials by recurrence
            p2_, r2_ = A.fit_device(o0, P, x, y, mask=mask)
            torch.cuda.synchronize()
            assert torch.equal(torch.isnan(p2_), torch.isnan(p1_))
            assert ((p2_ - p1_).abs() / p1_.abs()).nan_to_num(0).max() < 1e-5
            assert (r2_ - r1_).abs().nan_to_num(0).max() < 1e-5
            # with the fast path the masked voxels go through the two-voxel list kernel (use_tma=0) or the
            # one-voxel TMA kernel (use_tma=1): same iteration, different packing -> equal to rounding
            o0, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), init=init, use_tma=0)
            o1, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), init=init, use_tma=1)
            p0_, r0_ = A.fit_device(o0, P, x, y, mask=mask)
            p1_, r1_ = A.fit_device(o1, P, x, y, mask=mask)
            torch.cuda.synchronize()
            assert torch.equal(torch.isnan(p0_), torch.isnan(p1_))
            assert ((p0_ - p1_).abs() / p1_.abs()).nan_to_num(0).max() < 1e-5
            assert (r0_ - r1_).abs().nan_to_num(0).max() < 1e-6
    # ineligible inputs must be refused loudly, not silently rerouted
    o1, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), use_tma=1)
    with pytest.raises(Exception):
        A.fit_device(o1, P, x, y.t().contiguous(), layout="echo_fastest")


def test_two_voxel_fast_kernels_agree_with_lm(D):
    """The dense mono-exponential fast path (variable-projection Newton, two voxels per lane; plain loads
    and the TMA-staged persistent variant) against the LM from p0 (fast_path=0): same minimiser, ragged
    and odd sizes, int16 samples, status / pass-count outputs, fused epilogue, degenerate voxels."""
    import torch

    from dosma_b200 import _cabi, device_api as A

    g = torch.Generator(device="cuda").manual_seed(11)
    x = np.arange(1, 9) * 10.0
    xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
    for n in (1, 2, 63, 64, 65, 100_003, 1_000_000):
        y = (500 + 1000 * torch.rand(n, device="cuda", generator=g)) * torch.exp(
            -xt / (10 + 70 * torch.rand(n, device="cuda", generator=g))) + 10 * torch.randn(8, n, device="cuda", generator=g)
        if n > 100:
            y[:, 5] = 0.0          # all-zero voxel: skipped (fitting.py:1065-1067)
            y[:, 8] = -y[:, 8]     # negative amplitude
            y[:, 11] = 7.0         # constant signal
            y[:, 14] = 10 * torch.randn(8, device="cuda", generator=g)  # pure noise
        o_lm, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), fast_path=0, use_tma=0)
        p_lm, r_lm = A.fit_device(o_lm, P, x, y)
        ref = None
        for kw in (dict(fast_path=2, use_tma=0), dict(fast_path=1, use_tma=0), dict(fast_path=1, use_tma=1), dict()):
            if kw.get("use_tma") == 1 and n % 4:
                continue  # an explicit use_tma=1 refuses rows whose pitch is not a multiple of 16 bytes
            o, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), **kw)
            st = torch.zeros(n, dtype=torch.uint8, device="cuda")
            it = torch.zeros(n, dtype=torch.uint8, device="cuda")
            p, r = A.fit_device(o, P, x, y, status=st, niter=it)
            torch.cuda.synchronize()
            stats = _cabi.get_handle(0).stats()
            ok = ~torch.isnan(p_lm[:, 0]) & ~torch.isnan(p[:, 0])
            if n > 100:
                # the constant voxel has b = 0 (no relative error) and pure noise has several local minima
                # (the LM finds the one in p0's basin): checked on their own terms below
                ok[11] = ok[14] = False
                assert abs(float(p[11, 1])) < 1e-6 and abs(float(p[11, 0]) - 7.0) < 1e-4 and int(st[11]) in (1, 2, 3, 4)
                assert int(st[14]) in (1, 2, 3, 4, 5) and int(st[8]) in (1, 2, 3, 4) and float(p[8, 0]) < 0
            assert ok.float().mean() > 0.999
            rel = ((p[ok] - p_lm[ok]).abs() / p_lm[ok].abs())
            assert rel.max() < 2e-3 and (rel > 1e-4).float().mean() < 1e-3, (n, kw, float(rel.max()))
            assert (r[ok] - r_lm[ok]).abs().max() < 1e-5
            assert int(((st >= 1) & (st <= 4)).sum()) == stats["n_fitted"] and stats["n_voxels"] == n
            assert int(it.sum()) == stats["sum_iters"]
            if n > 100:
                assert int(st[5]) == 0 and torch.isnan(p[5]).all() and float(r[5]) == 0.0
            if ref is None:
                ref = (p, r)
            else:  # all fast variants run the same arithmetic per voxel up to packing order
                okr = ~torch.isnan(ref[0][:, 0]) & ~torch.isnan(p[:, 0])
                assert ((p[okr] - ref[0][okr]).abs() / ref[0][okr].abs().clamp_min(1e-30)).max() < 1e-5
    # int16 samples and the fused MonoExponentialFit epilogue (ufunc, bounds, r2 threshold, fill, rounding)
    for n in (200_001, 200_064):  # odd: plain 4-byte pair loads; multiple of 8: 16-bit tiles through TMA
        yi = ((500 + 1000 * torch.rand(n, device="cuda", generator=g)) * torch.exp(
            -xt / (10 + 70 * torch.rand(n, device="cuda", generator=g))) + 10 * torch.randn(8, n, device="cuda", generator=g)
              ).round().to(torch.int16)
        post = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, 3], r2_threshold=0.9, nan_to_num=0.0)
        outs = []
        for kw in (dict(fast_path=0, use_tma=0), dict(fast_path=1, use_tma=0), dict()):
            o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post, **kw)
            p, r = A.fit_device(o, P, x, yi, out_dtype=torch.float64)
            torch.cuda.synchronize()
            outs.append((p, r))
        for p, r in outs[1:]:
            same = (p[:, 1] - outs[0][0][:, 1]).abs() <= 1.001e-3  # one rounding step
            assert same.float().mean() > 0.999
            assert (r - outs[0][1]).abs().max() < 1e-5
    # non-uniform echo times: the general (exp-based) fast path, dense and through a mask, against the LM
    for xn in (np.array([10.0, 20.0, 40.0, 80.0]), np.array([0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0])):
        for n in (300_001, 300_032):  # odd: one-voxel kernel (no pair loads); multiple of 4: TMA two-voxel kernel
            xg = torch.tensor(xn, device="cuda", dtype=torch.float32)[:, None]
            yn = ((500 + 1000 * torch.rand(n, device="cuda", generator=g)) * torch.exp(
                -xg / (10 + 70 * torch.rand(n, device="cuda", generator=g))) + 10 * torch.randn(len(xn), n, device="cuda", generator=g))
            mask = torch.rand(n, device="cuda", generator=g) > 0.7
            o0, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), fast_path=0)
            o1, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
            pa_, ra_ = A.fit_device(o0, P, xn, yn)
            pb_, rb_ = A.fit_device(o1, P, xn, yn)
            pm_, rm_ = A.fit_device(o1, P, xn, yn, mask=mask)
            torch.cuda.synchronize()
            assert _cabi.get_handle(0).stats()["n_fitted"] == int(mask.sum())
            ok = ~torch.isnan(pa_[:, 0]) & ~torch.isnan(pb_[:, 0])
            assert ok.float().mean() > 0.999
            rel = (pb_[ok] - pa_[ok]).abs() / pa_[ok].abs()
            assert rel.max() < 2e-3 and (rel > 1e-4).float().mean() < 1e-3 and (rb_[ok] - ra_[ok]).abs().max() < 1e-5
            # the mask path (two-voxel list kernel) runs the same per-voxel arithmetic as the dense two-voxel kernel:
            # identical inside, NaN outside; the dense one-voxel kernel (odd n) runs the loop form of the same
            # iteration: equal to rounding
            if n % 4 == 0:
                assert torch.equal(pm_[mask].nan_to_num(-1), pb_[mask].nan_to_num(-1)) and torch.equal(rm_[mask], rb_[mask])
            else:
                okm = ~torch.isnan(pm_[mask][:, 0]) & ~torch.isnan(pb_[mask][:, 0])
                assert okm.float().mean() > 0.999
                assert ((pm_[mask][okm] - pb_[mask][okm]).abs() / pb_[mask][okm].abs()).max() < 2e-5
                assert (rm_[mask][okm] - rb_[mask][okm]).abs().max() < 2e-6
            assert torch.isnan(pm_[~mask]).all()
    # y_bounds keep the voxel-skipping rules with the LM: identical results with and without the fast path
    xn = np.arange(1, 9) * 10.0
    o0, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), fast_path=0, y_bounds=(0, 1400))
    o1, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), y_bounds=(0, 1400))
    pa_, ra_ = A.fit_device(o0, P, xn, y)
    pb_, rb_ = A.fit_device(o1, P, xn, y)
    torch.cuda.synchronize()
    assert torch.equal(pa_.nan_to_num(-1), pb_.nan_to_num(-1)) and torch.equal(ra_, rb_)


def test_scaling_and_permutation_properties(D):
    """Size-independent properties at a BASELINE-sized workload (384 x 384 x 16 slab, 8 echoes):
    scaling y by 2 scales a by 2 and leaves b (to rounding); permuting voxels permutes results exactly."""
    import torch

    from dosma_b200 import device_api as A

    g = torch.Generator(device="cuda").manual_seed(3)
    n = 384 * 384 * 16
    x = np.arange(1, 9) * 10.0
    xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
    a = 500 + 1000 * torch.rand(n, device="cuda", generator=g)
    t2 = 10 + 70 * torch.rand(n, device="cuda", generator=g)
    y = a * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device="cuda", generator=g)
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
    p, r = A.fit_device(o, P, x, y)
    p2, r2 = A.fit_device(o, P, x, 2 * y)
    perm = torch.randperm(n, device="cuda", generator=g)
    pp, rp = A.fit_device(o, P, x, y[:, perm].contiguous())
    torch.cuda.synchronize()
    ok = ~torch.isnan(p[:, 0]) & ~torch.isnan(p2[:, 0])
    assert ok.float().mean() > 0.999
    assert torch.equal(pp.nan_to_num(-1), p[perm].nan_to_num(-1)) and torch.equal(rp, r[perm])
    assert ((p2[ok, 0] / p[ok, 0] - 2).abs() < 1e-4).float().mean() > 0.999
    assert (((p2[ok, 1] - p[ok, 1]) / p[ok, 1]).abs() < 1e-4).float().mean() > 0.999
    # round trip: recovered parameters near the truth (SNR 100 -> a few % statistical error at most)
    assert ((p[ok, 1] + 1 / t2[ok]).abs() * t2[ok]).median() < 0.02


def test_full_size_round_trip(D):
    """BASELINE config 2 (384 x 384 x 160, 8 echoes, 23.6 M voxels), noise-free: encode -> fit -> decode."""
    import torch

    from dosma_b200 import device_api as A

    g = torch.Generator(device="cuda").manual_seed(1)
    n = 384 * 384 * 160
    x = np.arange(1, 9) * 10.0
    xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
    a = 500 + 1000 * torch.rand(n, device="cuda", generator=g)
    t2 = 10 + 70 * torch.rand(n, device="cuda", generator=g)
    y = a * torch.exp(-xt / t2)
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
    p, r = A.fit_device(o, P, x, y)
    torch.cuda.synchronize()
    assert not torch.isnan(p).any()
    assert ((p[:, 0] - a).abs() / a).max() < 1e-4
    assert ((p[:, 1] + 1 / t2).abs() * t2).max() < 1e-4
    assert (r > 1 - 1e-5).all()
    st = A._cabi.get_handle(0).stats()
    assert st["n_fitted"] == n and st["n_failed"] == 0


def test_config3_masked_t1rho_full_size(D):
    """BASELINE config 3 (512 x 512 x 256, 7 MAPSS-like spin-lock times, ~10 % tissue mask): fitting with the
    mask equals fitting everything and selecting (bitwise), outside the mask reads NaN / fill, and the
    statistics count exactly the masked voxels."""
    import torch

    from dosma_b200 import _cabi, device_api as A

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(2)
    x7 = [0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]  # tests/scan_sequences/mri/test_mapss.py:43
    shape = (512, 512, 256)
    n = int(np.prod(shape))
    xt = torch.tensor(x7, device=dev, dtype=torch.float32)[:, None]
    y = (500 + 1000 * torch.rand(n, device=dev, generator=g)) * torch.exp(-xt / (20 + 100 * torch.rand(n, device=dev, generator=g)))
    y += 10 * torch.randn(7, n, device=dev, generator=g)
    zz, yy, xx = torch.meshgrid(*[torch.linspace(-1, 1, s, device=dev) for s in shape], indexing="ij")
    rad = (zz ** 2 + yy ** 2 + (xx * 1.6) ** 2).sqrt()
    mask = ((rad > 0.55) & (rad < 0.62)).reshape(-1)
    del zz, yy, xx, rad
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
    p_all, r_all = A.fit_device(o, P, x7, y)
    p_m, r_m = A.fit_device(o, P, x7, y, mask=mask)
    torch.cuda.synchronize()
    st = _cabi.get_handle(0).stats()
    assert st["n_fitted"] == int(mask.sum())
    assert torch.equal(p_m[mask].nan_to_num(-1), p_all[mask].nan_to_num(-1)) and torch.equal(r_m[mask], r_all[mask])
    assert torch.isnan(p_m[~mask]).all() and torch.isnan(r_m[~mask]).all()
    post = {"ufunc": [0, 1], "lb": [-np.inf, 0.0], "ub": [np.inf, 500.0], "decimals": [-1, 3], "r2_threshold": 0.9,
            "nan_to_num": 0.0}
    o2, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post)
    p_f, r_f = A.fit_device(o2, P, x7, y, mask=mask)
    torch.cuda.synchronize()
    assert (p_f[~mask] == 0).all() and (r_f[~mask] == 0).all()  # MonoExponentialFit convention (fitting.py:731)
    keep = mask & (r_m >= 0.9) & (p_m[:, 1] != 0)
    tc = (1 / p_m[keep, 1].abs().double())
    inb = tc <= 500
    assert torch.allclose(p_f[keep, 1][inb].double(), torch.round(tc[inb] * 1e3) / 1e3, atol=2e-3)


def test_config4_biexp_full_size_round_trip(D):
    """BASELINE config 4 (256 x 256 x 128, 16 echoes, bi-exponential), noise-free: fp64 arithmetic recovers the
    generating parameters; fp32 arithmetic converges on the same voxels with the accuracy fp32 allows for this
    ill-conditioned model."""
    import torch

    from dosma_b200 import device_api as A

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    x16 = [5.0 * i for i in range(1, 17)]
    n = 256 * 256 * 128
    xt = torch.tensor(x16, device=dev, dtype=torch.float64)[:, None]
    amp = 500 + 1000 * torch.rand(n, device=dev, generator=g, dtype=torch.float64)
    fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g, dtype=torch.float64)
    ts = 8 + 12 * torch.rand(n, device=dev, generator=g, dtype=torch.float64)
    tl = 50 + 50 * torch.rand(n, device=dev, generator=g, dtype=torch.float64)
    y = amp * fs * torch.exp(-xt / ts) + amp * (1 - fs) * torch.exp(-xt / tl)
    o, P = A.make_opts(D.biexponential, p0=(500.0, -1 / 10, 500.0, -1 / 60), compute_dtype="f64")
    p, r = A.fit_device(o, P, x16, y, out_dtype=torch.float64)
    torch.cuda.synchronize()
    ok = ~torch.isnan(p[:, 0])
    assert ok.double().mean() > 0.99
    truth = torch.stack([amp * fs, -1 / ts, amp * (1 - fs), -1 / tl], dim=1)
    rel = ((p[ok] - truth[ok]).abs() / truth[ok].abs()).max(dim=1).values
    assert (rel < 1e-6).double().mean() > 0.99, float((rel < 1e-6).double().mean())
    assert (r[ok] > 1 - 1e-9).all()
    o32, _ = A.make_opts(D.biexponential, p0=(500.0, -1 / 10, 500.0, -1 / 60), compute_dtype="f32")
    p32, r32 = A.fit_device(o32, P, x16, y.float())
    torch.cuda.synchronize()
    ok32 = ~torch.isnan(p32[:, 0])
    assert ok32.float().mean() > 0.99
    rel32 = ((p32[ok32].double() - truth[ok32]).abs() / truth[ok32].abs()).max(dim=1).values
    assert rel32.median() < 1e-3 and (r32[ok32] > 1 - 1e-5).all()


def test_fp32_fused_epilogue_equals_float64_epilogue(D):
    """The fused MonoExponentialFit epilogue (1/|b|, bounds, r2 threshold, fill, rounding) computed in fp32 by the
    kernels (float32 maps) must be, bit for bit, the float64 epilogue's result converted to float32 -- on a million
    voxels whose time constants straddle the bounds and the rounding ties, for every kernel variant."""
    import torch

    from dosma_b200 import device_api as A

    g = torch.Generator(device="cuda").manual_seed(21)
    x = np.arange(1, 9) * 10.0
    xt = torch.tensor(x, device="cuda", dtype=torch.float32)[:, None]
    n = 1_000_000
    t2 = 5 + 120 * torch.rand(n, device="cuda", generator=g)  # beyond the upper bound of 100 as well
    y = (500 + 1000 * torch.rand(n, device="cuda", generator=g)) * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device="cuda", generator=g)
    for decimals in (1, 3, -1):
        post = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, decimals], r2_threshold=0.9, nan_to_num=0.0)
        for kw in (dict(), dict(use_tma=0), dict(fast_path=2, use_tma=0), dict(fast_path=0, use_tma=0)):
            o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post, **kw)
            p32, r32 = A.fit_device(o, P, x, y, out_dtype=torch.float32)
            p64, r64 = A.fit_device(o, P, x, y, out_dtype=torch.float64)
            torch.cuda.synchronize()
            assert torch.equal(p32, p64.float()) and torch.equal(r32, r64.float()), (decimals, kw)
            if decimals >= 0:  # float64 maps are on numpy's decimal grid (np.around is idempotent on them); fill = 0
                tc = p64[:, 1].cpu().numpy()
                assert np.array_equal(np.around(tc, decimals), tc)
            assert float((p32[:, 1] == 0).float().mean()) > 0.05 and float((p32[:, 1] > 0).float().mean()) > 0.5


class _Hdr(dict):
    """Header stand-in (pydicom is not installed): `.get` access, deep-copyable."""


def _hdrs(shape, **fields):
    arr = np.empty(int(np.prod(shape)), dtype=object)
    for i in range(arr.size):
        arr[i] = _Hdr(fields)
    return arr.reshape(shape)


def test_fit_marshalling_headers_4d_copy_headers(D):
    """`_Fitter.fit` marshalling on the real engine (fitting.py:157-235; the reference's TestCurveFitter.test_headers
    :433-477 and TestMonoExponentialFit.test_headers :220-236): 4-D inputs give 5-D parameter maps, headers of y[0]
    are deep-copied and get a trailing axis, `copy_headers=False` drops them, a one-parameter model works, the
    time-constant map of MonoExponentialFit keeps (1, 1, Z) headers, reoriented inputs come back in y[0]'s frame."""
    rng = np.random.default_rng(8)
    x = np.asarray([0.5, 1.0, 2.0, 4.0])
    b = rng.random((10, 10, 20, 4)) + 0.1
    y = [D.MedicalVolume(D.monoexponential(t, 1.0, b), np.eye(4), headers=_hdrs((1, 1, 20, 4), EchoNumbers=i))
         for i, t in enumerate(x)]
    popt, r2 = D.CurveFitter(D.monoexponential).fit(x, y)
    assert popt.shape == (10, 10, 20, 4, 2) and r2.shape == (10, 10, 20, 4) and popt.volume.dtype == np.float64
    assert np.allclose(popt.volume[..., 0], 1.0) and np.allclose(popt.volume[..., 1], b)
    assert popt.headers().shape == (1, 1, 20, 4, 1) and r2.headers().shape == (1, 1, 20, 4)
    assert all(h.get("EchoNumbers") == 0 for h in popt.headers().flatten())
    assert popt.headers().flatten()[0] is not y[0].headers().flatten()[0]  # deep copy (fitting.py:225)
    a_hat, b_hat = popt[..., 0], popt[..., 1]  # the reference test's own indexing
    assert np.allclose(b_hat.volume, b) and b_hat.headers().shape == (1, 1, 20, 4) and a_hat.shape == b.shape
    popt, r2 = D.CurveFitter(D.monoexponential).fit(x, y, copy_headers=False)
    assert np.allclose(popt.volume[..., 1], b) and popt.headers() is None and r2.headers() is None
    a = rng.random((10, 10, 20, 4)) + 0.1
    yl = [D.MedicalVolume(a * t, np.eye(4), headers=_hdrs((1, 1, 20, 4), EchoNumbers=i)) for i, t in enumerate(x)]
    popt, _ = D.CurveFitter(D.linear).fit(x, yl)
    assert popt.shape == (10, 10, 20, 4, 1) and np.allclose(popt.volume[..., 0], a)
    # MonoExponentialFit (3-D, headers): TestMonoExponentialFit.test_headers
    b3 = rng.random((10, 10, 20)) + 0.1
    y3 = [D.MedicalVolume(D.monoexponential(t, 1.0, b3), np.eye(4),
                          headers=_hdrs((1, 1, 20), StudyDescription="Sample study", EchoNumbers=i)) for i, t in enumerate(x)]
    t_hat = D.MonoExponentialFit(decimal_precision=8).fit(x, y3)[0]
    assert np.allclose(t_hat.volume, 1 / np.abs(b3)) and t_hat.headers().shape == (1, 1, 20)
    assert all(h.get("StudyDescription") == "Sample study" for h in t_hat.headers().flatten())
    # echoes stored in another orientation are brought to y[0]'s frame (fitting.py:189-190)
    y3r = [y3[0]] + [v.reformat(("SI", "AP", "LR")) for v in y3[1:]]
    assert y3r[1].orientation != y3r[0].orientation
    t_hat2 = D.MonoExponentialFit(decimal_precision=8).fit(x, y3r)[0]
    assert t_hat2.orientation == y3[0].orientation and np.allclose(t_hat2.volume, t_hat.volume)


def test_patch_dosma_swaps_the_engine_in(D):
    """`dosma_b200.patch_dosma()` rebinds the three names in every DOSMA module that holds them
    (fitting.py:24-32 and the scan pipelines that bind them at import, cube_quant.py:9 / mapss.py:10 / cones.py:10):
    checked on stand-in modules (DOSMA itself is not importable on the GPU box), and a pipeline-style call through the
    patched module runs on the GPU."""
    import sys
    import types

    names = ["dosma", "dosma.core", "dosma.core.fitting", "dosma.scan_sequences", "dosma.scan_sequences.mri",
             "dosma.scan_sequences.mri.cube_quant", "dosma.scan_sequences.mri.mapss", "dosma.scan_sequences.mri.cones"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        for n in names:
            mod = types.ModuleType(n)
            mod.__path__ = []
            sys.modules[n] = mod
        sentinel = object()
        for n in ("dosma", "dosma.core", "dosma.core.fitting"):
            sys.modules[n].CurveFitter = sys.modules[n].MonoExponentialFit = sys.modules[n].curve_fit = sentinel
        for n in names[5:]:
            sys.modules[n].MonoExponentialFit = sentinel  # `from dosma.core.fitting import MonoExponentialFit`
        patched = D.patch_dosma()
        assert len(patched) == 12, patched
        F = sys.modules["dosma.core.fitting"]
        assert F.CurveFitter is D.CurveFitter and F.curve_fit is D.curve_fit
        for n in names[5:]:
            assert sys.modules[n].MonoExponentialFit is D.MonoExponentialFit and not hasattr(sys.modules[n], "CurveFitter")
        # what cube_quant.generate_t1_rho_map does (cube_quant.py:170-178), through the patched module global
        x = [10.0, 20.0, 40.0, 80.0]
        t = np.random.default_rng(3).uniform(10, 80, (16, 16, 4))
        vols = [D.MedicalVolume((1000 * np.exp(-xi / t)).astype(np.float32), np.eye(4)) for xi in x]
        fitter = sys.modules["dosma.scan_sequences.mri.cube_quant"].MonoExponentialFit(
            bounds=(0, 100), tc0="polyfit", decimal_precision=1, num_workers=4)
        tc, r2 = fitter.fit(x, vols)
        assert np.abs(tc.volume - np.around(t, 1)).max() < 0.11 and fitter.last_stats["n_launches"] >= 1
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def test_host_widening_equals_device_float64_maps(D, monkeypatch):
    """fp32 arithmetic with float64 result maps: by default the maps cross PCIe as float32 and the host threads widen
    them (rounded parameters back onto numpy's decimal grid); DFIT_HOST_WIDEN=0 has the device write float64.  Raw
    parameters, r2 and rounded time constants must be bit-identical either way; an unrounded 1/|b| carries the float32
    rounding of the epilogue's result (1e-7 relative)."""
    rng = np.random.default_rng(12)
    x = np.arange(1, 9) * 10.0
    n = 300_001
    t2 = rng.uniform(5, 130, n)
    y = (rng.uniform(500, 1500, n) * np.exp(-x[:, None] / t2) + rng.normal(0, 10, (8, n))).astype(np.float32)
    shape = (n, 1, 1)
    vols = [D.MedicalVolume(y[e].reshape(shape), np.eye(4)) for e in range(8)]
    mask = rng.random(shape) > 0.2

    def run():
        out = [D.curve_fit(D.monoexponential, x, y, p0=(1.0, -1 / 30))]
        out.append(tuple(v.volume for v in D.MonoExponentialFit(tc0="polyfit", decimal_precision=1).fit(x, vols)))
        out.append(tuple(v.volume for v in D.MonoExponentialFit(decimal_precision=3, bounds=(0, 60)).fit(x, vols, mask=mask)))
        f = D.CurveFitter(D.monoexponential, p0=(1.0, -1 / 30), out_ufuncs=(None, lambda v: 1 / np.abs(v)), out_bounds=((0, 2000), (0, 100)),
                          r2_threshold=0.9, nan_to_num=None)
        out.append(tuple(v.volume for v in f.fit(x, vols, mask=mask)))
        return out

    wide = run()
    monkeypatch.setenv("DFIT_HOST_WIDEN", "0")
    dev64 = run()
    for k, (a, b) in enumerate(zip(wide, dev64)):
        for u, v in zip(a, b):
            assert u.dtype == np.float64 and v.dtype == np.float64 and u.shape == v.shape
            if k == 3:  # unrounded ufunc output
                assert np.array_equal(np.isnan(u), np.isnan(v)) and np.allclose(u, v, rtol=1.2e-7, atol=0, equal_nan=True)
            else:
                assert np.array_equal(u, v, equal_nan=True), k
    tc = wide[1][0]
    assert np.array_equal(np.around(tc, 1), tc) and (tc > 0).mean() > 0.5  # on numpy's float64 decimal grid


def test_nonfinite_input_raises(D):
    x = np.arange(1, 5) * 10.0
    y = np.ones((4, 100), dtype=np.float32)
    y[2, 17] = np.nan
    with pytest.raises(ValueError):
        D.curve_fit(D.monoexponential, x, y)


def test_api_errors_and_forms(D):
    """Mirror of tests/core/test_fitting.py error/argument-form checks that need a launch."""
    rng = np.random.default_rng(2)
    x = np.asarray([0.5, 1.0, 2.0, 4.0])
    b = rng.random((10, 10, 20)) + 0.1
    y = [D.MedicalVolume(np.exp(b * t), np.eye(4)) for t in x]
    t = 1 / np.abs(b)
    t_hat = D.MonoExponentialFit(decimal_precision=8).fit(x, y)[0]
    assert np.allclose(t_hat.volume, t)
    t_hat = D.MonoExponentialFit(tc0="polyfit", decimal_precision=8).fit(x, y)[0]
    assert np.allclose(t_hat.volume, t)
    # p0 forms (TestCurveFitter.test_p0)
    for p0 in ((1.0, b), {"a": 1.0, "b": b}, {"a": 1.0, "b": D.MedicalVolume(b, np.eye(4))},
               np.stack([np.ones(b.shape), b], axis=-1)):
        popt, _ = D.CurveFitter(D.monoexponential).fit(x, y, p0=p0)
        assert np.allclose(popt.volume[..., 0], 1.0) and np.allclose(popt.volume[..., 1], b)
    # mask -> NaN outside (TestCurveFitter.test_mask)
    mask = rng.random(b.shape) > 0.5
    popt, r2 = D.CurveFitter(D.monoexponential).fit(x, y, mask=mask)
    assert np.isnan(popt.volume[~mask]).all() and np.allclose(popt.volume[mask][:, 1], b[mask])
    with pytest.raises(TypeError):
        D.CurveFitter(D.monoexponential).fit(x, y, mask="foo")
    with pytest.raises(RuntimeError):
        D.CurveFitter(D.monoexponential).fit(x, y, mask=rng.random((5, 5, 5)) > 0.5)
    # arbitrary python ufunc falls back to host post-processing only (TestCurveFitter.test_out_ufuncs)
    uf = lambda v: 2 * np.abs(v) + 5  # noqa: E731
    popt, _ = D.CurveFitter(D.monoexponential, out_ufuncs=uf).fit(x, y)
    assert np.allclose(popt.volume[..., 1], uf(b))
    # bounds (TestCurveFitter.test_bounds)
    popt, _ = D.CurveFitter(D.monoexponential, out_bounds=[(-np.inf, np.inf), (0, 0.6)]).fit(x, y)
    assert np.isnan(popt.volume[..., 1][b > 0.6001]).all() and np.allclose(popt.volume[..., 1][b < 0.5999], b[b < 0.5999])


def test_float32_result_maps_option(D):
    """out_dtype="f32": the same fit, results as float32 maps (what the kernel computes in), half the D2H bytes."""
    c = G.load("curvefit_mono8_snr100_f32")
    p64, r64 = D.curve_fit(D.monoexponential, c["x"], c["y"], p0=(1.0, -1 / 30))
    p32, r32 = D.curve_fit(D.monoexponential, c["x"], c["y"], p0=(1.0, -1 / 30), out_dtype="f32")
    assert p32.dtype == np.float32 and r32.dtype == np.float32 and p64.dtype == np.float64
    assert np.array_equal(p32, p64.astype(np.float32), equal_nan=True) and np.array_equal(r32, r64.astype(np.float32))
    with pytest.raises(ValueError):
        D.curve_fit(D.monoexponential, c["x"], c["y"], out_dtype="f16")
