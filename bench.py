#!/usr/bin/env python
"""Benchmark of the per-voxel mono-exponential fit (BASELINE.json metric) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one synthetic multi-echo volume.
  value    voxels/s, samples resident in HBM when the timed region starts (device entry point of the
           C-ABI, CUDA events on the launching stream, max over ranks, barrier + sync on both sides)
  e2e      the same metric through the reference-facing host entry point (`dfit_fit_host`): pinned
           HOST buffers in, pinned HOST buffers out, copies inside the timed region
  roofline achieved = algorithmic bytes per launch (4*E + 4*(P+1) per voxel) / mean kernel time
  cpu_baseline  the CPU oracle port (numpy + scipy.optimize.curve_fit per voxel, i.e. what the
           reference does) on a bounded sample of the same workload, all host cores
`--impl reference` times that CPU port alone (rank 0 only) and prints the same JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxels/sec monoexp T2 fit (8 echoes, 384^3)"
UNIT = "voxels/s"
SHAPE = (384, 384, 384)
ECHOES = 8
X_MS = [10.0 * i for i in range(1, ECHOES + 1)]
P0 = (1.0, -1.0 / 30.0)  # MonoExponentialFit default tc0 = 30 (fitting.py:638, 720)
SNR_SIGMA = 10.0  # a ~ U(500, 1500), T2 ~ U(10, 80) ms, sigma = 10 (SNR 100)
BYTES_PER_VOXEL = 4 * ECHOES + 4 * 3  # read y (fp32) + write popt[2] + r2


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload_config(n_gpus):
    return {
        "workload": f"{SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]} x {ECHOES}-echo monoexponential T2 fit, fp32 samples, "
                    f"one such volume per GPU (z-slab shard of a {SHAPE[0]}x{SHAPE[1]}x{SHAPE[2] * n_gpus} volume)",
        "voxels_per_gpu": int(np.prod(SHAPE)),
        "echo_times_ms": X_MS,
        "p0": "a=1, b=-1/30 (MonoExponentialFit default tc0=30)",
        "noise": "gaussian sigma=10 on a~U(500,1500) (SNR 100), T2~U(10,80) ms",
        "l2_policy": "inputs (1.8 GB per GPU) exceed the 126 MB L2; no explicit flush",
        "parallelism": f"voxel-slab x{n_gpus}, one all-gather of the parameter map" if n_gpus > 1 else "single GPU",
    }


def synth_numpy(n, seed, echoes=ECHOES):
    rng = np.random.default_rng(seed)
    x = np.asarray(X_MS[:echoes])
    a = rng.uniform(500, 1500, n)
    t2 = rng.uniform(10, 80, n)
    y = a * np.exp(-x[:, None] / t2) + rng.normal(0, SNR_SIGMA, (echoes, n))
    return x, y.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
_REF = {}


def reference_impl():
    """The CPU implementation the CPU legs time: the UNMODIFIED reference `dosma.core.fitting` from `baseline/_ref`
    (installed there by `pip install --no-deps --target baseline/_ref /root/reference`; loaded
    through tests/golden/ref_loader.py, which stands in for the third-party imports the image lacks) -- kind
    "reference" -- or, where that directory did not travel, the oracle port of the same code -- kind "port"."""
    if not _REF:
        ref_root = os.path.join(ROOT, "baseline", "_ref")
        try:
            if not os.path.isdir(os.path.join(ref_root, "dosma", "core")):
                raise FileNotFoundError(ref_root)
            os.environ["DOSMA_REFERENCE_ROOT"] = ref_root
            import importlib

            R = importlib.import_module("tests.golden.ref_loader")
            R.REFERENCE_ROOT = ref_root
            F, _ = R.load_reference_fitting()
            _REF.update(kind="reference", curve_fit=F.curve_fit, model=F.monoexponential,
                        what="dosma.core.fitting.curve_fit of the unmodified reference (baseline/_ref)")
        except Exception as e:  # noqa: BLE001
            from oracle import dosma_oracle as O

            _REF.update(kind="port", curve_fit=O.curve_fit, model=O.monoexponential,
                        what=f"oracle/dosma_oracle.py (baseline/_ref unavailable: {type(e).__name__})")
    return _REF


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def cpu_port_rate(n_sample, workers, seed=1234):
    impl = reference_impl()
    x, y = synth_numpy(n_sample, seed)
    t0 = time.perf_counter()
    popt, r2 = impl["curve_fit"](impl["model"], x, y, p0=P0, num_workers=workers if workers > 1 else 0,
                                 chunksize=max(64, min(1000, n_sample // (4 * max(workers, 1)))))
    dt = time.perf_counter() - t0
    assert popt.shape == (n_sample, 2)
    return n_sample / dt, dt


def cpu_baseline(target_seconds=12.0):
    cores = host_cores()
    impl = reference_impl()
    rate, _ = cpu_port_rate(256 * cores, cores)  # calibration (dominated by pool start-up: a lower bound)
    n = int(min(max(rate * target_seconds, 1024), 2_000_000))
    rate, dt = cpu_port_rate(n, cores)
    if dt < 0.6 * target_seconds:  # the calibration under-estimated the rate: one more run at the right size
        n = int(min(max(rate * target_seconds, 1024), 2_000_000))
        rate, dt = cpu_port_rate(n, cores)
    # BASELINE.json's configs[0] names the reference's num_workers=1 case: time that too (a few seconds)
    n1 = int(min(max(rate / max(cores, 1) * 3.0, 512), 65536))
    rate1, dt1 = cpu_port_rate(n1, 1, seed=4321)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": impl["kind"],
            "sample": f"{n} voxels of the same workload (seed 1234), scipy.optimize.curve_fit per voxel via "
                      f"{impl['what']} with a {cores}-process pool, {dt:.1f} s",
            "single_core_value": rate1, "single_core_sample": f"{n1} voxels, num_workers=1, {dt1:.1f} s"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = host_cores()
    impl = reference_impl()
    rate, _ = cpu_port_rate(256 * cores, cores)  # dominated by pool start-up: a lower bound
    rate, _ = cpu_port_rate(int(min(max(rate * 3.0, 2048), 200_000)), cores)  # second, larger calibration sample
    per_step = int(min(max(rate * 8.0, 1024), 1_000_000))  # ~8 s per step
    for _ in range(args.warmup):
        cpu_port_rate(max(per_step // 8, 512), cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_port_rate(per_step, cores, seed=1234 + k)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = (f"{per_step} voxels per step of the same workload, scipy.optimize.curve_fit per voxel "
              f"({impl['what']}, {cores}-process pool)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": impl["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and the threads / page-locked allocations it makes from here on) to the CPUs next to its GPU:
    with one rank per GPU, host staging and the copy threads of `dfit_fit_host` then stay on the socket the GPU's PCIe
    root hangs off.  Returns a short description for the JSON line, or None when the topology cannot be read."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return f"{len(allowed)} CPUs local to GPU {local_rank} ({spec})"
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """DRAM bytes per voxel of the fit kernel from the committed `ncu --set full` capture of this same command
    (profiles/traffic.json names the capture), or (None, None).  A profiler cannot run inside a timed bench, so the
    figure is the committed measurement, not one of this run."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                t = json.load(f)
            return t.get("dram_bytes_per_voxel"), t.get("source")
        except Exception:
            return None, None
    return None, None


def other_configs(torch, D, A, _cabi, dev, handle):
    """BASELINE.json configs[2] and [3] on one GPU, samples in HBM, CUDA events on the launching stream (median of 5 after
    2 warm-ups): config 3 = 512 x 512 x 256, 7-echo T1rho (non-uniform spin-lock times) with a 2.3 % tissue mask, through
    the fused MonoExponentialFit epilogue ([tc, r2] maps); config 4 = 256 x 256 x 128, 16-echo bi-exponential, fp32."""
    def timed(fn):
        fn()
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    out = {}
    g = torch.Generator(device=dev).manual_seed(2)
    x7 = [0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]
    shape = (512, 512, 256)
    n = int(np.prod(shape))
    xt = torch.tensor(x7, device=dev, dtype=torch.float32)[:, None]
    y = (500 + 1000 * torch.rand(n, device=dev, generator=g)) * torch.exp(-xt / (20 + 100 * torch.rand(n, device=dev, generator=g)))
    y += 10 * torch.randn(7, n, device=dev, generator=g)
    zz, yy, xx = torch.meshgrid(*[torch.linspace(-1, 1, s, device=dev) for s in shape], indexing="ij")
    rad = (zz ** 2 + yy ** 2 + (xx * 1.6) ** 2).sqrt()
    mask = ((rad > 0.55) & (rad < 0.62)).reshape(-1)
    del zz, yy, xx, rad
    post = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, 3], r2_threshold=0.9, nan_to_num=0.0)
    o, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post, out_param=1)
    tc = torch.empty((n,), device=dev)
    r2 = torch.empty((n,), device=dev)
    ms = timed(lambda: A.fit_device(o, P, x7, y, mask=mask, popt=tc, r2=r2, handle=handle))
    nm = int(mask.sum())
    out["config3_512x512x256_7echo_t1rho_tissue_mask"] = {
        "ms": ms, "voxels": n, "masked_voxels": nm, "volume_voxels_per_s": n / ms * 1e3, "masked_voxels_per_s": nm / ms * 1e3,
        "fitted_voxels": handle.stats()["n_fitted"], "outputs": "[tc, r2] maps, fused MonoExponentialFit epilogue"}
    ms = timed(lambda: A.fit_device(o, P, x7, y, popt=tc, r2=r2, handle=handle))
    out["config3_dense_no_mask"] = {"ms": ms, "voxels_per_s": n / ms * 1e3}
    del y, mask, tc, r2

    g = torch.Generator(device=dev).manual_seed(3)
    x16 = [5.0 * i for i in range(1, 17)]
    n = 256 * 256 * 128
    xt = torch.tensor(x16, device=dev, dtype=torch.float32)[:, None]
    amp = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    fs = 0.3 + 0.4 * torch.rand(n, device=dev, generator=g)
    ts = 8 + 12 * torch.rand(n, device=dev, generator=g)
    tl = 50 + 50 * torch.rand(n, device=dev, generator=g)
    y = amp * fs * torch.exp(-xt / ts) + amp * (1 - fs) * torch.exp(-xt / tl)
    o, P = A.make_opts(D.biexponential, p0=(500.0, -1 / 10, 500.0, -1 / 60))
    popt = torch.empty((n, 4), device=dev)
    r2 = torch.empty((n,), device=dev)
    for name, data in (("config4_256x256x128_16echo_biexp_f32", y), ("config4_at_snr100", y + 10 * torch.randn(16, n, device=dev, generator=g))):
        ms = timed(lambda: A.fit_device(o, P, x16, data, popt=popt, r2=r2, handle=handle))
        st = handle.stats()
        out[name] = {"ms": ms, "voxels": n, "voxels_per_s": n / ms * 1e3, "mean_passes": st["sum_iters"] / max(st["n_fitted"], 1),
                     "failed_fraction": st["n_failed"] / n}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import dosma_b200 as D
    from dosma_b200 import _cabi, device_api as A, sharding

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    # stdout carries exactly one JSON line: anything libraries print on the way (NCCL's version banner ...) goes
    # to stderr -- file descriptor 1 is pointed at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = int(np.prod(SHAPE))
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    xt = torch.tensor(X_MS, device=dev, dtype=torch.float32)[:, None]
    a = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    t2 = 10 + 70 * torch.rand(n, device=dev, generator=g)
    y = a * torch.exp(-xt / t2)
    y += SNR_SIGMA * torch.randn(ECHOES, n, device=dev, generator=g)
    del a, t2
    opts, P = A.make_opts(D.monoexponential, p0=P0)
    handle = _cabi.get_handle(local_rank)
    popt = torch.empty((n, P), dtype=torch.float32, device=dev)
    r2 = torch.empty((n,), dtype=torch.float32, device=dev)
    counts = [n] * world

    # N > 1: the reassembly of the T2 map is fused into the fit kernel's epilogue -- each voxel's [b, r2] row (what a
    # T2 map needs of the fit: 8 bytes) is stored straight into every rank's map, by one NVLS multicast store that the
    # NVSwitch replicates where torch's symmetric memory provides a multicast mapping, else by one peer store per rank
    # over NVLink (dosma_b200.sharding.PeerMaps).  If neither mapping is available the same map is produced by a plain
    # NCCL all-gather after the fit.  The local [a, b] / r2 maps are written as at N = 1: same workload per GPU.
    GATHER_COLS = [1, 2]  # b, r2
    peer = None
    gather_mode = "none"
    if world > 1:
        try:
            # Peer stores by default: a multicast store also delivers the rank's OWN copy through the switch, so every
            # GPU receives N instead of N - 1 maps -- measured slower (2 GPUs: 1.42 against 0.83 ms per step).
            # DFIT_BENCH_MULTICAST=auto selects the NVLS multicast path where torch's symmetric memory offers it.
            transport = os.environ.get("DFIT_BENCH_GATHER", "stores")  # "stores" (fused epilogue) | "copies" (copy engines)
            peer = sharding.PeerMaps(n, len(GATHER_COLS), dev, param_mask=0b10, copy_engine=transport == "copies",
                                     multicast=os.environ.get("DFIT_BENCH_MULTICAST", "off"))
            gather_mode = (f"all-gather of [b, r2] rows: {peer.transport}" if peer.copy_engine
                           else f"fused in-kernel all-gather of [b, r2] rows: {peer.transport}")
        except Exception as e:  # pragma: no cover
            peer = None
            gather_mode = f"nccl all_gather_into_tensor (peer mapping unavailable: {type(e).__name__})"
        flag = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and peer is not None:
            peer.close()
            peer = None
            gather_mode = "nccl all_gather_into_tensor (peer mapping unavailable on some rank)"

    def packed():
        return torch.cat([popt[:, 1:2], r2[:, None]], dim=1)

    def fit_chunk(lo, hi):
        A.fit_device(opts, P, X_MS, y[:, lo:hi], popt=popt[lo:hi], r2=r2[lo:hi], handle=handle)

    def fit_once():
        if peer is not None and peer.copy_engine:
            peer.fit_pipelined(fit_chunk, n)
        else:
            A.fit_device(opts, P, X_MS, y, popt=popt, r2=r2, handle=handle)

    def step():
        fit_once()
        if world > 1 and peer is None:
            return sharding.gather_maps(packed(), counts)
        return popt

    if peer is not None:  # one-off check of the fused gather against NCCL
        step()
        peer.synchronize()
        ref = sharding.gather_maps(packed(), counts)
        same = torch.equal(peer.local.nan_to_num(-1.0), ref.nan_to_num(-1.0))
        del ref
        if not same:
            raise SystemExit("fused all-gather disagrees with NCCL all-gather")

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    ev[0].record()
    for k in range(args.steps):
        kev[k][0].record()
        fit_once()
        kev[k][1].record()
        if world > 1 and peer is None:
            sharding.gather_maps(packed(), counts)
        ev[k + 1].record()
    sync()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = float(np.mean([s.elapsed_time(e) for s, e in kev]))
    stats = handle.stats()

    # the same step back to back for >= 1 s: the figure under sustained clocks / power, next to the burst of K steps
    # above (and long enough for nvidia-smi to see the clocks under load)
    n_sus = int(min(max(1200.0 / max(total_ms / args.steps, 1e-3), args.steps), 20000))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    s0.record()
    for _ in range(n_sus):
        step()
    s1.record()
    sync()
    sustained_ms = s0.elapsed_time(s1)
    clocks = sampler.stop() if rank == 0 else None
    peer_chunks = getattr(peer, "last_chunks", None) if peer is not None else None
    if peer is not None:
        peer.close()

    # ---- end-to-end through the host entry point (pinned host buffers) -------------------------
    yh = torch.empty((ECHOES, n), dtype=torch.float32).pin_memory()
    yh.copy_(y)
    popt_h = torch.empty((n, P), dtype=torch.float32).pin_memory()
    r2_h = torch.empty((n,), dtype=torch.float32).pin_memory()
    import ctypes

    lib = _cabi.load()
    xs = np.asarray(X_MS, dtype=np.float64)
    planes = (ctypes.c_void_p * ECHOES)(*[yh[e].data_ptr() for e in range(ECHOES)])

    def e2e_step():
        _cabi.check(lib.dfit_fit_host(handle.ptr, ctypes.byref(opts), ECHOES, n, xs.ctypes.data,
                                      ctypes.cast(planes, ctypes.c_void_p), _cabi.F32, None, None, _cabi.F32,
                                      popt_h.data_ptr(), r2_h.data_ptr(), _cabi.F32, None, None))
        return float(r2_h[n // 2])  # host read of the step's result

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync()
    e2e_s = time.perf_counter() - t0
    e2e_launches = handle.stats()["n_launches"]

    # the same step through the Python drop-in API on pageable numpy data (what a DOSMA user calls), N = 1 only
    py_api = None
    if world == 1:
        y_np = yh.numpy().copy()  # pageable
        D.curve_fit(D.monoexponential, xs, y_np, p0=P0)  # warm-up at full size: the staging blocks get their size
        t0 = time.perf_counter()
        for _ in range(2):
            p_np, r_np = D.curve_fit(D.monoexponential, xs, y_np, p0=P0)
        py_s = (time.perf_counter() - t0) / 2
        py_api = {"value": n / py_s, "unit": UNIT, "api": "dosma_b200.curve_fit on pageable numpy arrays, float64 results",
                  "seconds_per_step": py_s}
        del p_np, r_np
        # ... and what every DOSMA pipeline calls: MonoExponentialFit.fit on a list of volumes (tc0='polyfit' like
        # cube_quant.py:172 / mapss.py:172), float64 T2 and r2 maps
        vols = [D.MedicalVolume(y_np[e].reshape(SHAPE), np.eye(4)) for e in range(ECHOES)]
        fitter = D.MonoExponentialFit(tc0="polyfit", decimal_precision=1)
        fitter.fit(xs, vols)
        t0 = time.perf_counter()
        for _ in range(2):
            tc_v, r2_v = fitter.fit(xs, vols)
        mf_s = (time.perf_counter() - t0) / 2
        py_api["monoexponentialfit"] = {"value": n / mf_s, "unit": UNIT, "seconds_per_step": mf_s,
                                        "api": "dosma_b200.MonoExponentialFit(tc0='polyfit').fit on volumes, float64 maps"}
        del y_np, vols, tc_v, r2_v

    # the other single-GPU configurations of BASELINE.json, device-resident like `value` (N = 1 only; reported beside the
    # headline, never part of it -- a failure here must not cost the line)
    other = None
    if world == 1:
        try:
            other = other_configs(torch, D, A, _cabi, dev, handle)
        except Exception as e:  # pragma: no cover
            other = {"error": f"{type(e).__name__}: {e}"[:200]}

    times = torch.tensor([total_ms, e2e_s * 1e3, kernel_ms, sustained_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms, sustained_ms = [float(t) for t in times.tolist()]

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * n / (ms_per_step * 1e-3)
        peak, peak_src = peak_hbm()
        achieved = n * BYTES_PER_VOXEL / (kernel_ms * 1e-3) / 1e9
        tpv, traffic_src = recorded_traffic()
        hbm = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
               "traffic": tpv * n if tpv else None, "traffic_source": traffic_src, "peak_source": peak_src,
               "kernel": ("dfit::fit_kernel_mono2_tma<MonoExp,8,GATHER=false,float>" if world == 1 else
                          "dfit::fit_kernel_mono2_tma<MonoExp,8,GATHER=true,float>" if peer is not None else
                          "dfit::fit_kernel_mono2_tma<MonoExp,8,GATHER=false,float> + ncclAllGather"),
               "kernel_ms": kernel_ms, "algorithmic_bytes_per_voxel": BYTES_PER_VOXEL,
               "note": "variable-projection Newton on q = exp(b dx), two voxels per lane, straight-line two-pass fit, "
                       "tiles staged by TMA; FP32-pipe / issue bound, HBM fraction is the contract figure"}
        if world == 1:
            roofline = hbm
        else:
            # N > 1: the step is bound by what every GPU must RECEIVE -- (N - 1) ranks' rows of the reassembled map --
            # against the peer-copy rate measured on this pool (B200_PROFILING.md: 770 GB/s per direction per GPU)
            ingress = (world - 1) * n * 4 * len(GATHER_COLS)
            nv = ingress / (ms_per_step * 1e-3) / 1e9
            roofline = {"bound": "nvlink", "achieved": nv, "peak": 770.0, "unit": "GB/s", "frac": nv / 770.0,
                        "traffic": None, "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md)",
                        "ingress_bytes_per_gpu_per_step": ingress, "row_bytes": 4 * len(GATHER_COLS),
                        "kernel": hbm["kernel"], "kernel_ms": kernel_ms,
                        "note": "fused compute + collective: target time = max(fit at the HBM roofline, (N-1) x map bytes / 770 GB/s); "
                                "the HBM view of the same kernel is in roofline_hbm",
                        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "gather": gather_mode,
            "clocks": clocks,
            "sustained": {"value": world * n * n_sus / (sustained_ms * 1e-3), "unit": UNIT, "steps": n_sus,
                          "seconds": sustained_ms * 1e-3, "ms_per_step": sustained_ms / n_sus,
                          "note": "the same step back to back for >= 1 s (the clock samples cover this region too)"},
            "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": 4 * ECHOES * n, "d2h_bytes_per_step": 4 * (P + 1) * n,
                    "steps": e2e_steps, "launches_per_step": e2e_launches, "host_affinity": numa,
                    "api": "dfit_fit_host (C-ABI), pinned host buffers in and out"},
            "gpu_launches": args.steps * stats["n_launches"] * (peer_chunks or 1),
            "roofline": roofline,
            "lm": {"mean_iters": stats["sum_iters"] / max(stats["n_fitted"], 1), "max_iters": stats["max_iters"],
                   "failed_voxels": stats["n_failed"], "fitted_voxels": stats["n_fitted"]},
        }
        if world > 1:
            line["roofline_hbm"] = hbm
        if py_api is not None:
            line["e2e_python_api"] = py_api
        if other is not None:
            line["other_configs"] = other
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dfit", choices=["dfit", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "dfit" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
