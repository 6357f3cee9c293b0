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
def cpu_port_rate(n_sample, workers, seed=1234):
    from oracle import dosma_oracle as O

    x, y = synth_numpy(n_sample, seed)
    t0 = time.perf_counter()
    popt, r2 = O.curve_fit(O.monoexponential, x, y, p0=P0, num_workers=workers if workers > 1 else 0,
                           chunksize=max(64, min(1000, n_sample // (4 * max(workers, 1)))))
    dt = time.perf_counter() - t0
    assert popt.shape == (n_sample, 2)
    return n_sample / dt, dt


def cpu_baseline(target_seconds=12.0):
    from oracle import dosma_oracle as O

    cores = O.host_cores()
    rate, _ = cpu_port_rate(256 * cores, cores)  # calibration (dominated by pool start-up: a lower bound)
    n = int(min(max(rate * target_seconds, 1024), 2_000_000))
    rate, dt = cpu_port_rate(n, cores)
    if dt < 0.6 * target_seconds:  # the calibration under-estimated the rate: one more run at the right size
        n = int(min(max(rate * target_seconds, 1024), 2_000_000))
        rate, dt = cpu_port_rate(n, cores)
    # BASELINE.json's configs[0] names the reference's num_workers=1 case: time that too (a few seconds)
    n1 = int(min(max(rate / max(cores, 1) * 3.0, 512), 65536))
    rate1, dt1 = cpu_port_rate(n1, 1, seed=4321)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} voxels of the same workload (seed 1234), scipy.optimize.curve_fit per voxel via "
                      f"oracle/dosma_oracle.py with a {cores}-process pool, {dt:.1f} s",
            "single_core_value": rate1, "single_core_sample": f"{n1} voxels, num_workers=1, {dt1:.1f} s"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import dosma_oracle as O

    cores = O.host_cores()
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
              f"(oracle/dosma_oracle.py, {cores}-process pool)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
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
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
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
    """dram bytes per launch of the fit kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                t = json.load(f)
            return t.get("dram_bytes_per_voxel")
        except Exception:
            return None
    return None


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

    # N > 1: the reassembly of the parameter map is fused into the fit kernel's epilogue -- each voxel's
    # [a, b, r2] row is stored straight into every rank's map over NVLink (dosma_b200.sharding.PeerMaps).
    # If peer mapping is unavailable the same result is produced by a plain NCCL all-gather.
    peer = None
    gather_mode = "none"
    if world > 1:
        try:
            peer = sharding.PeerMaps(n, P + 1, dev)
            gather_mode = "fused peer stores over NVLink (in-kernel all-gather)"
        except Exception as e:  # pragma: no cover
            peer = None
            gather_mode = f"nccl all_gather_into_tensor (peer mapping unavailable: {type(e).__name__})"
        flag = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and peer is not None:
            peer.close()
            peer = None
            gather_mode = "nccl all_gather_into_tensor (peer mapping unavailable on some rank)"

    def step():
        A.fit_device(opts, P, X_MS, y, popt=popt, r2=r2, handle=handle)
        if world > 1 and peer is None:
            packed = torch.cat([popt, r2[:, None]], dim=1)
            return sharding.gather_maps(packed, counts)
        return popt

    if peer is not None:  # one-off check of the fused gather against NCCL
        step()
        peer.synchronize()
        ref = sharding.gather_maps(torch.cat([popt, r2[:, None]], dim=1), counts)
        same = torch.equal(peer.local.nan_to_num(-1.0), ref.nan_to_num(-1.0))
        del ref
        if not same:
            raise SystemExit("fused all-gather disagrees with NCCL all-gather")

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    ev[0].record()
    for k in range(args.steps):
        kev[k][0].record()
        A.fit_device(opts, P, X_MS, y, popt=popt, r2=r2, handle=handle)
        kev[k][1].record()
        if world > 1 and peer is None:
            packed = torch.cat([popt, r2[:, None]], dim=1)
            sharding.gather_maps(packed, counts)
        ev[k + 1].record()
    sync()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = float(np.mean([s.elapsed_time(e) for s, e in kev]))
    stats = handle.stats()
    clocks = sampler.stop() if rank == 0 else None
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
        del y_np, p_np, r_np

    times = torch.tensor([total_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms = [float(t) for t in times.tolist()]

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * n / (ms_per_step * 1e-3)
        peak, peak_src = peak_hbm()
        achieved = n * BYTES_PER_VOXEL / (kernel_ms * 1e-3) / 1e9
        tpv = recorded_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), gather=gather_mode),
            "clocks": clocks,
            "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": 4 * ECHOES * n, "d2h_bytes_per_step": 4 * (P + 1) * n,
                    "steps": e2e_steps, "launches_per_step": e2e_launches,
                    "api": "dfit_fit_host (C-ABI), pinned host buffers in and out"},
            "gpu_launches": args.steps * stats["n_launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": tpv * n if tpv else None, "peak_source": peak_src,
                         "kernel": ("dfit::fit_kernel_mono2_tma<MonoExp,8>" if world == 1 else
                                    "dfit::fit_kernel<MonoExp,float,8,true,GATHER>"),
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_voxel": BYTES_PER_VOXEL,
                         "note": "variable-projection Newton on q = exp(b dx), two voxels per lane, tiles staged by "
                                 "TMA; FP32-pipe / issue bound, HBM fraction is the contract figure"},
            "lm": {"mean_iters": stats["sum_iters"] / max(stats["n_fitted"], 1), "max_iters": stats["max_iters"],
                   "failed_voxels": stats["n_failed"], "fitted_voxels": stats["n_fitted"]},
        }
        if py_api is not None:
            line["e2e_python_api"] = py_api
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
