"""BASELINE.json configurations 3 and 5 as written -- sharded over the GPUs of one box, each with the final gather of
the map fused into the fit (run under torchrun; N = 1 works too):

  C3  512 x 512 x 256, 7-echo T1rho mono-exponential with a tissue mask: ONE volume, all ranks scan the whole mask,
      each fits the masked voxels of its span (spans balanced on the mask, dosma_b200.sharding.masked_spans) from the
      samples of that span only, every rank ends with the complete [tc, r2] map (fill outside the mask).
  C5  32 subjects x 384 x 384 x 64 x 8 echoes: 32 / N whole subjects per GPU, [b, r2] rows of every subject gathered
      on every rank.  (Strong scaling: the batch is fixed.)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
        tests/gpu_scripts/configs_multi.py > gpurun_out/configs_multi_Ngpu.json
Prints one JSON object (rank 0).  Times: CUDA events on the launching stream, barrier + sync on both sides, max over ranks.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dosma_b200 as D  # noqa: E402
from dosma_b200 import _cabi, device_api as A, sharding  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29551")
dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
MC = os.environ.get("DFIT_BENCH_MULTICAST", "off")
lib, h = _cabi.load(), _cabi.get_handle(local)
REPS = 10


def sync():
    dist.barrier()
    torch.cuda.synchronize()


def timed(fn):
    for _ in range(3):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / REPS], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def synth(n, x, seed, t_rng):
    g = torch.Generator(device=dev).manual_seed(seed)
    xt = torch.tensor(x, device=dev, dtype=torch.float32)[:, None]
    a = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    t = t_rng[0] + (t_rng[1] - t_rng[0]) * torch.rand(n, device=dev, generator=g)
    return a * torch.exp(-xt / t) + 10.0 * torch.randn(len(x), n, device=dev, generator=g)


out = {"n_gpus": world, "multicast": MC}

# ---------------------------------------------------------------- C3
x7 = [0.0, 10.0, 12.847, 25.695, 40.0, 51.39, 80.0]  # MAPSS spin-lock / echo times (tests/scan_sequences/mri/test_mapss.py:43)
shape = (512, 512, 256)
n = int(np.prod(shape))
zz, yy, xx = torch.meshgrid(*[torch.linspace(-1, 1, s, device=dev) for s in shape], indexing="ij")
rad = (zz ** 2 + yy ** 2 + (xx * 1.6) ** 2).sqrt()
mask = ((rad > 0.55) & (rad < 0.62)).reshape(-1).to(torch.uint8)  # an ellipsoid shell: thin, z-clustered tissue
del zz, yy, xx, rad
spans = sharding.masked_spans(mask.cpu().numpy(), world)
v_lo, v_hi = int(spans[rank]), int(spans[rank + 1])
y_span = synth(v_hi - v_lo, x7, 200 + rank, (20, 120))
post = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 500.0], decimals=[-1, 1], r2_threshold=0.9, nan_to_num=0.0)
o3, P = A.make_opts(D.monoexponential, init="polyfit", post=post)  # MonoExponentialFit(tc0="polyfit") as in mapss.py:172
peer = sharding.PeerMaps(n, 2, dev, total_rows=n, row0=0, param_mask=0b10, split_list=True, fit_span=(v_lo, v_hi),
                         y_voxel0=v_lo, multicast=MC)
xs = np.ascontiguousarray(x7, dtype=np.float64)
stream = torch.cuda.current_stream(dev).cuda_stream


def c3_step():
    _cabi.check(lib.dfit_fit_device(h.ptr, ctypes.byref(o3), 7, n, xs.ctypes.data, y_span.data_ptr(), _cabi.F32, _cabi.PLANAR,
                                    y_span.stride(0), mask.data_ptr(), None, _cabi.F32, None, None, _cabi.F32, None, None,
                                    ctypes.c_void_p(stream)))


ms = timed(c3_step)
peer.synchronize()
fitted = torch.tensor([h.stats()["n_fitted"]], device=dev)
dist.all_reduce(fitted)
m = mask.bool()
tc_in, r2_in = peer.local[m, 0], peer.local[m, 1]
complete = bool((peer.local[~m] == 0).all()) and float((tc_in > 0).float().mean()) > 0.9 and float((r2_in > 0.9).float().mean()) > 0.9
flag = torch.tensor([1 if complete else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
n_mask = int(m.sum())
out["C3_512x512x256_7echo_t1rho_mask_split"] = {
    "voxels": n, "masked_voxels": n_mask, "mask_fraction": n_mask / n, "fitted_voxels_all_ranks": int(fitted),
    "span_voxels_this_rank0": v_hi - v_lo, "ms": ms, "volume_voxels_per_s": n / ms * 1e3, "masked_voxels_per_s": n_mask / ms * 1e3,
    "every_rank_has_the_complete_map": bool(int(flag)), "transport": peer.transport,
    "nvlink_ingress_bytes_per_gpu": (world - 1) * n_mask * 8 // max(world, 1)}
peer.close()
del y_span, mask
torch.cuda.empty_cache()

# ---------------------------------------------------------------- C5
x8 = [10.0 * i for i in range(1, 9)]
subj = 384 * 384 * 64
n_subj = 32
per = n_subj // world
n_loc = per * subj
y5 = synth(n_loc, x8, 300 + rank, (10, 80))
o5, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30))
popt = torch.empty((n_loc, 2), device=dev)
r2 = torch.empty((n_loc,), device=dev)
peer = sharding.PeerMaps(n_loc, 2, dev, param_mask=0b10, multicast=MC) if world > 1 else None
ms = timed(lambda: A.fit_device(o5, P, x8, y5, popt=popt, r2=r2, handle=h))
ok = True
if peer is not None:
    peer.synchronize()
    ref = sharding.gather_maps(torch.cat([popt[:, 1:2], r2[:, None]], dim=1), [n_loc] * world)
    ok = torch.equal(peer.local.nan_to_num(-1.0), ref.nan_to_num(-1.0))
    del ref
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
out["C5_32subjects_384x384x64_8echo"] = {
    "voxels": n_subj * subj, "subjects_per_gpu": per, "ms": ms, "voxels_per_s": n_subj * subj / ms * 1e3,
    "gather_equals_nccl_all_gather": bool(int(flag)), "transport": peer.transport if peer is not None else "none (one GPU)",
    "nvlink_ingress_bytes_per_gpu": (world - 1) * n_loc * 8,
    "nvlink_ingress_GBps": (world - 1) * n_loc * 8 / ms / 1e6}
if peer is not None:
    peer.close()
if rank == 0:
    print(json.dumps(out, indent=1), flush=True)
dist.barrier()
dist.destroy_process_group()
