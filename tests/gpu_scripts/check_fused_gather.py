"""Multi-GPU check (run under torchrun on >= 2 GPUs): the fused in-kernel all-gather (peer stores over
NVLink, dosma_b200.sharding.PeerMaps) must equal a plain NCCL all-gather of the per-rank results.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/gpu_scripts/check_fused_gather.py [n_voxels]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import dosma_b200 as D  # noqa: E402
from dosma_b200 import device_api as A, sharding  # noqa: E402


def main():
    sizes = [int(v) for v in sys.argv[1:]] or [100_000]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for n in sizes:  # odd sizes take the one-voxel kernel, multiples of 4 the two-voxel TMA kernel (ragged or not)
        ok = check(n, rank, world, dev) and ok
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


def check(n, rank, world, dev):
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = np.arange(1, 9) * 10.0
    xt = torch.tensor(x, device=dev, dtype=torch.float32)[:, None]
    a = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    t2 = 10 + 70 * torch.rand(n, device=dev, generator=g)
    y = a * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device=dev, generator=g)
    ok = True
    # default: the one-voxel kernel's warp-transposed peer stores; use_tma=1 (16-byte-aligned pitch only): the
    # two-voxel TMA kernel with bulk stores of 768-byte row blocks
    for kw in [dict()] + ([dict(use_tma=1)] if n % 4 == 0 else []):
        ok = check_one(n, rank, world, dev, x, y, kw) and ok
    return ok


def check_one(n, rank, world, dev, x, y, kw):
    opts, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), **kw)
    # the NCCL reference must come from the same kernel arithmetic as the fused run: without use_tma=1 the
    # gather runs in the one-voxel kernel, which a plain fit only uses with fast_path=2 (or an odd pitch)
    ref_opts, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), **(kw or dict(fast_path=2)))

    popt, r2 = A.fit_device(ref_opts, P, x, y)
    torch.cuda.synchronize()
    ref = sharding.gather_maps(torch.cat([popt, r2[:, None]], dim=1), [n] * world)

    peer = sharding.PeerMaps(n, P + 1, dev)
    # plumbing check first: a torch copy into every peer's map
    for r in range(world):
        peer.maps[r][rank * n: rank * n + 4, :] = float(rank + 1)
    peer.synchronize()
    for r in range(world):
        assert torch.all(peer.local[r * n: r * n + 4] == float(r + 1)), "peer mapping broken"
    print(f"[rank {rank}] peer mapping ok", flush=True)
    A.fit_device(opts, P, x, y, popt=popt, r2=r2)
    peer.synchronize()
    same = torch.equal(peer.local.nan_to_num(-1.0), ref.nan_to_num(-1.0))
    print(f"[rank {rank}] n={n} {kw} fused gather == nccl all_gather: {same}", flush=True)
    peer.close()
    dist.barrier()
    return same


if __name__ == "__main__":
    main()
