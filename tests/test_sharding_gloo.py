"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: voxel-range partition + the single
all-gather that reassembles the parameter map, including ragged slabs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dosma_b200 import sharding as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_vox, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bounds = S.voxel_ranges(n_vox, world, align=128)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        y_local = torch.arange(lo, hi, dtype=torch.float32)

        def fake_fit(y):  # stands in for the CUDA fit: "parameters" derived from the voxel id
            return torch.stack([y * 2, -y], dim=1), y + 0.5

        full = S.fit_sharded(fake_fit, y_local, counts=np.diff(bounds))
        ids = torch.arange(n_vox, dtype=torch.float32)
        expect = torch.stack([ids * 2, -ids, ids + 0.5], dim=1)
        assert full.shape == expect.shape and torch.equal(full, expect)
        torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_vox", [1024, 1000, 130])
def test_gather_maps_two_ranks(tmp_path, n_vox):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_vox, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, "r0.pt"))
    b = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(a, b) and a.shape == (n_vox, 3)
