"""Multi-GPU GPU test (skipped on a 1-GPU box): fused in-kernel all-gather == NCCL all-gather."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_gather_matches_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "gpu_scripts", "check_fused_gather.py"), "200003", "200192", "200068"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    # 2 transports x 3 sizes x (raw, fused epilogue) x (all columns, 8-byte rows); 2 transports x 2 split-list runs
    assert res.stdout.count("fused gather == nccl all_gather: True") == 2 * 3 * 4 * world, res.stdout[-3000:]
    assert res.stdout.count("complete map == single-GPU masked fit: True") == 2 * 2 * world, res.stdout[-3000:]
    # the copy-engine transport: the two sizes with a TMA-aligned pitch x (raw, fused epilogue)
    assert res.stdout.count("pipelined copies == nccl all_gather: True") == 2 * 2 * world, res.stdout[-3000:]
