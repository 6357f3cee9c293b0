"""Multi-GPU check (run under torchrun on >= 2 GPUs): the fused in-kernel all-gather (dosma_b200.sharding.PeerMaps:
peer stores over NVLink or NVLS multicast stores) must equal a plain NCCL all-gather of the per-rank results --
all columns and the 8-byte [b or tc, r2] rows, raw parameters and the fused MonoExponentialFit epilogue, the one-voxel
and the two-voxel TMA kernel, ragged sizes; and the masked split-list mode must equal the single-GPU masked fit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/gpu_scripts/check_fused_gather.py [n_voxels ...]
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

POST = dict(ufunc=[0, 1], lb=[-np.inf, 0.0], ub=[np.inf, 100.0], decimals=[-1, 1], r2_threshold=0.9, nan_to_num=0.0)
X = np.arange(1, 9) * 10.0


def synth(n, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    xt = torch.tensor(X, device=dev, dtype=torch.float32)[:, None]
    a = 500 + 1000 * torch.rand(n, device=dev, generator=g)
    t2 = 10 + 70 * torch.rand(n, device=dev, generator=g)
    return a * torch.exp(-xt / t2) + 10 * torch.randn(8, n, device=dev, generator=g)


def main():
    sizes = [int(v) for v in sys.argv[1:]] or [100_000]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for mc in ("auto", "off"):  # NVLS multicast / symmetric memory where available, and the CUDA-IPC peer-store path
        for n in sizes:  # odd sizes take the one-voxel kernel, multiples of 4 the two-voxel TMA kernel (ragged or not)
            ok = check_dense(n, rank, world, dev, mc) and ok
        ok = check_split_list(rank, world, dev, mc) and ok
    for n in sizes:
        ok = check_copy_engine(n, rank, world, dev) and ok
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


def check_dense(n, rank, world, dev, mc):
    y = synth(n, dev, 100 + rank)
    ok = True
    for post in (None, POST):
        for mask_bits, cols in ((0, [0, 1, 2]), (0b10, [1, 2])):  # all columns / [b or tc, r2]
            opts, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post)
            # the NCCL reference must come from the kernel the fused run uses: the two-voxel TMA kernel when the pitch
            # allows it (n % 4 == 0), else the one-voxel kernel (which a plain fit only uses with fast_path=2)
            ref_opts, _ = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post, **({} if n % 4 == 0 else dict(fast_path=2)))
            popt, r2 = A.fit_device(ref_opts, P, X, y)
            torch.cuda.synchronize()
            ref = sharding.gather_maps(torch.cat([popt, r2[:, None]], dim=1)[:, cols].contiguous(), [n] * world)
            peer = sharding.PeerMaps(n, len(cols), dev, param_mask=mask_bits, multicast=mc)
            A.fit_device(opts, P, X, y, popt=popt, r2=r2)
            peer.synchronize()
            same = torch.equal(peer.local.nan_to_num(-1.0), ref.nan_to_num(-1.0))
            print(f"[rank {rank}] n={n} cols={cols} post={post is not None} via {peer.transport}: "
                  f"fused gather == nccl all_gather: {same}", flush=True)
            peer.close()
            dist.barrier()
            ok = ok and same
    return ok


def check_copy_engine(n, rank, world, dev):
    """The third transport: the kernel stores into the own map only, the copy engines push every chunk's rows to the peers
    while the next chunk is fitted (PeerMaps.fit_pipelined).  Chunk boundaries must not change a single bit."""
    if n % 4 != 0:
        return True  # (chunked views need the TMA-aligned pitch of the two-voxel kernel to take the same kernel)
    y = synth(n, dev, 100 + rank)
    ok = True
    for post in (None, POST):
        opts, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post)
        popt, r2 = A.fit_device(opts, P, X, y)
        torch.cuda.synchronize()
        ref = sharding.gather_maps(torch.cat([popt, r2[:, None]], dim=1)[:, [1, 2]].contiguous(), [n] * world)
        peer = sharding.PeerMaps(n, 2, dev, param_mask=0b10, copy_engine=True)
        p2, q2 = torch.empty_like(popt), torch.empty_like(r2)
        for _ in range(2):  # twice: the second fit waits for the first one's copies before it overwrites the rows
            peer.fit_pipelined(lambda lo, hi: A.fit_device(opts, P, X, y[:, lo:hi], popt=p2[lo:hi], r2=q2[lo:hi]), n, chunks=5)
        peer.synchronize()
        same = torch.equal(peer.local.nan_to_num(-1.0), ref.nan_to_num(-1.0)) and torch.equal(p2.nan_to_num(-1.0), popt.nan_to_num(-1.0))
        print(f"[rank {rank}] n={n} post={post is not None} via {peer.transport} ({peer.last_chunks} chunks): "
              f"pipelined copies == nccl all_gather: {same}", flush=True)
        peer.close()
        dist.barrier()
        ok = ok and same
    return ok


def check_split_list(rank, world, dev, mc):
    """Masked fit of one volume by all ranks: everyone holds the whole mask, each fits its share of the compacted voxel
    list from the samples of that share's voxel span only, every rank ends with the complete map."""
    n = 300_032
    y_full = synth(n, dev, 7)  # the same volume on every rank (only the rank's span is handed to the fit)
    g = torch.Generator(device=dev).manual_seed(9)
    zz = torch.arange(n, device=dev, dtype=torch.float32) / n
    mask = (torch.rand(n, device=dev, generator=g) < 0.25 * torch.exp(-((zz - 0.3) / 0.1) ** 2)).to(torch.uint8)  # clustered
    ok = True
    for post, fill in ((None, float("nan")), (POST, 0.0)):
        opts, P = A.make_opts(D.monoexponential, p0=(1.0, -1 / 30), post=post)
        p1, r1 = A.fit_device(opts, P, X, y_full, mask=mask)  # single-GPU masked fit: the expected map
        torch.cuda.synchronize()
        expect = torch.stack([p1[:, 1], r1], dim=1)
        idx = torch.nonzero(mask)[:, 0]
        b = sharding.masked_spans(mask.cpu().numpy(), world)
        v_lo, v_hi = int(b[rank]), int(b[rank + 1])
        share = int(mask[v_lo:v_hi].sum())
        lo, hi = int(mask[:v_lo].sum()), int(mask[:v_lo].sum()) + share
        y_span = y_full[:, v_lo:v_hi].contiguous()
        peer = sharding.PeerMaps(n, 2, dev, total_rows=n, row0=0, param_mask=0b10, split_list=True, fit_span=(v_lo, v_hi),
                                 y_voxel0=v_lo, multicast=mc)
        import ctypes

        from dosma_b200 import _cabi
        lib, h = _cabi.load(), _cabi.get_handle(dev.index)
        xs = np.ascontiguousarray(X, dtype=np.float64)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(lib.dfit_fit_device(h.ptr, ctypes.byref(opts), 8, n, xs.ctypes.data, y_span.data_ptr(), _cabi.F32,
                                        _cabi.PLANAR, y_span.stride(0), mask.data_ptr(), None, _cabi.F32, None, None, _cabi.F32,
                                        None, None, ctypes.c_void_p(stream)))
        peer.synchronize()
        same = torch.equal(peer.local.nan_to_num(-7.0), expect.nan_to_num(-7.0))
        if not same:  # say where: inside / outside the mask, own share / the peers'
            bad = (peer.local.nan_to_num(-7.0) != expect.nan_to_num(-7.0)).any(dim=1)
            inside = mask.bool()
            own = torch.zeros(n, dtype=torch.bool, device=dev)
            own[idx[lo:hi]] = True
            print(f"[rank {rank}] MISMATCH rows: {int(bad.sum())} (inside mask {int((bad & inside).sum())}, own share "
                  f"{int((bad & own).sum())}, outside mask {int((bad & ~inside).sum())}); first: "
                  f"{[(int(i), peer.local[i].tolist(), expect[i].tolist()) for i in torch.nonzero(bad)[:4, 0]]}", flush=True)
        fitted = h.stats()["n_fitted"]
        print(f"[rank {rank}] split list post={post is not None} via {peer.transport}: share {hi - lo} of {idx.numel()} voxels "
              f"(fitted {fitted}), span [{v_lo}, {v_hi}); complete map == single-GPU masked fit: {same}", flush=True)
        peer.close()
        dist.barrier()
        ok = ok and same and fitted <= hi - lo and abs(share - idx.numel() / world) < 8
    return ok


if __name__ == "__main__":
    main()
