"""Multi-GPU plumbing: one process per GPU, z-slab / voxel-range partition, ONE all-gather at the end.

The fit is embarrassingly parallel over voxels (the reference's only parallelism is a process pool
over voxels, dosma/core/fitting.py:861-868), so ranks never exchange data during the fit; the only
collective is the final reassembly of the parameter map.  Ragged slabs are
padded to the largest slab so that the reassembly stays a single `all_gather_into_tensor`.
"""
import numpy as np

__all__ = ["slab_bounds", "voxel_ranges", "gather_maps", "fit_sharded", "PeerMaps"]


def slab_bounds(n_slices, world_size):
    """Contiguous, balanced z-slab boundaries: returns world_size + 1 slice indices."""
    base, rem = divmod(int(n_slices), int(world_size))
    sizes = [base + (1 if r < rem else 0) for r in range(world_size)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def voxel_ranges(n_vox, world_size, align=1):
    """Balanced contiguous voxel ranges (multiples of `align` except the last) -- the partition to use
    on the flattened / mask-compacted voxel list, which also balances z-clustered tissue masks."""
    n_units = (int(n_vox) + align - 1) // align
    b = slab_bounds(n_units, world_size) * align
    b[-1] = n_vox
    return np.minimum(b, n_vox)


def gather_maps(local, counts, group=None):
    """All-gather per-rank row blocks `local` ([n_r, C] tensors, n_r = counts[rank]) into the full
    [sum(counts), C] tensor on every rank with a single collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    counts = [int(c) for c in counts]
    nmax = max(counts)
    cols = local.shape[1:]
    if local.shape[0] != nmax:
        padded = local.new_zeros((nmax,) + tuple(cols))
        padded[: local.shape[0]] = local
    else:
        padded = local.contiguous()
    out = local.new_empty((world * nmax,) + tuple(cols))
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(c == nmax for c in counts):
        return out
    return torch.cat([out[r * nmax: r * nmax + counts[r]] for r in range(world)], dim=0)


def fit_sharded(fit_local, y_local, counts, group=None):
    """Run `fit_local(y_local) -> (popt [n, P], r2 [n])` on this rank's slab and reassemble
    `[N, P + 1]` (parameters and r2 side by side) on every rank."""
    import torch

    popt, r2 = fit_local(y_local)
    packed = torch.cat([popt, r2[:, None]], dim=1)
    return gather_maps(packed, counts, group=group)


class _DevArray:
    """Zero-copy torch view of a raw device allocation (via __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerMaps:
    """Reassembled parameter maps that every rank's fit kernel stores into directly (fused all-gather).

    Each rank allocates its own full map `[world * rows_per_rank, P + 1]` (fp32) through the C-ABI
    (`dfit_ipc_alloc`) and publishes the CUDA IPC handle; every rank maps all peers' allocations with
    ITS device current (`dfit_ipc_open`: lazy peer access over NVLink) and hands the `world` device
    pointers to `dfit_set_gather`, after which `fit_device` stores each voxel's row into all maps
    while the fit is running.  `synchronize()` (stream drain + barrier) makes the local map complete.
    """

    def __init__(self, rows_per_rank, ncols, device, group=None):
        import ctypes

        import torch
        import torch.distributed as dist

        from . import _cabi

        lib = _cabi.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows_per_rank = int(rows_per_rank)
        self.device = device
        self._handle = _cabi.get_handle(device.index)
        shape = (self.world * self.rows_per_rank, int(ncols))
        nbytes = shape[0] * shape[1] * 4
        own = ctypes.c_void_p()
        hbuf = ctypes.create_string_buffer(64)
        _cabi.check(lib.dfit_ipc_alloc(self._handle.ptr, nbytes, ctypes.byref(own), hbuf))
        self._own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(hbuf.raw), group=group)
        self._peer_ptrs = []
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(own.value)
            else:
                p = ctypes.c_void_p()
                _cabi.check(lib.dfit_ipc_open(self._handle.ptr, handles[r], ctypes.byref(p)))
                self._peer_ptrs.append(p)
                ptrs.append(p.value)
        with torch.cuda.device(device):
            self.maps = [torch.as_tensor(_DevArray(p, shape), device=device) for p in ptrs]
        self.local = self.maps[self.rank]
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        _cabi.check(lib.dfit_set_gather(self._handle.ptr, self.world, self.rank,
                                        ctypes.cast(self._ptrs, ctypes.c_void_p), self.rows_per_rank))
        dist.barrier(group=group)

    def synchronize(self):
        """All ranks' peer stores into this rank's map are complete after this returns."""
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def close(self):
        import torch
        import torch.distributed as dist

        from . import _cabi

        lib = _cabi.load()
        torch.cuda.synchronize(self.device)
        _cabi.check(lib.dfit_set_gather(self._handle.ptr, 0, 0, None, 0))
        self.maps = []
        self.local = None
        dist.barrier(group=self.group)  # nobody stores into a map that is about to go away
        for p in self._peer_ptrs:
            _cabi.check(lib.dfit_ipc_close(self._handle.ptr, p))
        self._peer_ptrs = []
        dist.barrier(group=self.group)
        if self._own is not None:
            _cabi.check(lib.dfit_ipc_free(self._handle.ptr, self._own))
            self._own = None
