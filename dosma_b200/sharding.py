"""Multi-GPU plumbing: one process per GPU, z-slab / voxel-range partition, ONE all-gather at the end.

The fit is embarrassingly parallel over voxels (the reference's only parallelism is a process pool
over voxels, dosma/core/fitting.py:861-868), so ranks never exchange data during the fit; the only
collective is the final reassembly of the parameter map.  Ragged slabs are
padded to the largest slab so that the reassembly stays a single `all_gather_into_tensor`.
"""
import numpy as np

__all__ = ["slab_bounds", "voxel_ranges", "masked_spans", "gather_maps", "fit_sharded", "PeerMaps"]


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


def masked_spans(mask, world_size, align=4):
    """Voxel spans [lo_r, hi_r) of the ranks for a masked fit of ONE volume by all ranks (the kernels' split mode,
    `dfit_gather_desc.split_list`): consecutive, covering the whole volume, cut so that every span holds (nearly) the
    same number of masked voxels -- the balanced partition of the mask-compacted voxel list, expressed in voxel
    indices so that each rank knows which samples it needs.  Cuts fall on multiples of `align` voxels.
    Returns world_size + 1 boundaries."""
    m = np.asarray(mask).reshape(-1) != 0
    n = m.shape[0]
    idx = np.flatnonzero(m)
    b = [0]
    for r in range(1, int(world_size)):
        k = idx.shape[0] * r // int(world_size)
        cut = int(idx[k]) // align * align if idx.shape[0] else n * r // int(world_size) // align * align
        b.append(max(cut, b[-1]))
    b.append(n)
    return np.asarray(b, dtype=np.int64)


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

    Every rank owns one full map `[rows, ncols]` (fp32; `ncols` = parameters carried + r2) and can store into all of
    them.  Two ways to get there, tried in this order:

    * torch symmetric memory (`torch.distributed._symmetric_memory`: plumbing, like `torch.distributed` itself): the
      buffers are exchanged by the runtime and, where the fabric offers it, bound to an NVLS multicast object -- one
      `multimem.st` per row leaves the GPU and the NVSwitch replicates it to all ranks;
    * CUDA IPC through the C-ABI (`dfit_ipc_alloc` / `dfit_ipc_open`, lazy peer access over NVLink): one store per rank.

    `dfit_set_gather_ex` then hands the pointers to the kernels, after which `fit_device` stores each voxel's row into
    all maps while the fit is running.  `synchronize()` (stream drain + barrier) makes the local map complete.

    rows_per_rank, ncols: the dense layout (rank r's voxels are rows [r rows_per_rank, (r + 1) rows_per_rank)).
    total_rows / row0: any other layout (e.g. the split mode for masked fits of one volume: every rank addresses the
    whole volume, row0 = 0, and fits the masked voxels of `fit_span` from samples that start at voxel `y_voxel0`).
    param_mask: parameters carried per row (bit i = parameter i, 0 = all); ncols must equal their number + 1.
    multicast: "off" (default: one peer store per rank), "auto" (NVLS multicast stores where available) or "require".
        A multicast store also delivers the rank's own copy through the switch -- N instead of N - 1 maps arrive at every
        GPU -- and measured slower than peer stores on this pool; it saves SM store instructions.
    copy_engine: True = the third transport, for dense layouts: the kernel stores its rows into the rank's OWN map only
        and `fit_pipelined` cuts the fit into chunks whose rows the copy engines push to the peers' maps (peer-to-peer
        cudaMemcpyAsync over NVLink on a side stream) while the SMs fit the next chunk -- the copy engines move full-size
        NVLink packets, which the 16-byte-per-lane stores of the fused epilogue do not reach.
    """

    def __init__(self, rows_per_rank, ncols, device, group=None, *, total_rows=None, row0=None, param_mask=0,
                 split_list=False, fit_span=(0, 0), y_voxel0=0, multicast="off", copy_engine=False):
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
        rows = int(total_rows) if total_rows is not None else self.world * self.rows_per_rank
        row0 = int(row0) if row0 is not None else self.rank * self.rows_per_rank
        shape = (rows, int(ncols))
        self._own, self._peer_ptrs, self._symm = None, [], None
        self.transport = None
        self.copy_engine = bool(copy_engine)
        if self.copy_engine and (split_list or multicast != "off"):
            raise ValueError("copy_engine is for dense layouts and excludes multicast")
        self._base_row0 = row0
        mc_ptr = 0
        ptrs = None
        if multicast not in ("off", "auto", "require"):
            raise ValueError("multicast must be 'off', 'auto' or 'require'")
        if multicast != "off":
            ptrs, mc_ptr = self._try_symmetric(shape, device, group)
        if ptrs is None:
            nbytes = shape[0] * shape[1] * 4
            own = ctypes.c_void_p()
            hbuf = ctypes.create_string_buffer(64)
            _cabi.check(lib.dfit_ipc_alloc(self._handle.ptr, nbytes, ctypes.byref(own), hbuf))
            self._own = own
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(hbuf.raw), group=group)
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(own.value)
                else:
                    p = ctypes.c_void_p()
                    _cabi.check(lib.dfit_ipc_open(self._handle.ptr, handles[r], ctypes.byref(p)))
                    self._peer_ptrs.append(p)
                    ptrs.append(p.value)
            self.transport = "cuda-ipc peer stores"
        if multicast == "require" and not mc_ptr:
            raise RuntimeError("NVLS multicast was required but is not available")
        self.multicast_ptr = int(mc_ptr or 0)
        with torch.cuda.device(device):
            self.maps = [torch.as_tensor(_DevArray(p, shape), device=device) for p in ptrs]
        self.local = self.maps[self.rank]
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        gd = _cabi.DfitGatherDesc()
        gd.struct_size = ctypes.sizeof(gd)
        gd.world, gd.rank = self.world, self.rank
        gd.maps = ctypes.cast(self._ptrs, ctypes.c_void_p)
        if self.copy_engine:  # the kernel sees a world of one: its own map
            self._own_ptr = (ctypes.c_void_p * 1)(ptrs[self.rank])
            gd.world, gd.rank = 1, 0
            gd.maps = ctypes.cast(self._own_ptr, ctypes.c_void_p)
            self.transport = "copy-engine peer copies, pipelined with the fit"
            self._copy_stream = torch.cuda.Stream(device=device)
            self._events = {}
        gd.multicast = self.multicast_ptr or None
        gd.rows, gd.row0 = rows, row0
        gd.param_mask = int(param_mask)
        gd.split_list = int(bool(split_list))
        gd.fit_lo, gd.fit_hi = int(fit_span[0]), int(fit_span[1])
        gd.y_voxel0 = int(y_voxel0)
        self._desc = gd
        _cabi.check(lib.dfit_set_gather_ex(self._handle.ptr, ctypes.byref(gd)))
        dist.barrier(group=group)

    def _try_symmetric(self, shape, device, group):
        """Allocate the maps as torch symmetric memory; returns (pointers, multicast pointer) or (None, 0).  All
        ranks take the same branch (the outcome is agreed on with an all-reduce)."""
        import torch
        import torch.distributed as dist

        ok, ptrs, mc = 1, None, 0
        try:
            import torch.distributed._symmetric_memory as symm

            t = symm.empty(shape, dtype=torch.float32, device=device)
            hdl = symm.rendezvous(t, group=group if group is not None else dist.group.WORLD)
            t.zero_()
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            self._symm = (t, hdl)
        except Exception:  # not built / not permitted / no fabric support: the IPC path does the same job
            ok = 0
        flag = torch.tensor([ok, 1 if mc else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        torch.cuda.synchronize(device)
        if int(flag[0]) == 0:
            self._symm = None
            return None, 0
        self.transport = "symmetric memory, NVLS multicast stores" if int(flag[1]) else "symmetric memory, peer stores"
        return ptrs, (mc if int(flag[1]) else 0)

    def reconfigure(self, **kw):
        """Change row0 / param_mask / split_list / y_voxel0 of the gather without re-allocating the maps."""
        import ctypes

        from . import _cabi

        for k, v in kw.items():
            setattr(self._desc, k, int(v))
        _cabi.check(_cabi.load().dfit_set_gather_ex(self._handle.ptr, ctypes.byref(self._desc)))

    def fit_pipelined(self, fit_chunk, n_rows, chunks=8, align=128):
        """copy_engine transport: `fit_chunk(lo, hi)` launches the fit of this rank's voxels [lo, hi) on the current stream
        (the gather is already pointed at their rows); the rows of every chunk then travel to the peers' maps on the copy
        stream while the next chunk is fitted.  Returns immediately (asynchronous); `synchronize()` completes the maps."""
        import torch

        assert self.copy_engine
        cur = torch.cuda.current_stream(self.device)
        step = max(align, (int(n_rows) + chunks - 1) // chunks // align * align)
        lo, c = 0, 0
        while lo < n_rows:
            hi = min(int(n_rows), lo + step)
            if hi + align > n_rows:
                hi = int(n_rows)
            fitted, copied = self._events.setdefault(c, (torch.cuda.Event(), torch.cuda.Event()))
            cur.wait_event(copied)  # the previous fit's copies of these rows have left before they are overwritten
            self.reconfigure(row0=self._base_row0 + lo)
            fit_chunk(lo, hi)
            fitted.record(cur)
            self._copy_stream.wait_event(fitted)
            with torch.cuda.stream(self._copy_stream):
                r0, r1 = self._base_row0 + lo, self._base_row0 + hi
                for k in range(1, self.world):  # round the ring, starting behind the own rank
                    r = (self.rank + k) % self.world
                    self.maps[r][r0:r1].copy_(self.local[r0:r1], non_blocking=True)
                copied.record(self._copy_stream)
            lo, c = hi, c + 1
        self.last_chunks = c
        cur.wait_event(copied)  # the fit is complete on the current stream when its last rows have left

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
        _cabi.check(lib.dfit_set_gather_ex(self._handle.ptr, None))
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
        self._symm = None
