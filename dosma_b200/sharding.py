"""Multi-GPU plumbing: one process per GPU, z-slab / voxel-range partition, ONE all-gather at the end.

The fit is embarrassingly parallel over voxels (the reference's only parallelism is a process pool
over voxels, dosma/core/fitting.py:861-868), so ranks never exchange data during the fit; the only
collective is the final reassembly of the parameter map.  Ragged slabs are
padded to the largest slab so that the reassembly stays a single `all_gather_into_tensor`.
"""
import numpy as np

__all__ = ["slab_bounds", "voxel_ranges", "gather_maps", "fit_sharded"]


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
