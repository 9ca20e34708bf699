"""Host-side helpers for the one-process-per-GPU launch (torch.distributed is plumbing only)."""
from __future__ import annotations

from typing import List, Tuple


def aggregate(ms_local: float, units_local: float) -> Tuple[float, float]:
    """(max over ranks of the device time, sum over ranks of the processed units)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms_local, units_local
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u = torch.tensor([units_local], dtype=torch.float64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def slab_bounds(n_leaf_planes: int, world: int) -> List[Tuple[int, int]]:
    """Slab decomposition along x in whole leaf planes: rank r owns [lo, hi)."""
    base, rem = divmod(n_leaf_planes, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def leaf_layer(pos, dx):
    """leaf layer (voxel >> 3) along x of world positions, with the reference's voxel rule ijk = floor(pos/dx + 0.5)
    in double (openvdb/math/Transform.h:111)"""
    import numpy as np
    inv = 1.0 / np.float64(np.float32(dx))
    return (np.floor(pos[:, 0].astype(np.float64) * inv + 0.5).astype(np.int64)) >> 3


def balanced_slabs(layer_counts, world: int, min_layers: int = 2) -> List[Tuple[int, int]]:
    """Slabs of whole leaf layers with (nearly) equal particle counts: cut the prefix sum of the per-layer histogram
    (first non-empty layer .. last non-empty layer) at multiples of total/world, keeping >= min_layers per rank
    (boundaries chosen by occupied-leaf count)."""
    import numpy as np
    c = np.asarray(layer_counts, dtype=np.int64)
    nz = np.nonzero(c)[0]
    first, last = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, len(c))
    if last - first < min_layers * world:
        raise ValueError(f"{last - first} occupied leaf layers cannot give {world} ranks {min_layers} layers each")
    pre = np.concatenate([[0], np.cumsum(c[first:last])])
    total = pre[-1]
    cuts = [first]
    for r in range(1, world):
        k = int(np.searchsorted(pre, total * r / world, side="left")) + first
        k = max(k, cuts[-1] + min_layers)
        k = min(k, last - min_layers * (world - r))
        cuts.append(k)
    cuts.append(last)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]
