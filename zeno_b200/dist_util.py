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
