"""Shared helpers of the parity tests: scene set-up on both worlds and grid/particle comparison."""
from __future__ import annotations

import numpy as np

from zeno_b200 import scenes


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return float(d / n) if n > 0 else float(d)


def compare_grids(g_gpu, g_ref, what: str, tol: float = 0.0, check_inactive: bool = True):
    """Active masks bit-exact; active values bit-exact (tol == 0) or relative-L2 <= tol.
    Returns the relative L2 error of the active values."""
    a = scenes.canonical_grid(g_gpu)
    b = scenes.canonical_grid(g_ref)
    assert a["origins"].shape == b["origins"].shape, f"{what}: leaf count {a['origins'].shape[0]} vs {b['origins'].shape[0]}"
    assert np.array_equal(a["origins"], b["origins"]), f"{what}: leaf origins differ"
    ham = int(np.unpackbits((a["masks"] ^ b["masks"]).view(np.uint8)).sum())
    assert ham == 0, f"{what}: active masks differ in {ham} voxels"
    mb = scenes.mask_bits(a["masks"])  # [n,512]
    nch = a["values"].shape[1]
    err = 0.0
    for c in range(nch):
        va, vb = a["values"][:, c][mb], b["values"][:, c][mb]
        if tol == 0.0:
            bad = int((va.view(np.uint32) != vb.view(np.uint32)).sum())
            # +0.0 and -0.0 compare equal numerically; only a numeric difference is an error
            if bad:
                bad = int((va != vb).sum())
            assert bad == 0, f"{what}[{c}]: {bad} active values differ (max abs {np.abs(va - vb).max()})"
        else:
            e = rel_l2(va, vb)
            err = max(err, e)
            assert e <= tol, f"{what}[{c}]: relative L2 {e:.3e} > {tol:.1e}"
        if check_inactive:
            ia, ib = a["values"][:, c][~mb], b["values"][:, c][~mb]
            if tol == 0.0:
                assert int((ia != ib).sum()) == 0, f"{what}[{c}]: inactive voxel values differ"
    return err


def compare_particles(p_gpu, p_ref, what: str = "particles"):
    a = scenes.canonical_particles(p_gpu)
    b = scenes.canonical_particles(p_ref)
    assert a.shape == b.shape, f"{what}: particle count {a.shape[0]} vs {b.shape[0]}"
    assert np.array_equal(a[:, :3], b[:, :3]), f"{what}: particle-to-voxel binning differs"
    assert np.array_equal(a, b), f"{what}: quantised particle state differs"


def check_store_invariants(p):
    """voxel_end is a per-leaf cumulative count consistent with the attribute arrays."""
    ve = p["voxel_end"].astype(np.int64)
    assert (np.diff(ve, axis=1) >= 0).all()
    assert ve[:, -1].sum() == p["P"].shape[0] == p["v"].shape[0]
    o = p["origins"]
    assert ((o % 8) == 0).all()
    assert len({tuple(x) for x in o.tolist()}) == o.shape[0]


def make_worlds(N: int, seed: int = 1, ppc: int = 8, random_velocity: bool = False, solid: bool = True,
                gpu_world_cls=None, oracle_world_cls=None, side=None):
    pos, vel, dx = scenes.dam_break_points(N, seed=seed, ppc=ppc, random_velocity=random_velocity, side=side)
    worlds = []
    for cls in (gpu_world_cls, oracle_world_cls):
        if cls is None:
            worlds.append(None)
            continue
        w = cls(dx)
        if solid:
            w.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
        w.PrimToVDBPointDataGrid(pos, vel)
        worlds.append(w)
    return worlds[0], worlds[1], dx, pos, vel


def sync_state(dst, src, grids=("Velocity", "PostAdvVelocity", "LiquidSDF", "CellFWeight", "Pressure", "Divergence")):
    """one-step-synchronised parity: start the next stage from the oracle's state"""
    for name in grids:
        dst.set_grid(name, src.get_grid(name))
    dst.set_particles(src.get_particles())
