"""The two P2G transfer kernels are the same function: p2g_xrow_kernel (thread per x-row of targets, the default) against
p2g_gather_kernel (thread per target voxel, round 1; FLIPB200_P2G_OLD=1), bit for bit, on ordinary, dense (row batches)
and crowded (leaf handed back to the old kernel) stores -- and against the oracle where the oracle is quick."""
import os

import numpy as np
import pytest

from tests import util
from zeno_b200 import scenes

pytestmark = pytest.mark.gpu


def _p2g(w, dx, old):
    saved = os.environ.get("FLIPB200_P2G_OLD")
    try:
        if old:
            os.environ["FLIPB200_P2G_OLD"] = "1"
        else:
            os.environ.pop("FLIPB200_P2G_OLD", None)
        w.FLIP_P2G(dx, 3)
    finally:
        if saved is None:
            os.environ.pop("FLIPB200_P2G_OLD", None)
        else:
            os.environ["FLIPB200_P2G_OLD"] = saved
    return {name: w.get_grid(name) for name in ("Velocity", "PostAdvVelocity", "LiquidSDF")}


def _same(a, b, what):
    for name in a:
        util.compare_grids(a[name], b[name], f"{what} {name}", tol=0.0)


@pytest.mark.parametrize("ppc", [8, 27])
def test_p2g_paths_identical(gpu_lib, ppc):
    """8 ppc: one batch per plane; 27 ppc (the reference's per-voxel cap): every plane in two or three row batches"""
    from zeno_b200 import abi
    N = 64 if ppc == 8 else 32
    gw, _, dx, pos, vel = util.make_worlds(N, seed=3, ppc=ppc, random_velocity=True, gpu_world_cls=abi.World)
    new = _p2g(gw, dx, old=False)
    old = _p2g(gw, dx, old=True)
    _same(new, old, f"P2G x-row vs gather, ppc={ppc}")
    # after a substep the per-voxel counts are ragged
    gw.CutCellWeight(); gw.PushOutLiquidSDF(dx); gw.FieldAddVector(0.0, -0.098, 0.0)
    gw.AssembleSolvePPE(0.01, dx); gw.SubtractPressureGradient(0.01, dx, 3)
    gw.G2PAdvectorSheetty(0.01, dx, 4, 3, 0.03, 0.05, True)
    new = _p2g(gw, dx, old=False)
    old = _p2g(gw, dx, old=True)
    _same(new, old, f"P2G x-row vs gather after a substep, ppc={ppc}")
    gw.close()


def test_p2g_crowded_rows_fall_back(gpu_lib, oracle_lib):
    """A few voxels with hundreds of particles: their cell rows do not fit the x-row kernel's staging buffer, the leaf is
    handed back to the gather kernel; everything else stays on the x-row kernel. Same grids as the oracle, bit for bit."""
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    rng = np.random.default_rng(11)
    N = 32
    dx = 1.0 / N
    cells = rng.integers(6, 22, size=(400, 3))
    counts = rng.integers(0, 12, size=400)
    crowded = rng.choice(400, size=6, replace=False)
    counts[crowded] = rng.integers(300, 700, size=6)
    idx = np.repeat(cells, counts, axis=0).astype(np.float64) + rng.uniform(-0.49, 0.49, size=(counts.sum(), 3))
    pos = (idx * dx).astype(np.float32)
    vel = rng.uniform(-1, 1, size=pos.shape).astype(np.float32)
    res = []
    for cls in (abi.World, OracleWorld):
        w = cls(dx)
        w.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
        w.PrimToVDBPointDataGrid(pos, vel)
        res.append(w)
    gw, ow = res
    new = _p2g(gw, dx, old=False)
    old = _p2g(gw, dx, old=True)
    _same(new, old, "crowded: x-row + fallback vs gather")
    ow.FLIP_P2G(dx, 3)
    for name in new:
        util.compare_grids(new[name], ow.get_grid(name), f"crowded P2G {name} vs oracle", tol=0.0, check_inactive=False)
    gw.close()
