"""The multigrid preconditioner has several execution paths on the GPU (one kernel per pass, the round-1 one-launch
cycle kernel, the thread-block-cluster bottom + tiled sweeps). They all execute the reference's op sequence
(FF/simd_vdb_poisson_uaamg.cpp:1993-2126, :1109-1150) voxel by voxel, so they must agree BIT FOR BIT: same PCG residual
history, same pressure. Run on one device, path selected through the environment (read at every solve)."""
import os

import numpy as np
import pytest

from tests import util
from zeno_b200 import scenes

pytestmark = pytest.mark.gpu

DT = 0.008


def _world(N, seed=2):
    from zeno_b200.abi import World
    pos, vel, dx = scenes.dam_break_points(N, seed=seed, random_velocity=True)
    vel = vel * 0.3
    w = World(dx)
    w.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
    w.PrimToVDBPointDataGrid(pos, vel)
    w.FLIP_P2G(dx, 3)
    w.CutCellWeight()
    w.PushOutLiquidSDF(dx)
    w.FieldAddVector(0.0, -9.8 * DT, 0.0)
    return w, dx


def _solve(w, dx, env, tol=None):
    saved = {k: os.environ.get(k) for k in ("FLIPB200_MG_PATH", "FLIPB200_CLUSTER_FIRST", "FLIPB200_BRICK_MIN", "FLIPB200_CG_COMPAT",
                                            "FLIPB200_NO_CYCLE_KERNEL", "FLIPB200_MG_L1")}
    try:
        for k in saved:
            os.environ.pop(k, None)
        os.environ.update(env)
        vel = w.get_grid("Velocity")
        res = w.AssembleSolvePPE(DT, dx, rel_tol=tol)
        w.set_grid("Velocity", vel)
        info = w.solver_info()
        res["levels"] = info["levels"]
        return res, info["history"].copy(), w.get_grid("Pressure")
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("N,first", [(64, None), (64, 1), (128, None), (128, 2), (192, None)])
def test_paths_agree_bitwise(gpu_lib, N, first):
    w, dx = _world(N)
    # the first solve of a world sees the velocity as FieldAddVector left it, the later ones a download / upload copy of it
    # (inactive voxels read back as background): solve once so that both paths start from the same grids
    _solve(w, dx, {"FLIPB200_MG_PATH": "cycle"}, tol=1e-6)
    ref = _solve(w, dx, {"FLIPB200_MG_PATH": "cycle"}, tol=1e-6)
    env = {"FLIPB200_MG_PATH": "tiles", "FLIPB200_CG_COMPAT": "1"}   # the coarsest CG with the old path's summation order
    if first is not None:
        env["FLIPB200_CLUSTER_FIRST"] = str(first)
    new = _solve(w, dx, env, tol=1e-6)
    assert ref[0]["status"] == 0 and new[0]["status"] == 0, (ref[0], new[0])
    assert ref[0]["levels"] == new[0]["levels"]
    assert np.array_equal(ref[1], new[1]), (ref[1], new[1])
    util.compare_grids(new[2], ref[2], f"pressure, tiles vs cycle path, N={N}", tol=0.0)
    w.close()


@pytest.mark.parametrize("N", [64, 128])
def test_default_coarsest_cg_is_equivalent(gpu_lib, N):
    """The default coarsest-level CG sums in a different (fixed) order: same iteration count, pressure within 1e-5."""
    w, dx = _world(N)
    _solve(w, dx, {"FLIPB200_MG_PATH": "cycle"}, tol=1e-6)
    ref = _solve(w, dx, {"FLIPB200_MG_PATH": "tiles", "FLIPB200_CG_COMPAT": "1"}, tol=1e-6)
    new = _solve(w, dx, {"FLIPB200_MG_PATH": "tiles"}, tol=1e-6)
    again = _solve(w, dx, {"FLIPB200_MG_PATH": "tiles"}, tol=1e-6)
    assert new[0]["status"] == 0 and new[0]["iterations"] == ref[0]["iterations"], (ref[0], new[0])
    assert np.array_equal(new[1], again[1]), "the solve must be deterministic"
    util.compare_grids(new[2], ref[2], f"pressure, default CG vs compat CG, N={N}", tol=1e-5)
    util.compare_grids(new[2], again[2], f"pressure, two runs, N={N}", tol=0.0)
    w.close()


@pytest.mark.parametrize("N", [64, 128, 192])
def test_l1_cached_passes_agree_bitwise(gpu_lib, N):
    """The one-launch kernel reads the iterate through L1 between grid barriers (M_GRID_L1); FLIPB200_MG_L1=0 selects the
    L2-coherent loads it replaced. Any stale line would show up as a different residual history."""
    w, dx = _world(N)
    _solve(w, dx, {"FLIPB200_MG_L1": "0"}, tol=1e-6)
    for _ in range(2):
        ref = _solve(w, dx, {"FLIPB200_MG_L1": "0"}, tol=1e-6)
        new = _solve(w, dx, {"FLIPB200_MG_L1": "1"}, tol=1e-6)
        assert ref[0]["status"] == 0 and new[0]["status"] == 0, (ref[0], new[0])
        assert np.array_equal(ref[1], new[1]), (ref[1], new[1])
        util.compare_grids(new[2], ref[2], f"pressure, L1 vs L2 loads, N={N}", tol=0.0)
    w.close()


@pytest.mark.parametrize("N", [64, 128])
def test_one_kernel_per_pass_is_equivalent(gpu_lib, N):
    """The plain path (one kernel per colour pass, its own coarsest solve and reduction order): same iteration count,
    pressure within 1e-5."""
    w, dx = _world(N)
    _solve(w, dx, {"FLIPB200_MG_PATH": "cycle"}, tol=1e-6)
    ref = _solve(w, dx, {"FLIPB200_NO_CYCLE_KERNEL": "1"}, tol=1e-6)
    new = _solve(w, dx, {}, tol=1e-6)
    assert ref[0]["status"] == 0 and new[0]["status"] == 0, (ref[0], new[0])
    assert ref[0]["iterations"] == new[0]["iterations"], (ref[0], new[0])
    util.compare_grids(new[2], ref[2], f"pressure, one-launch vs one kernel per pass, N={N}", tol=1e-5)
    w.close()
