"""GPU parity of the node variants the substep benchmark never takes ("untested GPU configurations"):
every RK order of the advection, a separate ViscousVelocity field, a moving solid (SolidVelocity), and the pure-multigrid
fallback after a failed PCG (FF/simd_vdb_poisson_uaamg.cpp:2405-2444, FF/FLIP_vdb.cpp:3089-3097). CUDA through the C ABI
against the CPU oracle on the same seeded inputs; bars as in tests/test_parity_gpu.py."""
import numpy as np
import pytest

from tests import util
from zeno_b200 import scenes

pytestmark = pytest.mark.gpu

DT = 0.008
G = (0.0, -9.8, 0.0)


def _pair(N=48, seed=4, vel_scale=0.6):
    """A GPU world and an oracle world after one projection: a divergence-free velocity to advect in."""
    from oracle.pyoracle import OracleWorld
    from zeno_b200.abi import World
    pos, vel, dx = scenes.dam_break_points(N, seed=seed, random_velocity=True)
    vel = vel * vel_scale
    solid = scenes.box_solid_sdf(N, dx)
    gw, ow = World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.set_grid("SolidSDF", solid)
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIP_P2G(dx, 3)
        w.CutCellWeight()
        w.PushOutLiquidSDF(dx)
        w.FieldAddVector(G[0] * DT, G[1] * DT, G[2] * DT)
    ow.AssembleSolvePPE(DT, dx)
    ow.SubtractPressureGradient(DT, dx, 3)
    util.sync_state(gw, ow)
    return gw, ow, dx


def _compare_advected(gw, ow, what):
    n = ow.particles_info()[1]
    a = scenes.canonical_particles(gw.get_particles())
    b = scenes.canonical_particles(ow.get_particles())
    assert a.shape == b.shape, f"{what}: {a.shape[0]} particles on the GPU vs {b.shape[0]} in the oracle"
    assert np.array_equal(a, b), f"{what}: quantised particle state differs from the oracle in {(a != b).any(axis=1).sum()} of {n} particles"
    assert gw.dropped() == ow.dropped()
    util.check_store_invariants(gw.get_particles())


@pytest.mark.parametrize("rk", [1, 2, 3, 4])
@pytest.mark.parametrize("same_field", [True, False], ids=["viscous=velocity", "viscous=separate"])
def test_g2p_rk_orders_and_viscous_field(gpu_lib, oracle_lib, rk, same_field):
    gw, ow, dx = _pair()
    if not same_field:
        # a ViscousVelocity that differs from Velocity: the carried velocity is sampled from it (FF/FLIP_vdb.cpp:596-606)
        v = ow.get_grid("Velocity")
        v = dict(v)
        v["values"] = (v["values"] * np.float32(0.75)).astype(np.float32)
        for w in (gw, ow):
            w.set_grid("ViscousVelocity", v)
    before = scenes.canonical_particles(ow.get_particles())
    for w in (gw, ow):
        w.capture_precodec(True)
        # surface_size 0: every particle with a negative liquid SDF takes the RK branch (the P2G SDF never gets deeper than ~ -0.8 dx,
        # so with the default band of 4 voxels every particle would take the Euler step whatever RK_ORDER says)
        w.G2PAdvectorSheetty(DT, dx, 0, rk, 0.03, 0.05, same_field)
    n = before.shape[0]
    pos, vel, alive = gw.get_precodec(n)
    rpos, rvel, ralive = ow.get_precodec(n)
    assert np.array_equal(alive, ralive)
    m = alive.astype(bool)
    assert util.rel_l2(pos[m], rpos[m]) <= 1e-5 and util.rel_l2(vel[m], rvel[m]) <= 1e-5   # north star: 1e-5 relative L2 on the pre-codec state
    _compare_advected(gw, ow, f"G2PAdvectorSheetty RK{rk} same_field={same_field}")
    after = scenes.canonical_particles(ow.get_particles())
    assert after.shape != before.shape or not np.array_equal(after, before)
    gw.close()


def test_g2p_rk_orders_differ(gpu_lib):
    """The RK branch taken is really the one asked for: the four orders give four different particle states."""
    from zeno_b200.abi import World
    pos, vel, dx = scenes.dam_break_points(48, seed=4, random_velocity=True)
    outs = []
    for rk in (1, 2, 3, 4):
        w = World(dx)
        w.set_grid("SolidSDF", scenes.box_solid_sdf(48, dx))
        w.PrimToVDBPointDataGrid(pos, vel * 0.6)
        w.FLIP_P2G(dx, 3)
        w.G2PAdvectorSheetty(DT, dx, 0, rk, 0.03, 0.05, True)   # surface_size 0: the RK branch for every particle inside the liquid
        outs.append(scenes.canonical_particles(w.get_particles()))
        w.close()
    for i in range(4):
        for j in range(i + 1, 4):
            assert outs[i].shape != outs[j].shape or not np.array_equal(outs[i], outs[j]), f"RK{i + 1} and RK{j + 1} give the same result"


def test_g2p_moving_solid_pushes_particles(gpu_lib, oracle_lib):
    """Particles driven into a wall that moves (SolidVelocity set): the push-out with the solid's normal velocity
    (FF/FLIP_vdb.cpp:687-703, :3270-3367) must execute and agree with the oracle; a second run without SolidVelocity differs."""
    from oracle.pyoracle import OracleWorld
    from zeno_b200.abi import World
    N = 48
    pos, vel, dx = scenes.dam_break_points(N, seed=6, random_velocity=True)
    vel = vel * 0.3
    vel[:, 0] -= 4.0            # towards the x = 0 wall, the block sits a few voxels from it
    vel[:, 1] -= 3.0
    solid = scenes.box_solid_sdf(N, dx)
    results = {}
    for with_vel in (True, False):
        gw, ow = World(dx), OracleWorld(dx)
        for w in (gw, ow):
            w.set_grid("SolidSDF", solid)
            w.PrimToVDBPointDataGrid(pos, vel)
            w.FLIP_P2G(dx, 3)
        if with_vel:
            sv = dict(ow.get_grid("Velocity"))
            sv["values"] = np.zeros_like(sv["values"])
            sv["values"][:, 0] = 0.8     # the wall moves in +x
            sv["values"][:, 1] = 0.4
            sv["masks"] = np.full_like(sv["masks"], np.uint64(0xffffffffffffffff))
            for w in (gw, ow):
                w.set_grid("SolidVelocity", sv)
        for w in (gw, ow):
            w.G2PAdvectorSheetty(0.02, dx, 4, 3, 0.03, 0.05, True)
        _compare_advected(gw, ow, f"moving solid, SolidVelocity={with_vel}")
        results[with_vel] = scenes.canonical_particles(ow.get_particles())
        gw.close()
    a, b = results[True], results[False]
    assert a.shape == b.shape
    changed = int((a != b).any(axis=1).sum())
    assert changed > 0, "no particle was pushed out of the moving solid: the SolidVelocity branch did not execute"
    print(f"particles whose state depends on SolidVelocity: {changed} of {a.shape[0]}")


def test_pure_multigrid_fallback(gpu_lib, oracle_lib):
    """PCG limited to one iteration fails (status 1); the node then warm-starts from the previous pressure and runs the pure
    mu-cycle iteration (w = 1, prolongation x 0.5, 10 n coarsest sweeps). Same fallback in the oracle: same status, pressure close."""
    gw, ow, dx = _pair(N=72, seed=7)   # 5832 DOFs: two multigrid levels (see tests/test_ref_pin_cpu.py for the one-level quirk of the reference)
    res = []
    for w in (gw, ow):
        w.FieldAddVector(0.0, -0.05, 0.0)    # a new right-hand side on top of the projected field
        res.append(w.AssembleSolvePPE(DT, dx, rel_tol=1e-4, max_iter=1))
    assert res[0]["status"] == 1 and res[1]["status"] == 1, res
    assert res[0]["iterations"] == res[1]["iterations"]
    # (the reported residual is relative to the residual of the ZERO guess, the fallback's tolerance to that of the warm start)
    assert res[0]["rel_residual"] <= 1e-3 and abs(res[0]["rel_residual"] - res[1]["rel_residual"]) <= 0.05 * res[1]["rel_residual"], res
    util.compare_grids(gw.get_grid("Divergence"), ow.get_grid("Divergence"), "fallback: right-hand side", tol=0.0)
    util.compare_grids(gw.get_grid("Pressure"), ow.get_grid("Pressure"), "fallback: pressure", tol=2e-3, check_inactive=False)
    gw.close()
