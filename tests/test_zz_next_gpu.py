"""GPU parity of the first node beyond the substep chain: KillParticlesInSDF.

The CPU side of this node is pinned in tests/test_ref_pin_cpu.py (the reference's own node class == oracle == the drop-in's node).
This file compares the CUDA implementation with the oracle through the C ABI (first run on a B200: the round-1 driver run, all green;
the expected-failure markers of that first run are gone). The device code of these kernels is also executed on the CPU and matches the
oracle bit for bit (tests/test_next_kernels_emul_cpu.py).
"""
import numpy as np
import pytest

from tests import util
from zeno_b200 import scenes

pytestmark = pytest.mark.gpu



@pytest.mark.parametrize("keep", [True, False], ids=["KEEP", "DEL"])
def test_kill_particles_in_sdf_matches_oracle(gpu_lib, oracle_lib, keep):
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    N = 32
    pos, vel, dx = scenes.dam_break_points(N, seed=9, random_velocity=True)
    killer = scenes.sphere_sdf(centre=(3.3, 4.1, 2.7), radius=4.6, lo=(-8, -8, -8), hi=(16, 16, 16), bg=3.0)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.set_grid("KillerSDF", killer)
    n0 = gw.particles_info()[1]
    for w in (gw, ow):
        w.KillParticlesInSDF("KillerSDF", keep)
    a = scenes.canonical_particles(gw.get_particles())
    b = scenes.canonical_particles(ow.get_particles())
    assert 0 < b.shape[0] < n0
    assert a.shape == b.shape, f"{a.shape[0]} survivors on the GPU vs {b.shape[0]} in the oracle"
    assert np.array_equal(a, b), "surviving particles differ from the oracle"
    util.check_store_invariants(gw.get_particles())
    # the filtered store is a valid input of the next node
    for w in (gw, ow):
        w.FLIP_P2G(dx, 3)
    for name in ("Velocity", "LiquidSDF"):
        util.compare_grids(gw.get_grid(name), ow.get_grid(name), f"P2G after KillParticlesInSDF: {name}", tol=0.0, check_inactive=False)
    gw.close()


def test_particle_add_dv_matches_oracle(gpu_lib, oracle_lib):
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    pos, vel, dx = scenes.dam_break_points(32, seed=2, random_velocity=True)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.ParticleAddDV(0.013, -0.1633333, 1.0e-4)
    util.compare_particles(gw.get_particles(), ow.get_particles(), "ParticleAddDV vs oracle")
    gw.close()


def test_plain_g2p_advector_matches_oracle(gpu_lib, oracle_lib):
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    N, dt = 32, 0.01
    pos, vel, dx = scenes.dam_break_points(N, seed=6, random_velocity=True)
    vel *= np.float32(0.3)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIP_P2G(dx, 3)
        w.FieldAddVector(0.0, -9.8 * dt, 0.0)
        w.G2P_Advector(dt, dx, 3, 0.3)
    a = scenes.canonical_particles(gw.get_particles())
    b = scenes.canonical_particles(ow.get_particles())
    assert a.shape == b.shape
    assert (a[:, :3] == b[:, :3]).all(axis=1).mean() > 0.999
    assert np.array_equal(a, b), "plain G2P_Advector: quantised particle state differs from the oracle"
    gw.close()


def test_renormalize_sdf_matches_oracle(gpu_lib, oracle_lib):
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    pos, vel, dx = scenes.dam_break_points(32, seed=8, random_velocity=True)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIP_P2G(dx, 3)
        w.VDBRenormalizeSDF("LiquidSDF", 4, 0)
    util.compare_grids(gw.get_grid("LiquidSDF"), ow.get_grid("LiquidSDF"), "VDBRenormalizeSDF vs oracle", tol=0.0, check_inactive=False)
    with pytest.raises(abi.FlipB200Error):
        gw.VDBRenormalizeSDF("LiquidSDF", 1, 2)      # the tracker's dilate is not accelerated
    gw.close()


def test_erode_sdf_matches_oracle(gpu_lib, oracle_lib):
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    pos, vel, dx = scenes.dam_break_points(32, seed=8)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIP_P2G(dx, 3)
        w.VDBErodeSDF("LiquidSDF", 0.37 * dx)
    util.compare_grids(gw.get_grid("LiquidSDF"), ow.get_grid("LiquidSDF"), "VDBErodeSDF vs oracle", tol=0.0, check_inactive=False)
    gw.close()


def test_smooth_sdf_matches_oracle(gpu_lib, oracle_lib):
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    pos, vel, dx = scenes.dam_break_points(32, seed=8)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIP_P2G(dx, 3)
        w.VDBSmoothSDF("LiquidSDF", 2, 2)
    util.compare_grids(gw.get_grid("LiquidSDF"), ow.get_grid("LiquidSDF"), "VDBSmoothSDF vs oracle", tol=0.0, check_inactive=False)
    gw.close()


@pytest.mark.parametrize("seed", [6, 31])
def test_fluid_reseed_matches_oracle(gpu_lib, oracle_lib, seed):
    """FluidReseed (FF/nosys/FLIP_Reseed.cpp -> FLIP_vdb::reseed_fluid): the oracle's restatement is pinned leaf by leaf against the
    reference's own node class in the seeded build (tests/test_ref_pin_cpu.py); here the CUDA kernels against the oracle with the
    same per-leaf draw starts (include/flipb200.h): identical stores, and the topped-up store is a valid input of FLIP_P2G."""
    from oracle.pyoracle import OracleWorld
    from tests.test_ref_pin_cpu import _reseed_worlds
    from zeno_b200 import abi
    (gw, ow), dx = _reseed_worlds((abi.World, OracleWorld), seed=seed)
    n0 = ow.particles_info()[1]
    for w in (gw, ow):
        w.FluidReseed(1000 + seed)
    n1 = ow.particles_info()[1]
    assert n1 > 1.5 * n0, f"the scene must make the reseeder work: {n0} -> {n1}"
    a = scenes.canonical_particles(gw.get_particles())
    b = scenes.canonical_particles(ow.get_particles())
    assert a.shape == b.shape, f"{a.shape[0]} particles on the GPU vs {b.shape[0]} in the oracle"
    assert np.array_equal(a, b), "the reseeded store differs from the oracle"
    util.check_store_invariants(gw.get_particles())
    # a second pass with another seed (voxels that are still short of 8 try again), then the transfer
    for w in (gw, ow):
        w.FluidReseed(2000 + seed)
        w.FLIP_P2G(dx, 3)
    util.compare_particles(gw.get_particles(), ow.get_particles(), "second FluidReseed")
    for name in ("Velocity", "LiquidSDF"):
        util.compare_grids(gw.get_grid(name), ow.get_grid(name), f"P2G after FluidReseed: {name}", tol=0.0, check_inactive=False)
    gw.close()


def test_fluid_reseed_through_the_node_class(gpu_lib, oracle_lib):
    """The drop-in's FluidReseed NODE on real OpenVDB objects with the CUDA library behind it == the oracle."""
    from oracle import pyoracle
    if not pyoracle.plugin_gpu_available():
        pytest.skip("oracle/_ref/libflipplugin_gpu.so is not built")
    from oracle.pyoracle import OracleWorld, PluginGpuWorld
    from tests.test_ref_pin_cpu import _reseed_worlds
    (pw, ow), dx = _reseed_worlds((PluginGpuWorld, OracleWorld), seed=8)
    for w in (pw, ow):
        w.FluidReseed(4242)
    util.compare_particles(pw.get_particles(), ow.get_particles(), "FluidReseed node (GPU) vs oracle")
    pw.close()


@pytest.mark.parametrize("seed", [4, 12])
def test_particle_emitter_matches_oracle(gpu_lib, oracle_lib, seed):
    """ParticleEmitter (FF/nosys/ParticleEmitter.cpp -> FLIP_vdb::emit_liquid), constant-velocity branch: the oracle is pinned against
    the reference's node class (tests/test_ref_pin_cpu.py); here CUDA == oracle: same leaves created, identical stores, twice in a
    row (the second call tops up what the first left short), and the result feeds FLIP_P2G."""
    from oracle.pyoracle import OracleWorld
    from tests.test_ref_pin_cpu import _emitter_scene
    from zeno_b200 import abi
    pos, vel, dx, shape = _emitter_scene(seed=seed)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.set_grid("KillerSDF", shape)
    n0 = ow.particles_info()[1]
    for k in range(2):
        for w in (gw, ow):
            w.ParticleEmitter("KillerSDF", 0.5, 0.0, -0.75, seed=100 * seed + k)
        a = scenes.canonical_particles(gw.get_particles())
        b = scenes.canonical_particles(ow.get_particles())
        assert a.shape == b.shape, f"call {k}: {a.shape[0]} particles on the GPU vs {b.shape[0]} in the oracle"
        assert np.array_equal(a, b), f"call {k}: the store differs from the oracle"
    assert ow.particles_info()[1] > n0 + 1000
    util.check_store_invariants(gw.get_particles())
    for w in (gw, ow):
        w.FLIP_P2G(dx, 3)
    for name in ("Velocity", "LiquidSDF"):
        util.compare_grids(gw.get_grid(name), ow.get_grid(name), f"P2G after ParticleEmitter: {name}", tol=0.0, check_inactive=False)
    gw.close()


def test_particle_emitter_into_an_empty_world(gpu_lib, oracle_lib):
    from oracle.pyoracle import OracleWorld
    from tests.test_ref_pin_cpu import _emitter_scene
    from zeno_b200 import abi
    _, _, dx, shape = _emitter_scene()
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.set_grid("KillerSDF", shape)
        w.ParticleEmitter("KillerSDF", 0.0, 1.0, 0.0, seed=5)
    a = scenes.canonical_particles(gw.get_particles())
    b = scenes.canonical_particles(ow.get_particles())
    assert b.shape[0] > 1000 and a.shape == b.shape and np.array_equal(a, b)
    gw.close()


def test_flip_apply_boundary_matches_oracle(gpu_lib, oracle_lib):
    """FLIPApplyBoundary: CUDA == oracle (pinned against the reference's node class in tests/test_ref_pin_cpu.py), bit for bit, and the
    merged SDF is what the following nodes see (CutCellWeight after it agrees too)."""
    from oracle.pyoracle import OracleWorld
    from tests.test_ref_pin_cpu import _boundary_scene
    from zeno_b200 import abi
    pos, vel, dx, solid, sphere = _boundary_scene(seed=5)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.set_grid("SolidSDF", solid)
        w.set_grid("KillerSDF", sphere)
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIP_P2G(dx, 3)
        w.FLIPApplyBoundary("KillerSDF")
    a, b = gw.get_grid("SolidSDF"), ow.get_grid("SolidSDF")
    assert b["origins"].shape[0] > solid["origins"].shape[0]
    util.compare_grids(a, b, "FLIPApplyBoundary: SolidSDF", tol=0.0, check_inactive=True)
    for w in (gw, ow):
        w.CutCellWeight()
    util.compare_grids(gw.get_grid("CellFWeight"), ow.get_grid("CellFWeight"), "CutCellWeight after FLIPApplyBoundary", tol=0.0, check_inactive=False)
    gw.close()


def test_surface_tension_matches_oracle(gpu_lib, oracle_lib):
    """The tension terms of AssembleSolvePPE / SubtractPressureGradient (pinned against the reference's node classes in
    tests/test_ref_pin_cpu.py): CUDA right-hand side == oracle bit for bit, pressure within the solver tolerance, same iterations
    (<= 1.1x), projected velocity from one pressure field bit for bit."""
    import math
    from oracle.pyoracle import OracleWorld
    from tests.test_ref_pin_cpu import _tension_worlds
    from zeno_b200 import abi
    (gw, ow), dx, dt = _tension_worlds((abi.World, OracleWorld))
    rg, ro = gw.AssembleSolvePPE(dt, dx), ow.AssembleSolvePPE(dt, dx)
    util.compare_grids(gw.get_grid("Divergence"), ow.get_grid("Divergence"), "tension RHS", tol=0.0, check_inactive=False)
    assert rg["status"] == 0 and ro["status"] == 0 and rg["iterations"] <= math.ceil(1.1 * ro["iterations"]), (rg, ro)
    util.compare_grids(gw.get_grid("Pressure"), ow.get_grid("Pressure"), "tension pressure", tol=util.REF_TOL["ppe"], check_inactive=False)
    gw.set_grid("Pressure", ow.get_grid("Pressure"))
    for w in (gw, ow):
        w.SubtractPressureGradient(dt, dx, 3)
    util.compare_grids(gw.get_grid("Velocity"), ow.get_grid("Velocity"), "tension gradient", tol=0.0, check_inactive=False)
    # switching the terms off again restores the plain right-hand side
    gw.set_surface_tension(1000.0, 0.0); ow.set_surface_tension(1000.0, 0.0)
    gw.AssembleSolvePPE(dt, dx); ow.AssembleSolvePPE(dt, dx)
    util.compare_grids(gw.get_grid("Divergence"), ow.get_grid("Divergence"), "plain RHS after tension", tol=0.0, check_inactive=False)
    gw.close()


def test_points_to_primitive_matches_oracle(gpu_lib, oracle_lib):
    """VDBPointsToPrimitive (projects/zenvdb/GetVDBPoints.cpp): world positions and velocities straight from the device
    store == the oracle (pinned against the node's arithmetic in tests/test_plugin_cpu.py), bit for bit, same order."""
    from oracle.pyoracle import OracleWorld
    from zeno_b200 import abi
    pos, vel, dx = scenes.dam_break_points(64, seed=11, random_velocity=True)
    gw, ow = abi.World(dx), OracleWorld(dx)
    for w in (gw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
    gp, gv = gw.VDBPointsToPrimitive()
    op, ov = ow.VDBPointsToPrimitive()
    assert gp.shape == op.shape and np.array_equal(gp, op), "world positions differ"
    assert np.array_equal(gv, ov), "velocities differ"
    assert np.abs(np.sort(gp, axis=0) - np.sort(pos, axis=0)).max() < dx, "positions must be the input points up to the codec"
    gw.close()
