"""The Zeno-side drop-in END TO END on the GPU: the node classes of zeno_b200/plugin/flipb200_nodes.cpp (reference node names,
sockets and params), instantiated through the stand-in of the Zeno node runtime and wired like the packaged graph, operate on
REAL OpenVDB objects and call libflipb200.so (oracle/_ref/libflipplugin_gpu.so). One substep chain + the nodes beyond it,
stage by stage against the CPU oracle: bit-exact masks and stencil values, bit-identical particle state, pressure within the
solver tolerance. This is the path a Zeno user runs."""
import numpy as np
import pytest

from tests import util
from zeno_b200 import scenes

pytestmark = pytest.mark.gpu

DT = 0.008
G = (0.0, -9.8, 0.0)


@pytest.fixture(scope="module")
def worlds(gpu_lib, oracle_lib):
    from oracle import pyoracle
    if not pyoracle.plugin_gpu_available():
        pytest.skip("oracle/_ref/libflipplugin_gpu.so not built (needs /root/reference at build time)")
    N = 64
    pos, vel, dx = scenes.dam_break_points(N, seed=5, random_velocity=True)
    vel = vel * 0.25
    solid = scenes.box_solid_sdf(N, dx)
    pw, ow = pyoracle.PluginGpuWorld(dx), pyoracle.OracleWorld(dx)
    for w in (pw, ow):
        w.set_grid("SolidSDF", solid)
        w.PrimToVDBPointDataGrid(pos, vel)
    return pw, ow, dx


def test_substep_through_node_classes(worlds):
    pw, ow, dx = worlds
    util.compare_particles(pw.get_particles(), ow.get_particles(), "binning (OpenVDB objects)")
    for step in range(2):
        for w in (pw, ow):
            w.FLIP_P2G(dx, 3)
        for name in ("Velocity", "PostAdvVelocity", "LiquidSDF"):
            util.compare_grids(pw.get_grid(name), ow.get_grid(name), f"node FLIP_P2G {name} step {step}", tol=0.0, check_inactive=False)
        for w in (pw, ow):
            w.CutCellWeight()
        util.compare_grids(pw.get_grid("CellFWeight"), ow.get_grid("CellFWeight"), "node CutCellWeight", tol=0.0, check_inactive=False)
        for w in (pw, ow):
            w.PushOutLiquidSDF(dx)
        util.compare_grids(pw.get_grid("LiquidSDF"), ow.get_grid("LiquidSDF"), "node PushOutLiquidSDF", tol=0.0, check_inactive=False)
        for w in (pw, ow):
            w.FieldAddVector(G[0] * DT, G[1] * DT, G[2] * DT)
        util.compare_grids(pw.get_grid("Velocity"), ow.get_grid("Velocity"), "node FieldAddVector", tol=0.0, check_inactive=False)
        cfl = [w.CFL_dt() for w in (pw, ow)]
        assert cfl[0] == pytest.approx(cfl[1], rel=1e-6)
        for w in (pw, ow):
            w.AssembleSolvePPE(DT, dx)
        util.compare_grids(pw.get_grid("Divergence"), ow.get_grid("Divergence"), "node AssembleSolvePPE rhs", tol=0.0, check_inactive=False)
        util.compare_grids(pw.get_grid("Pressure"), ow.get_grid("Pressure"), "node AssembleSolvePPE pressure", tol=2e-3, check_inactive=False)
        util.sync_state(pw, ow)      # one-step-synchronised from here (the two pressures differ in the last bits)
        for w in (pw, ow):
            w.SubtractPressureGradient(DT, dx, 3)
        util.compare_grids(pw.get_grid("Velocity"), ow.get_grid("Velocity"), "node SubtractPressureGradient", tol=0.0, check_inactive=False)
        for w in (pw, ow):
            w.G2PAdvectorSheetty(DT, dx, 4, 3, 0.03, 0.05, True)
        util.compare_particles(pw.get_particles(), ow.get_particles(), f"node G2PAdvectorSheetty step {step}")


def test_nodes_beyond_the_chain(worlds):
    pw, ow, dx = worlds
    killer = scenes.sphere_sdf(centre=(6.3, 5.1, 7.7), radius=5.6, lo=(-8, -8, -8), hi=(24, 24, 24), bg=3.0)
    for w in (pw, ow):
        w.set_grid("KillerSDF", killer)
        w.ParticleAddDV(0.0, -0.05, 0.01)
        w.KillParticlesInSDF("KillerSDF", True)
    a, b = scenes.canonical_particles(pw.get_particles()), scenes.canonical_particles(ow.get_particles())
    assert 0 < b.shape[0] and a.shape == b.shape and np.array_equal(a, b)
    for w in (pw, ow):
        w.FLIP_P2G(dx, 3)
        w.VDBRenormalizeSDF("LiquidSDF", 2, 0)
    util.compare_grids(pw.get_grid("LiquidSDF"), ow.get_grid("LiquidSDF"), "node VDBRenormalizeSDF", tol=1e-6, check_inactive=False)


def test_resident_mode_sees_host_edits(worlds):
    """Resident mode skips the upload of a grid the device already holds -- but an edit of the OpenVDB object by anyone else
    (here: the test writes a scaled velocity into the host grid) must be seen by the next accelerated node."""
    pw, ow, dx = worlds
    for w in (pw, ow):
        w.FLIP_P2G(dx, 3)
    v = dict(ow.get_grid("Velocity"))
    v["values"] = (v["values"] * np.float32(0.5)).astype(np.float32)
    for w in (pw, ow):
        w.set_grid("Velocity", v)          # a host-side edit between two accelerated nodes
        w.FieldAddVector(0.0, -0.1, 0.0)
    util.compare_grids(pw.get_grid("Velocity"), ow.get_grid("Velocity"), "FieldAddVector after a host edit", tol=0.0, check_inactive=False)
