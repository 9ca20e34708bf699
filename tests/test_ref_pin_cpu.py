"""Pins the CPU oracle (oracle/*.cpp, a restatement) against the REAL reference.

(1) fixtures: tests/golden/ref_chain48.npz / ref_chain96.npz hold the outputs of the reference's own code
    (FLIP_vdb.cpp, simd_vdb_poisson_uaamg.cpp, vdb_velocity_extrapolator.cpp + OpenVDB 9.0.1, compiled
    unmodified by oracle/ref/build_ref.sh) after every node of one dam-break substep. The oracle replays
    the chain one-step-synchronised: binning + every active mask bit-exact, values to fp32 rounding,
    PCG iteration count <= 1.1 x reference, levels / DOF counts identical.
(2) live: when oracle/_ref/libflipref.so is present (this container, and the GPU box: the .so travels),
    a second seeded case is run through both on the spot.
"""
import ctypes as C
import numpy as np
import pytest

from tests import util


@pytest.mark.parametrize("case", ["ref_chain48", "ref_chain96"])
def test_oracle_matches_reference_fixture(oracle_lib, case):
    from oracle.pyoracle import OracleWorld
    fx = np.load(f"{util.GOLDEN}/{case}.npz")
    # ref_chain96 keeps only the solver's inputs/outputs (2-level multigrid), so only those nodes replay
    rep = util.replay_ref_chain(fx, OracleWorld, stages=("ppe", "grad") if case == "ref_chain96" else None)
    assert "ppe" in rep
    if case == "ref_chain96":
        assert int(fx["ppe.levels"]) >= 2  # exercises restriction / prolongation / the coarse solve
    print(case, rep)


def test_oracle_matches_reference_live(oracle_lib):
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref/libflipref.so not built (needs /root/reference); fixtures cover this case")
    from oracle.pyoracle import OracleWorld, RefWorld
    from zeno_b200 import scenes
    pyoracle.ref_set_threads(0)
    N, dt = 40, 0.006
    pos, vel, dx = scenes.dam_break_points(N, seed=11, random_velocity=True)
    vel *= np.float32(0.25)
    solid = scenes.box_solid_sdf(N, dx)
    ow, rw = OracleWorld(dx), RefWorld(dx)
    for w in (ow, rw):
        w.set_grid("SolidSDF", solid)
        w.PrimToVDBPointDataGrid(pos, vel)
    util.compare_particles(ow.get_particles(), rw.get_particles(), "live binning")
    for name, grids in util.REF_STAGES:
        util.sync_state(ow, rw)  # the oracle starts every node from the reference's state
        res = [util.run_ref_stage(w, name, dx, dt) for w in (ow, rw)]
        for g in grids:
            util.compare_grids(ow.get_grid(g), rw.get_grid(g), f"live {name}.{g}", tol=util.REF_TOL[name], check_inactive=False)
        if name == "ppe":
            assert res[0]["iterations"] <= int(np.ceil(1.1 * res[1]["iterations"])), res
        if name == "g2p":
            a = scenes.canonical_particles(ow.get_particles())
            b = scenes.canonical_particles(rw.get_particles())
            assert a.shape == b.shape
            m = util.particle_code_report(a, b)
            assert m["same_voxel"] >= 0.999 and m["P_within_1lsb"] >= 0.999 and m["v_within_1ulp"] >= 0.999, m


def test_pure_multigrid_fallback_oracle_vs_reference(oracle_lib):
    """The fallback after a failed PCG (FF/FLIP_vdb.cpp:3089-3097 -> solvePureMultigrid, uaamg.cpp:2405-2444): forced on the
    reference's own solver objects by capping the PCG at one iteration (oracle/ref/ref_driver.cpp:ref_solve_ppe_ex) and in the
    oracle the same way. Both report the failure, both end within the tolerance, same pressure."""
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref/libflipref.so not built (needs /root/reference)")
    from oracle.pyoracle import OracleWorld, RefWorld
    from zeno_b200 import scenes
    pyoracle.ref_set_threads(0)
    # 72^3 tank: 5832 pressure DOFs = two multigrid levels. (With ONE level -- <= 4000 DOFs -- the reference's muCycleIterative
    # returns from its coarsest-level branch before it has bound the caller's grids, uaamg.cpp:2153-2185 vs :2187-2188, so its
    # fallback iterates on stale internal grids, prints err = 1.0 a hundred times and hands back the warm-start pressure; the oracle
    # and the CUDA path converge instead. That degenerate case is a documented deviation)
    N, dt = 72, 0.006
    pos, vel, dx = scenes.dam_break_points(N, seed=5, random_velocity=True)
    vel *= np.float32(0.25)
    solid = scenes.box_solid_sdf(N, dx)
    ow, rw = OracleWorld(dx), RefWorld(dx)
    for w in (ow, rw):
        w.set_grid("SolidSDF", solid)
        w.PrimToVDBPointDataGrid(pos, vel)
    for name in ("p2g0", "faceweight", "pushout", "addvec", "ppe"):   # a converged previous pressure to warm-start from
        util.run_ref_stage(rw, name, dx, dt)
    util.sync_state(ow, rw)
    res = []
    for w in (ow, rw):
        w.FieldAddVector(0.0, -0.04, 0.0)
        res.append(w.AssembleSolvePPE(dt, dx, rel_tol=1e-4, max_iter=1))
    assert res[0]["status"] == 1 and res[1]["status"] == 1, res
    hist = rw.solver_info()["history"]
    assert hist.shape[0] >= 2 and hist[-1] <= 1e-4, f"the reference's pure-multigrid iteration did not converge: {hist}"
    util.compare_grids(ow.get_grid("Divergence"), rw.get_grid("Divergence"), "fallback rhs", tol=1e-6, check_inactive=False)
    util.compare_grids(ow.get_grid("Pressure"), rw.get_grid("Pressure"), "fallback pressure", tol=2e-3, check_inactive=False)


def test_reference_fraction_inside_matches_oracle(oracle_lib):
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built")
    import ctypes as C
    ref = pyoracle.load_ref()
    rng = np.random.default_rng(3)
    for _ in range(2000):
        a, b, c, d = (float(x) for x in rng.normal(0, 1, 4).astype(np.float32))
        assert oracle_lib.orc_fraction_inside2(C.c_float(a), C.c_float(b)) == ref.ref_fraction_inside2(C.c_float(a), C.c_float(b))
        x = oracle_lib.orc_fraction_inside4(C.c_float(a), C.c_float(b), C.c_float(c), C.c_float(d))
        y = ref.ref_fraction_inside4(C.c_float(a), C.c_float(b), C.c_float(c), C.c_float(d))
        assert abs(x - y) <= 1e-6, (a, b, c, d, x, y)


def test_ref_driver_equals_reference_nodes(oracle_lib):
    """oracle/ref/ref_driver.cpp (ours) claims that every ref_<node> entry point performs exactly the calls of the node shim it
    stands for. Here the reference's node classes themselves (FF/nosys/*.cpp, compiled unmodified and run through a minimal
    stand-in of the Zeno node runtime, oracle/ref/ref_nodes_test.cpp) execute the same substep from the same state: every
    grid, mask and particle must be bit-identical -- except the pressure and the velocity projected with it, where the
    reference's own TBB reductions associate differently from run to run (1e-6 relative L2 there; bit-identical too whenever
    the TBB pool really is single-threaded, which a process that already ran multi-threaded reference code cannot guarantee).
    This anchors the fixtures, the oracle pinning and bench.py's CPU arm at the reference's nodes, not at our reading of them."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_world_create"):
        pytest.skip("oracle/_ref with the reference-node harness is not available here")
    from oracle.pyoracle import RefNodeWorld, RefWorld
    from zeno_b200 import scenes
    pyoracle.ref_set_threads(1)
    try:
        N, dt = 40, 0.006
        pos, vel, dx = scenes.dam_break_points(N, seed=11, random_velocity=True)
        vel *= np.float32(0.25)
        solid = scenes.box_solid_sdf(N, dx)
        nw, rw = RefNodeWorld(dx), RefWorld(dx)
        for w in (nw, rw):
            w.set_grid("SolidSDF", solid)
            w.PrimToVDBPointDataGrid(pos, vel)
        for name, grids in util.REF_STAGES:
            util.sync_state(nw, rw)
            for w in (nw, rw):
                util.run_ref_stage(w, name, dx, dt)
            for g in grids:
                tol = 1e-6 if (name, g) in (("ppe", "Pressure"), ("grad", "Velocity")) else 0.0
                util.compare_grids(nw.get_grid(g), rw.get_grid(g), f"reference node vs ref_driver: {name}.{g}", tol=tol, check_inactive=False)
            if name == "g2p":
                util.compare_particles(nw.get_particles(), rw.get_particles(), "reference node vs ref_driver: advected particles")
            if name == "grad":
                assert nw.CFL_dt() == rw.CFL_dt()
    finally:
        pyoracle.ref_set_threads(0)


@pytest.mark.parametrize("keep", [True, False], ids=["KEEP", "DEL"])
def test_kill_particles_in_sdf_oracle_plugin_and_reference_node(oracle_lib, keep):
    """SURVEY 8f-1, the first node beyond the substep chain. The REAL reference node class (FF/nosys/KillParticles.cpp, unmodified,
    through the node-runtime stand-in), the oracle restatement and the drop-in's node (oracle behind the C ABI) filter the same
    particle set with the same killer SDF: survivors identical code for code (the reference re-encodes positions on write-back)."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_kill_particles"):
        pytest.skip("oracle/_ref with the reference-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    from zeno_b200 import scenes
    N = 32
    pos, vel, dx = scenes.dam_break_points(N, seed=9, random_velocity=True)
    killer = scenes.sphere_sdf(centre=(3.3, 4.1, 2.7), radius=4.6, lo=(-8, -8, -8), hi=(16, 16, 16), bg=3.0)
    worlds = [cls(dx) for cls in (RefNodeWorld, OracleWorld, PluginWorld)]
    for w in worlds:
        w.PrimToVDBPointDataGrid(pos, vel)
        w.set_grid("KillerSDF", killer)
    n0 = worlds[0].particles_info()[1]
    for w in worlds:
        w.KillParticlesInSDF("KillerSDF", keep)
    ref = scenes.canonical_particles(worlds[0].get_particles())
    assert 0 < ref.shape[0] < n0, "the test shape must cut the particle set"
    for w, what in zip(worlds[1:], ("oracle", "plugin node")):
        got = scenes.canonical_particles(w.get_particles())
        assert got.shape == ref.shape, f"{what}: {got.shape[0]} survivors vs {ref.shape[0]} in the reference node"
        assert np.array_equal(got, ref), f"{what}: surviving particles differ from the reference node"


def test_particle_add_dv_oracle_plugin_and_reference_node(oracle_lib):
    """ParticleAddDV (FF/nosys/ParticleAddGravity.cpp -> FLIP_vdb::point_integrate_vector): the reference's node class, the oracle
    and the drop-in's node give the same half-precision velocities, bit for bit; positions and binning untouched."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_particles_add_dv"):
        pytest.skip("oracle/_ref with the reference-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    from zeno_b200 import scenes
    pos, vel, dx = scenes.dam_break_points(32, seed=2, random_velocity=True)
    worlds = [cls(dx) for cls in (RefNodeWorld, OracleWorld, PluginWorld)]
    for w in worlds:
        w.PrimToVDBPointDataGrid(pos, vel)
    before = scenes.canonical_particles(worlds[0].get_particles())
    for w in worlds:
        w.ParticleAddDV(0.013, -0.1633333, 1.0e-4)
    ref = worlds[0].get_particles()
    after = scenes.canonical_particles(ref)
    assert np.array_equal(after[:, :6], before[:, :6]) and not np.array_equal(after[:, 6:], before[:, 6:])
    for w, what in zip(worlds[1:], ("oracle", "plugin node")):
        util.compare_particles(w.get_particles(), ref, f"ParticleAddDV: {what} vs the reference node")


def test_plain_g2p_advector_oracle_plugin_and_reference_node(oracle_lib):
    """G2P_Advector, the plain node (FF/nosys/G2P_Advector.cpp -> FLIP_vdb::Advect): no liquid SDF, no solids, every particle takes
    an Euler step whatever RK_ORDER says, FLIP factor 1 - pic_smoothness with NO clamp to 0.05 (0.3 is used here). The reference's
    node class, the oracle and the drop-in's node from the same post-P2G state."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_g2p_advect"):
        pytest.skip("oracle/_ref with the reference-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    from zeno_b200 import scenes
    N, dt = 32, 0.01
    pos, vel, dx = scenes.dam_break_points(N, seed=6, random_velocity=True)
    vel *= np.float32(0.3)
    worlds = [cls(dx) for cls in (RefNodeWorld, OracleWorld, PluginWorld)]
    for w in worlds:
        w.PrimToVDBPointDataGrid(pos, vel)
    worlds[0].FLIP_P2G(dx, 3)
    worlds[0].FieldAddVector(0.0, -9.8 * dt, 0.0)      # Velocity != PostAdvVelocity, so the FLIP part is exercised
    for w in worlds[1:]:
        util.sync_state(w, worlds[0], grids=("Velocity", "PostAdvVelocity"))
    for w in worlds:
        w.G2P_Advector(dt, dx, 3, 0.3)
    ref = scenes.canonical_particles(worlds[0].get_particles())
    for w, what in zip(worlds[1:], ("oracle", "plugin node")):
        got = scenes.canonical_particles(w.get_particles())
        assert got.shape == ref.shape, (what, got.shape, ref.shape)
        m = util.particle_code_report(got, ref)
        assert m["same_voxel"] >= 0.999 and m["P_within_1lsb"] >= 0.999 and m["v_within_1ulp"] >= 0.999, (what, m)
    a, b = (scenes.canonical_particles(w.get_particles()) for w in worlds[1:])
    assert np.array_equal(a, b), "plugin node differs from the oracle driven directly"


@pytest.mark.parametrize("iterations", [1, 4])
def test_renormalize_sdf_oracle_plugin_and_reference_node(oracle_lib, iterations):
    """VDBRenormalizeSDF (projects/zenvdb/VDBRenormalize.cpp -> openvdb::tools::LevelSetTracker::normalize, FIRST_BIAS / TVD_RK3,
    no trimming) on the liquid SDF that FLIP_P2G produced: the real node class (compiled unmodified), the oracle
    restatement and the drop-in's node. Topology and inactive values must not change; active values agree to fp32 rounding with the
    reference (its binary contracts a*b+c into FMAs, the oracle does not), and the plugin node equals the oracle bit for bit."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_renormalize_sdf"):
        pytest.skip("oracle/_ref with the reference-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    from zeno_b200 import scenes
    N = 32
    pos, vel, dx = scenes.dam_break_points(N, seed=8, random_velocity=True)
    worlds = [cls(dx) for cls in (RefNodeWorld, OracleWorld, PluginWorld)]
    for w in worlds:
        w.PrimToVDBPointDataGrid(pos, vel)
    worlds[0].FLIP_P2G(dx, 3)
    before = worlds[0].get_grid("LiquidSDF")
    for w in worlds[1:]:
        w.set_grid("LiquidSDF", before)
    for w in worlds:
        w.VDBRenormalizeSDF("LiquidSDF", iterations, 0)
    ref, orc, plg = (w.get_grid("LiquidSDF") for w in worlds)
    e = util.compare_grids(orc, ref, "VDBRenormalizeSDF: oracle vs the reference node", tol=2e-6, check_inactive=False)
    util.compare_grids(plg, orc, "VDBRenormalizeSDF: plugin node vs oracle", tol=0.0)
    # it did something, and only to active voxels
    a, b = scenes.canonical_grid(before, drop_empty=False), scenes.canonical_grid(ref, drop_empty=False)
    mb = scenes.mask_bits(a["masks"])
    assert np.array_equal(a["masks"], b["masks"]) and np.array_equal(a["values"][:, 0][~mb], b["values"][:, 0][~mb])
    assert not np.array_equal(a["values"][:, 0][mb], b["values"][:, 0][mb])
    print(f"renormalize x{iterations}: oracle vs reference node rel L2 {e:.2e}")


def test_erode_sdf_oracle_plugin_and_reference_node(oracle_lib):
    """VDBErodeSDF (projects/zenvdb/VDBRenormalize.cpp:155-185): active voxels += depth; reference node class == oracle == plugin node."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_erode_sdf"):
        pytest.skip("oracle/_ref with the reference-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    from zeno_b200 import scenes
    pos, vel, dx = scenes.dam_break_points(32, seed=8)
    worlds = [cls(dx) for cls in (RefNodeWorld, OracleWorld, PluginWorld)]
    for w in worlds:
        w.PrimToVDBPointDataGrid(pos, vel)
    worlds[0].FLIP_P2G(dx, 3)
    before = worlds[0].get_grid("LiquidSDF")
    for w in worlds[1:]:
        w.set_grid("LiquidSDF", before)
    for w in worlds:
        w.VDBErodeSDF("LiquidSDF", 0.37 * dx)
    ref = worlds[0].get_grid("LiquidSDF")
    for w, what in zip(worlds[1:], ("oracle", "plugin node")):
        util.compare_grids(w.get_grid("LiquidSDF"), ref, f"VDBErodeSDF: {what} vs the reference node", tol=0.0)
    assert not np.array_equal(scenes.canonical_grid(before)["values"], scenes.canonical_grid(ref)["values"])


@pytest.mark.parametrize("width,iterations", [(1, 1), (2, 2)])
def test_smooth_sdf_oracle_plugin_and_reference_node(oracle_lib, width, iterations):
    """VDBSmoothSDF (projects/zenvdb/VDBRenormalize.cpp:108-133 -> openvdb::tools::Filter::gaussian): the real node class with the
    vendored OpenVDB filter, the oracle restatement (four X/Z/Y box-filter rounds per iteration) and the drop-in's node."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_smooth_sdf"):
        pytest.skip("oracle/_ref with the reference-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    from zeno_b200 import scenes
    pos, vel, dx = scenes.dam_break_points(32, seed=8)
    worlds = [cls(dx) for cls in (RefNodeWorld, OracleWorld, PluginWorld)]
    for w in worlds:
        w.PrimToVDBPointDataGrid(pos, vel)
    worlds[0].FLIP_P2G(dx, 3)
    before = worlds[0].get_grid("LiquidSDF")
    for w in worlds[1:]:
        w.set_grid("LiquidSDF", before)
    for w in worlds:
        w.VDBSmoothSDF("LiquidSDF", width, iterations)
    ref, orc, plg = (w.get_grid("LiquidSDF") for w in worlds)
    util.compare_grids(orc, ref, "VDBSmoothSDF: oracle vs the reference node", tol=0.0)
    util.compare_grids(plg, orc, "VDBSmoothSDF: plugin node vs oracle", tol=0.0)
    assert not np.array_equal(scenes.canonical_grid(before)["values"], scenes.canonical_grid(ref)["values"])


def _leaf_slices(p):
    counts = p["voxel_end"][:, 511].astype(np.int64)
    begin = np.concatenate([[0], np.cumsum(counts)])
    return begin


def _one_leaf(p, i):
    b = _leaf_slices(p)
    return {"origins": p["origins"][i:i + 1], "voxel_end": p["voxel_end"][i:i + 1], "P": p["P"][b[i]:b[i + 1]], "v": p["v"][b[i]:b[i + 1]]}


def _same_leaf(a, i, b, j):
    ba, bb = _leaf_slices(a), _leaf_slices(b)
    return (np.array_equal(a["origins"][i], b["origins"][j]) and np.array_equal(a["voxel_end"][i], b["voxel_end"][j])
            and np.array_equal(a["P"][ba[i]:ba[i + 1]], b["P"][bb[j]:bb[j + 1]]) and np.array_equal(a["v"][ba[i]:ba[i + 1]], b["v"][bb[j]:bb[j + 1]]))


def _reseed_scene(seed=5, ppc=3, side=20, W=2):
    """A 20^3-voxel block (27 particle leaves, all inside one 128^3 node so that tree order == store order) at 3 particles per
    voxel. FLIP_P2G's liquid SDF never goes below about -0.8 dx, and the reseeder only emits where it is <= -dx (the packaged
    graph renormalises it first), so the worlds get an analytic SDF of the block on the P2G topology: depth in voxels below the
    block's faces, with a ripple so that candidate acceptance is not uniform."""
    from zeno_b200 import scenes
    pos, vel, dx = scenes.dam_break_points(64, seed=seed, ppc=ppc, side=side, W=W, random_velocity=True)
    return pos, vel * np.float32(0.3), dx, (W, side)


def _analytic_liquid_sdf(grid, dx, W, side):
    g = {k: np.array(v, copy=True) for k, v in grid.items()}
    o = g["origins"].astype(np.float64)                                  # [n,3]
    off = np.arange(512)
    loc = np.stack([off >> 6, (off >> 3) & 7, off & 7], axis=1).astype(np.float64)  # [512,3]
    x = o[:, None, :] + loc[None, :, :]                                  # voxel centres, index space
    c, h = W + side / 2.0 - 0.5, side / 2.0
    d = np.max(np.abs(x - c) - h, axis=2)                                # < 0 inside the block
    d += 0.35 * np.sin(1.7 * x[..., 0] + 0.9 * x[..., 1]) * np.cos(1.3 * x[..., 2])
    g["values"] = (d * dx).astype(np.float32).reshape(g["values"].shape)
    return g


def _reseed_worlds(classes, seed=5):
    pos, vel, dx, (W, side) = _reseed_scene(seed=seed)
    worlds = [cls(dx) for cls in classes]
    for w in worlds:
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIP_P2G(dx, 3)
    sdf = _analytic_liquid_sdf(worlds[-1].get_grid("LiquidSDF"), dx, W, side)
    vel = worlds[-1].get_grid("Velocity")   # one velocity field for all (the reference's own P2G sums in another order: last bits differ)
    for w in worlds:
        w.set_grid("LiquidSDF", sdf)
        w.set_grid("Velocity", vel)
    return worlds, dx


@pytest.mark.parametrize("threads", [1, 0], ids=["one_thread", "all_threads"])
def test_fluid_reseed_oracle_vs_reference_node(oracle_lib, threads):
    """FluidReseed (FF/nosys/FLIP_Reseed.cpp -> FLIP_vdb::reseed_fluid, FF/FLIP_vdb.cpp:2047-2220), the REAL node class in the
    seeded build of the reference (std::random_device replaced by a fixed seed at compile time, oracle/ref/shims/seeded_random.h;
    the sources are untouched). The reference starts its jitter table once per TBB chunk and runs on through the chunk's leaves,
    and the chunking is the scheduler's business; so every leaf of its result must be what the oracle's per-leaf restatement gives
    when started EITHER at the chunk start the seed implies OR where the previous leaf ended. Code-for-code equality, every leaf."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_fluid_reseed"):
        pytest.skip("oracle/_ref with the FluidReseed reference node is not available here")
    from oracle.pyoracle import OracleWorld, RefNodeWorld
    seed = 20240
    lib = pyoracle.load()
    lib.orc_reseed_chunk_start.restype = C.c_uint64
    s0 = int(lib.orc_reseed_chunk_start(C.c_uint32(seed)))
    (rw, ow), dx = _reseed_worlds((RefNodeWorld, OracleWorld))
    before = ow.get_particles()
    util.compare_particles(rw.get_particles(), before, "reseed input: reference vs oracle store")
    try:
        pyoracle.ref_set_threads(threads)
        rw.FluidReseed(seed)
    finally:
        pyoracle.ref_set_threads(0)
    ref = rw.get_particles()
    n0, n1 = before["P"].shape[0], ref["P"].shape[0]
    assert n1 > n0 * 1.5, f"the scene must make the reseeder work: {n0} -> {n1} particles"
    assert np.array_equal(ref["origins"], before["origins"]), "the reference keeps the leaf set (store order == tree order inside one 128^3 block)"
    # the oracle, one leaf at a time on the same grids
    lw = OracleWorld(dx)
    for g in ("LiquidSDF", "Velocity"):
        lw.set_grid(g, ow.get_grid(g))
    nl = before["origins"].shape[0]
    prev_ends = set()
    chunk_starts = 0
    for i in range(nl):
        ok_ends = set()
        for c in [s0] + sorted(prev_ends - {s0}):
            lw.set_particles(_one_leaf(before, i))
            end = lw.FluidReseed(0, leaf_start=np.array([c], np.uint64), want_leaf_end=True)
            if _same_leaf(lw.get_particles(), 0, ref, i):
                ok_ends.add(int(end[0]))
                chunk_starts += int(c == s0 and int(end[0]) != c)
        assert ok_ends, (f"leaf {i} (origin {before['origins'][i]}) of the reference's result matches the oracle neither from the chunk start {s0} "
                         f"nor from the previous leaf's end {sorted(prev_ends)}")
        prev_ends = ok_ends
    assert chunk_starts >= 1


def test_fluid_reseed_plugin_node_equals_oracle(oracle_lib):
    """The drop-in's FluidReseed node (oracle behind the C ABI) against the oracle, both with the seeded per-leaf starts of
    include/flipb200.h: identical stores."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "pn_fluid_reseed"):
        pytest.skip("oracle/_ref with the plugin-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld
    (pw, ow), dx = _reseed_worlds((PluginWorld, OracleWorld), seed=6)
    n0 = ow.particles_info()[1]
    for w in (pw, ow):
        w.FluidReseed(77)
    assert ow.particles_info()[1] > 1.5 * n0
    util.compare_particles(pw.get_particles(), ow.get_particles(), "FluidReseed: plugin node vs oracle")
    # a second call tops up nothing where every voxel already holds > 4 particles or every octant is taken: the store stops growing
    n1 = ow.particles_info()[1]
    ow.FluidReseed(78)
    n2 = ow.particles_info()[1]
    assert n1 <= n2 < n1 * 1.2


def _emitter_scene(seed=3):
    """A 12^3 block at 8 particles per voxel and a sphere shape (world-unit distances on the world's own transform) that overlaps
    the block's corner and reaches into empty space: existing leaves get topped up, new leaves are created, leaves of the block
    the sphere does not touch must come through untouched."""
    from zeno_b200 import scenes
    pos, vel, dx = scenes.dam_break_points(64, seed=seed, ppc=8, side=12, W=2, random_velocity=True)
    shape = scenes.sphere_sdf(centre=(17.3, 12.1, 9.7), radius=7.6, lo=(0, 0, 0), hi=(32, 24, 24), bg=3.0)
    shape["values"] = (shape["values"] * np.float32(dx)).astype(np.float32)
    shape["bg"] = np.array([3.0 * dx], np.float32)
    return pos, vel * np.float32(0.3), dx, shape


def _by_origin(p):
    b = _leaf_slices(p)
    return {tuple(int(x) for x in p["origins"][i]): i for i in range(p["origins"].shape[0]) if b[i + 1] > b[i]}


@pytest.mark.parametrize("threads", [1, 0], ids=["one_thread", "all_threads"])
def test_particle_emitter_oracle_vs_reference_node(oracle_lib, threads):
    """ParticleEmitter (FF/nosys/ParticleEmitter.cpp -> FLIP_vdb::emit_liquid, FF/FLIP_vdb.cpp:2222-2642), constant-velocity branch, the
    REAL node class in the seeded build of the reference. The reference walks its touched leaves in the order of a concurrent hash
    map, one jitter-table start per TBB chunk, so a leaf of its result must equal the oracle's per-leaf restatement started either
    at the chunk start the seed implies or where SOME other leaf ended. Same leaf set, and code-for-code equality on every leaf."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_emit_liquid"):
        pytest.skip("oracle/_ref with the ParticleEmitter reference node is not available here")
    from oracle.pyoracle import OracleWorld, RefNodeWorld
    pos, vel, dx, shape = _emitter_scene()
    seed = 777
    lib = pyoracle.load()
    lib.orc_reseed_chunk_start.restype = C.c_uint64
    s0 = int(lib.orc_reseed_chunk_start(C.c_uint32(seed)))
    rw, ow = RefNodeWorld(dx), OracleWorld(dx)
    for w in (rw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.set_grid("KillerSDF", shape)
    before = ow.get_particles()
    V = (0.25, -1.5, 0.125)
    try:
        pyoracle.ref_set_threads(threads)
        rw.ParticleEmitter("KillerSDF", *V, seed=seed)
    finally:
        pyoracle.ref_set_threads(0)
    ref = rw.get_particles()
    ends = ow.ParticleEmitter("KillerSDF", *V, seed=seed, want_leaf_end=True, max_leaves=4096)
    orc = ow.get_particles()
    n0, n1 = before["P"].shape[0], ref["P"].shape[0]
    assert n1 > n0 + 1000, f"the emitter must add particles: {n0} -> {n1}"
    ro, oo = _by_origin(ref), _by_origin(orc)
    assert set(ro) == set(oo), "non-empty leaves of the reference's result vs the oracle's"
    assert len(ro) > len(_by_origin(before)), "the scene must create leaves"
    touched = {tuple(int(x) for x in orc["origins"][i]) for i in range(orc["origins"].shape[0]) if int(ends[i]) != 0xFFFFFFFFFFFFFFFF}
    # untouched leaves: identical to the input; touched ones: the oracle leaf by leaf from the candidate starts
    bo = _by_origin(before)
    for k, i in ro.items():
        if k not in touched:
            assert _same_leaf(ref, i, before, bo[k]), f"leaf {k} is not touched by the shape and must be unchanged"
    lw = OracleWorld(dx)
    lw.set_grid("KillerSDF", shape)
    pending = [k for k in sorted(ro) if k in touched]
    known_ends = set()
    progress = True
    while pending and progress:
        progress = False
        for k in list(pending):
            one = _one_leaf(before, bo[k]) if k in bo else {"origins": np.zeros((0, 3), np.int32), "voxel_end": np.zeros((0, 512), np.uint32),
                                                          "P": np.zeros((0, 3), np.uint16), "v": np.zeros((0, 3), np.uint16)}
            for c in [s0] + sorted(known_ends - {s0}):
                lw.set_particles(one)
                e = lw.ParticleEmitter("KillerSDF", *V, seed=0, leaf_start=np.full(4096, c, np.uint64), want_leaf_end=True, max_leaves=4096)
                got = lw.get_particles()
                j = _by_origin(got).get(k)
                if j is not None and _same_leaf(got, j, ref, ro[k]):
                    known_ends.add(int(e[j]))
                    pending.remove(k)
                    progress = True
                    break
    assert not pending, f"{len(pending)} leaves of the reference's result match the oracle from no candidate start, e.g. {pending[:3]}"


def test_particle_emitter_plugin_node_equals_oracle(oracle_lib):
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "pn_emit_liquid"):
        pytest.skip("oracle/_ref with the plugin-node harness is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld
    pos, vel, dx, shape = _emitter_scene(seed=4)
    pw, ow = PluginWorld(dx), OracleWorld(dx)
    for w in (pw, ow):
        w.PrimToVDBPointDataGrid(pos, vel)
        w.set_grid("KillerSDF", shape)
        w.ParticleEmitter("KillerSDF", 0.5, 0.0, -0.75, seed=99)
    util.compare_particles(pw.get_particles(), ow.get_particles(), "ParticleEmitter: plugin node vs oracle")
    # emitting into an EMPTY world creates the store
    ew = OracleWorld(dx)
    ew.set_grid("KillerSDF", shape)
    ew.ParticleEmitter("KillerSDF", 0.0, 1.0, 0.0, seed=5)
    assert ew.particles_info()[1] > 1000


def _boundary_scene(seed=3):
    """The tank's analytic solid SDF (vertex centred), a 12^3 water block, and a moving sphere (cell-centred grid, world-unit
    distances) that overlaps tank leaves, water leaves and empty space."""
    from zeno_b200 import scenes
    N = 64
    pos, vel, dx = scenes.dam_break_points(N, seed=seed, ppc=8, side=12, W=2, random_velocity=True)
    solid = scenes.box_solid_sdf(N, dx)
    sphere = scenes.sphere_sdf(centre=(15.3, 9.1, 11.7), radius=6.4, lo=(0, -8, 0), hi=(32, 24, 24), bg=3.0)
    sphere["values"] = (sphere["values"] * np.float32(dx)).astype(np.float32)
    sphere["bg"] = np.array([3.0 * dx], np.float32)
    return pos, vel, dx, solid, sphere


def test_flip_apply_boundary_oracle_plugin_and_reference_node(oracle_lib):
    """FLIPApplyBoundary (FF/nosys/Update_Solid_SDF.cpp -> FLIP_vdb::update_solid_sdf, FF/FLIP_vdb.cpp:1976-2046): the REAL node class,
    the oracle and the drop-in's node merge the same moving sphere into the same static SDF: same leaves, every voxel active,
    values bit for bit."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_apply_boundary"):
        pytest.skip("oracle/_ref with the FLIPApplyBoundary reference node is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    pos, vel, dx, solid, sphere = _boundary_scene()
    worlds = [cls(dx) for cls in (RefNodeWorld, OracleWorld, PluginWorld)]
    for w in worlds:
        w.set_grid("SolidSDF", solid)
        w.set_grid("KillerSDF", sphere)
        w.PrimToVDBPointDataGrid(pos, vel)
        w.FLIPApplyBoundary("KillerSDF")
    ref = worlds[0].get_grid("SolidSDF")
    assert ref["origins"].shape[0] > solid["origins"].shape[0], "the moving solid must add leaves"
    assert np.all(ref["masks"] == np.uint64(0xFFFFFFFFFFFFFFFF))
    assert int((ref["values"] < 0).sum()) > int((solid["values"] < 0).sum()) + 500, "the sphere must add solid voxels"
    for w, what in zip(worlds[1:], ("oracle", "plugin node")):
        util.compare_grids(w.get_grid("SolidSDF"), ref, f"FLIPApplyBoundary: {what} vs the reference node", tol=0.0, check_inactive=True)


def _tension_worlds(classes, seed=4):
    """A dam-break state right before the pressure solve, a synthetic curvature field on the liquid SDF's leaves (the reference reads
    it by voxel coordinate), Density 1000 and SurfaceTension 5 (tension = 2 coef / density = 0.01: ghost pressures of +-0.2)."""
    from zeno_b200 import scenes
    N = 48
    pos, vel, dx = scenes.dam_break_points(N, seed=seed, random_velocity=True)
    solid = scenes.box_solid_sdf(N, dx)
    worlds = [cls(dx) for cls in classes]
    dt = 0.006
    for w in worlds:
        w.set_grid("SolidSDF", solid)
        w.PrimToVDBPointDataGrid(pos, vel * np.float32(0.3))
        w.FLIP_P2G(dx, 3)
        w.CutCellWeight()
        w.PushOutLiquidSDF(dx)
        w.FieldAddVector(0.0, -9.8 * dt, 0.0)
    sdf = worlds[-1].get_grid("LiquidSDF")
    curv = {k: np.array(v, copy=True) for k, v in sdf.items()}
    o = curv["origins"].astype(np.float64)
    off = np.arange(512)
    loc = np.stack([off >> 6, (off >> 3) & 7, off & 7], axis=1).astype(np.float64)
    x = o[:, None, :] + loc[None, :, :]
    curv["values"] = (20.0 * np.sin(0.7 * x[..., 0] + 0.3 * x[..., 2]) * np.cos(0.5 * x[..., 1])).astype(np.float32).reshape(curv["values"].shape)
    curv["bg"] = np.array([0.0], np.float32)
    for w in worlds:
        w.set_grid("Curvature", curv)
        w.set_surface_tension(1000.0, 5.0)
    return worlds, dx, dt


def test_surface_tension_oracle_plugin_and_reference_nodes(oracle_lib):
    """SURVEY 8f-4, the tension terms of the two core nodes: BuildPoissonRhs_withTension (FF/simd_vdb_poisson_uaamg.cpp:95-209) inside
    AssembleSolvePPE and the ghost pressure of apply_pressure_gradient (FF/FLIP_vdb.cpp:2932-2939) inside SubtractPressureGradient.
    The REAL node classes (Density / SurfaceTension / Curvature sockets wired), the oracle and the drop-in's nodes from the same state."""
    from oracle import pyoracle
    if not pyoracle.ref_available() or not hasattr(pyoracle.load_ref(), "rn_set_surface_tension"):
        pytest.skip("oracle/_ref with the tension hooks is not available here")
    from oracle.pyoracle import OracleWorld, PluginWorld, RefNodeWorld
    (rw, ow, pw), dx, dt = _tension_worlds((RefNodeWorld, OracleWorld, PluginWorld))
    plain = OracleWorld(dx)   # the same state without tension: the terms must matter in this scene
    for g in ("SolidSDF", "LiquidSDF", "Velocity", "CellFWeight", "SolidVelocity"):
        try:
            plain.set_grid(g, ow.get_grid(g))
        except Exception:
            pass
    for w in (rw, ow, pw, plain):
        w.AssembleSolvePPE(dt, dx)
    ref_rhs, ref_p = rw.get_grid("Divergence"), rw.get_grid("Pressure")
    e0 = util.compare_grids(plain.get_grid("Divergence"), ow.get_grid("Divergence"), "rhs without vs with tension", tol=10.0, check_inactive=False)
    assert e0 > 1e-2, f"the tension term must change the right-hand side (relative L2 {e0})"
    util.compare_grids(ow.get_grid("Divergence"), ref_rhs, "tension RHS: oracle vs reference node", tol=1e-6, check_inactive=False)
    util.compare_grids(ow.get_grid("Pressure"), ref_p, "tension pressure: oracle vs reference node", tol=util.REF_TOL["ppe"], check_inactive=False)
    util.compare_grids(pw.get_grid("Divergence"), ow.get_grid("Divergence"), "tension RHS: plugin node vs oracle", tol=0.0, check_inactive=False)
    util.compare_grids(pw.get_grid("Pressure"), ow.get_grid("Pressure"), "tension pressure: plugin node vs oracle", tol=0.0, check_inactive=False)
    # the gradient from ONE pressure field (the reference's), so that only the gradient's own arithmetic is compared
    for w in (ow, pw):
        w.set_grid("Pressure", ref_p)
    for w in (rw, ow, pw):
        w.SubtractPressureGradient(dt, dx, 3)
    util.compare_grids(ow.get_grid("Velocity"), rw.get_grid("Velocity"), "tension gradient: oracle vs reference node", tol=1e-6, check_inactive=False)
    util.compare_grids(pw.get_grid("Velocity"), ow.get_grid("Velocity"), "tension gradient: plugin node vs oracle", tol=0.0, check_inactive=False)
