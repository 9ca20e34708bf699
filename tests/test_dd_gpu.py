"""GPU tests of the slab decomposition: the same dam break run (a) on one world and (b) split into
slabs along x over several ranks. The ranks here are worlds of ONE process on cuda:0, connected by the library's
in-process communicator (flipb200_comm_init_local), which carries exactly the message sequence of the NCCL backend --
so the decomposition is exercised on a single-GPU box. tests/test_nccl_gpu.py repeats the check over NCCL when the box
has two GPUs.

Bars: particle routing / ghost import, every active mask and the P2G + stencil values of the OWNED leaves are bit-exact
against the single world (the migration keeps the undecomposed store order inside every voxel); the distributed MGPCG
differs from the single-GPU one only in the association of its dot products: same iteration count +-1, pressure and
projected velocity within 1e-5 relative L2.
"""
import numpy as np
import pytest

from tests import util
from zeno_b200 import scenes

pytestmark = pytest.mark.gpu

DT = 0.008
G = (0.0, -9.8, 0.0)
OPEN = 1 << 29


def owned_part(g, lo, hi):
    lx = g["origins"][:, 0] >> 3
    k = (lx >= lo) & (lx < hi)
    return {"origins": g["origins"][k], "masks": g["masks"][k], "values": g["values"][k], "bg": g["bg"]}


def merge_grids(parts):
    return {"origins": np.concatenate([p["origins"] for p in parts]), "masks": np.concatenate([p["masks"] for p in parts]),
            "values": np.concatenate([p["values"] for p in parts]), "bg": parts[0]["bg"]}


def owned_particles(p, lo, hi):
    """rows (voxel, P, v) of the particles whose leaf lies in [lo, hi)"""
    rows = scenes.canonical_particles(p)
    lx = rows[:, 0] >> 3
    return rows[(lx >= lo) & (lx < hi)]


def split_points(pos, vel, dx, bounds):
    ijk = np.floor(pos.astype(np.float64) * (1.0 / np.float64(np.float32(dx))) + 0.5).astype(np.int64)
    lx = ijk[:, 0] >> 3
    out = []
    for r, (lo, hi) in enumerate(bounds):
        a = -OPEN if r == 0 else lo
        b = OPEN if r == len(bounds) - 1 else hi
        k = (lx >= a) & (lx < b)
        out.append((pos[k], vel[k]))
    return out


def make_dd(abi, N, bounds, pos, vel, dx, solid):
    worlds = [abi.World(dx) for _ in bounds]
    abi.comm_init_local(worlds)
    parts = split_points(pos, vel, dx, bounds)

    def setup(r, w):
        w.dd_set_slab(*bounds[r])
        w.set_grid("SolidSDF", solid)
        w.PrimToVDBPointDataGrid(*parts[r])
    abi.run_ranks(worlds, setup)
    return worlds


def gather_grid(worlds, name):
    return merge_grids([owned_part(w.get_grid(name), *w.dd_owned()) for w in worlds])


def gather_particles(worlds):
    return np.concatenate([owned_particles(w.get_particles(), *w.dd_owned()) for w in worlds])


@pytest.mark.parametrize("bounds", [[(0, 2), (2, 4)], [(0, 2), (2, 4), (4, 6)]], ids=["2ranks", "3ranks"])
def test_slab_decomposition_matches_single_world(gpu_lib, bounds):
    from zeno_b200 import abi
    N = 128
    side = 32 if len(bounds) == 2 else 48
    pos, vel, dx = scenes.dam_break_points(N, seed=3, random_velocity=True, side=side)
    vel = vel * 0.2
    solid = scenes.box_solid_sdf(N, dx)
    one = abi.World(dx)
    one.set_grid("SolidSDF", solid)
    one.PrimToVDBPointDataGrid(pos, vel)
    dd = make_dd(abi, N, bounds, pos, vel, dx, solid)

    # -- routing + ghost import: owned particles are exactly the single world's; every rank also holds the
    #    neighbours' boundary layer
    ref_rows = scenes.canonical_particles(one.get_particles())
    got = gather_particles(dd)
    got = got[np.lexsort(tuple(got[:, k] for k in range(8, -1, -1)))]
    assert np.array_equal(got, ref_rows), "owned particles differ from the single world after binning"
    for r, w in enumerate(dd):
        lo, hi = w.dd_owned()
        rows = scenes.canonical_particles(w.get_particles())
        lx = rows[:, 0] >> 3
        assert lx.min() >= lo - 1 and lx.max() <= hi, "a rank stores particles beyond its ghost layers"
        want = ref_rows[((ref_rows[:, 0] >> 3) >= lo - 1) & ((ref_rows[:, 0] >> 3) < hi + 1)]
        assert np.array_equal(rows, want), f"rank {r}: owned + ghost particle set differs"

    # -- P2G: bit-exact on owned leaves, and the refreshed ghost layers equal the owner's values
    one.FLIP_P2G(dx, 3)
    abi.run_ranks(dd, lambda r, w: w.FLIP_P2G(dx, 3))
    for name in ("Velocity", "PostAdvVelocity", "LiquidSDF"):
        util.compare_grids(gather_grid(dd, name), one.get_grid(name), f"dd P2G {name}", tol=0.0, check_inactive=False)
    ref_v = one.get_grid("Velocity")
    for r, w in enumerate(dd):
        lo, hi = w.dd_owned()
        util.compare_grids(owned_part(w.get_grid("Velocity"), lo - 1, hi + 1), owned_part(ref_v, lo - 1, hi + 1),
                           f"rank {r}: Velocity incl. ghost layers", tol=0.0, check_inactive=False)

    def chain(w):
        w.CutCellWeight()
        w.PushOutLiquidSDF(dx)
        w.FieldAddVector(G[0] * DT, G[1] * DT, G[2] * DT)
        return w.AssembleSolvePPE(DT, dx)

    res1 = chain(one)
    resd = abi.run_ranks(dd, lambda r, w: chain(w))
    for name in ("CellFWeight", "LiquidSDF", "Velocity"):
        util.compare_grids(gather_grid(dd, name), one.get_grid(name), f"dd stencils {name}", tol=0.0, check_inactive=False)
    # -- the sharded MGPCG
    info1, infod = one.solver_info(), dd[0].solver_info()
    assert all(r["status"] == 0 for r in resd) and res1["status"] == 0, (res1, resd)
    assert len({r["iterations"] for r in resd}) == 1, "ranks disagree on the iteration count"
    assert abs(resd[0]["iterations"] - res1["iterations"]) <= 1, (res1, resd)
    assert infod["levels"] == info1["levels"] and infod["num_dof"] == info1["num_dof"], (info1, infod)
    e_p = util.compare_grids(gather_grid(dd, "Pressure"), one.get_grid("Pressure"), "dd Pressure", tol=1e-5, check_inactive=False)
    util.compare_grids(gather_grid(dd, "Divergence"), one.get_grid("Divergence"), "dd Divergence", tol=0.0, check_inactive=False)

    one.SubtractPressureGradient(DT, dx, 3)
    abi.run_ranks(dd, lambda r, w: w.SubtractPressureGradient(DT, dx, 3))
    e_v = util.compare_grids(gather_grid(dd, "Velocity"), one.get_grid("Velocity"), "dd projected Velocity", tol=1e-5, check_inactive=False)

    # -- advection + migration (RK3), then a full second substep through flipb200_substep
    one.G2PAdvectorSheetty(DT, dx, 4, 3, 0.03, 0.05, True)
    abi.run_ranks(dd, lambda r, w: w.G2PAdvectorSheetty(DT, dx, 4, 3, 0.03, 0.05, True))
    a = scenes.canonical_particles(one.get_particles())
    b = gather_particles(dd)
    b = b[np.lexsort(tuple(b[:, k] for k in range(8, -1, -1)))]
    assert a.shape == b.shape, f"particle count after advection {b.shape[0]} vs {a.shape[0]}"
    same = (a[:, :3] == b[:, :3]).all(axis=1).mean()
    assert same > 0.999, same
    dt1 = one.CFL_dt()
    dts = abi.run_ranks(dd, lambda r, w: w.CFL_dt())
    assert all(abs(d - dt1) <= 1e-5 * dt1 for d in dts), (dt1, dts)
    one.substep(DT, dx, 4, 3, 0.03, 0.05, G, 3, True)
    abi.run_ranks(dd, lambda r, w: w.substep(DT, dx, 4, 3, 0.03, 0.05, G, 3, True))
    a = scenes.canonical_particles(one.get_particles())
    b = gather_particles(dd)
    assert a.shape == b.shape, f"particle count after the second substep {b.shape[0]} vs {a.shape[0]}"
    v1, vd = one.get_grid("Velocity"), gather_grid(dd, "Velocity")
    c1, cd = scenes.canonical_grid(v1), scenes.canonical_grid(vd)
    assert abs(c1["origins"].shape[0] - cd["origins"].shape[0]) <= max(2, c1["origins"].shape[0] // 100)
    print(f"dd ok ({len(bounds)} ranks): iterations {resd[0]['iterations']}/{res1['iterations']}, pressure rel L2 {e_p:.2e}, "
          f"velocity rel L2 {e_v:.2e}, same-voxel after advection {same:.6f}")
    for w in dd + [one]:
        w.close()


def test_slab_rules_are_enforced(gpu_lib):
    """include/flipb200.h: a slab holds at least two leaf layers, and the ranks' slabs continue each other without gap or
    overlap. A thin slab is refused at once; a gap is found at the first exchange, on every rank together."""
    from zeno_b200 import abi
    N = 128
    pos, vel, dx = scenes.dam_break_points(N, seed=3, side=32)
    solid = scenes.box_solid_sdf(N, dx)
    worlds = [abi.World(dx) for _ in range(2)]
    abi.comm_init_local(worlds)
    with pytest.raises(abi.FlipB200Error, match="at least two leaf layers"):
        worlds[0].dd_set_slab(0, 1)
    bounds = [(0, 2), (3, 5)]   # layer 2 belongs to nobody
    parts = split_points(pos, vel, dx, bounds)

    def setup(r, w):
        w.dd_set_slab(*bounds[r])
        w.set_grid("SolidSDF", solid)
        try:
            w.PrimToVDBPointDataGrid(*parts[r])
        except abi.FlipB200Error as e:
            return str(e)
        return None
    errs = abi.run_ranks(worlds, setup)
    assert all(e is not None and "contiguous" in e for e in errs), errs
    for w in worlds:
        w.close()


@pytest.mark.parametrize("bounds", [[(0, 2), (2, 4)], [(0, 2), (2, 4), (4, 6)]], ids=["2ranks", "3ranks"])
def test_sdf_nodes_under_decomposition(gpu_lib, bounds):
    """VDBRenormalizeSDF (4 iterations) and VDBSmoothSDF on the liquid SDF of a decomposed world: the ghost layers are refreshed
    between the passes that read across a slab face, so the owned leaves equal the single world's bit for bit."""
    from zeno_b200 import abi
    N = 128
    side = 32 if len(bounds) == 2 else 48
    pos, vel, dx = scenes.dam_break_points(N, seed=5, random_velocity=True, side=side)
    solid = scenes.box_solid_sdf(N, dx)
    one = abi.World(dx)
    one.set_grid("SolidSDF", solid)
    one.PrimToVDBPointDataGrid(pos, vel)
    dd = make_dd(abi, N, bounds, pos, vel, dx, solid)
    one.FLIP_P2G(dx, 3)
    abi.run_ranks(dd, lambda r, w: w.FLIP_P2G(dx, 3))
    one.VDBRenormalizeSDF("LiquidSDF", 4)
    abi.run_ranks(dd, lambda r, w: w.VDBRenormalizeSDF("LiquidSDF", 4))
    util.compare_grids(gather_grid(dd, "LiquidSDF"), one.get_grid("LiquidSDF"), "dd VDBRenormalizeSDF", tol=0.0, check_inactive=False)
    for width in (1, 2):
        one.VDBSmoothSDF("LiquidSDF", width, 1)
        abi.run_ranks(dd, lambda r, w: w.VDBSmoothSDF("LiquidSDF", width, 1))
        util.compare_grids(gather_grid(dd, "LiquidSDF"), one.get_grid("LiquidSDF"), f"dd VDBSmoothSDF width {width}", tol=0.0, check_inactive=False)
    for w in dd + [one]:
        w.close()


def _sorted_rows(rows):
    return rows[np.lexsort(tuple(rows[:, k] for k in range(8, -1, -1)))]


@pytest.mark.parametrize("bounds", [[(0, 2), (2, 4)], [(0, 2), (2, 4), (4, 6)]], ids=["2ranks", "3ranks"])
def test_reseed_and_emitter_under_decomposition(gpu_lib, bounds):
    """FluidReseed and ParticleEmitter on a decomposed world. A leaf's draws start at a hash of (seed, leaf origin) and read only
    the leaf's own particles and grids within two voxels of it, so the owner of a leaf and the neighbour that holds it as a ghost
    decide alike with no exchange: owned particles == the single world's, every rank's ghost layer == the owner's copy, and the
    P2G that follows is bit-identical on the owned leaves."""
    from zeno_b200 import abi
    N = 128
    side = 32 if len(bounds) == 2 else 48
    pos, vel, dx = scenes.dam_break_points(N, seed=9, random_velocity=True, side=side)
    rng = np.random.default_rng(9)
    keep = rng.random(pos.shape[0]) < 0.4          # ~3 particles per voxel: the reseeder has work everywhere
    pos, vel = pos[keep], (vel[keep] * 0.2).astype(np.float32)
    solid = scenes.box_solid_sdf(N, dx)
    one = abi.World(dx)
    one.set_grid("SolidSDF", solid)
    one.PrimToVDBPointDataGrid(pos, vel)
    dd = make_dd(abi, N, bounds, pos, vel, dx, solid)
    one.FLIP_P2G(dx, 3)
    abi.run_ranks(dd, lambda r, w: w.FLIP_P2G(dx, 3))

    def check(what):
        ref_rows = scenes.canonical_particles(one.get_particles())
        got = _sorted_rows(gather_particles(dd))
        assert got.shape == ref_rows.shape, f"{what}: {got.shape[0]} owned particles vs {ref_rows.shape[0]} in the single world"
        assert np.array_equal(got, ref_rows), f"{what}: owned particles differ from the single world"
        for r, w in enumerate(dd):
            lo, hi = w.dd_owned()
            rows = scenes.canonical_particles(w.get_particles())
            lx = ref_rows[:, 0] >> 3
            want = ref_rows[(lx >= lo - 1) & (lx < hi + 1)]
            assert np.array_equal(rows, want), f"{what}: rank {r} owned + ghost particle set differs"

    # FLIP_P2G's liquid SDF never goes below about -0.8 dx and the reseeder only emits where it is <= -dx (the packaged graph
    # renormalises it first): every world gets the analytic SDF of the block, as in tests/test_ref_pin_cpu.py
    from tests.test_ref_pin_cpu import _analytic_liquid_sdf
    sdf = _analytic_liquid_sdf(one.get_grid("LiquidSDF"), dx, 0, side)
    one.set_grid("LiquidSDF", sdf)
    abi.run_ranks(dd, lambda r, w: w.set_grid("LiquidSDF", sdf))
    n0 = one.particles_info()[1]
    one.FluidReseed(77)
    abi.run_ranks(dd, lambda r, w: w.FluidReseed(77))
    assert one.particles_info()[1] > 1.3 * n0, "the scene must make the reseeder work"
    check("FluidReseed")

    # a sphere across the slab face(s), partly over the block, partly in empty space (new leaves on both sides of a face)
    cx = 8.0 * bounds[0][1] + 1.3
    top = float(side)
    shape = scenes.sphere_sdf(centre=(cx, top + 2.1, 9.7), radius=7.6, lo=(int(cx) - 16, int(top) - 16, -8), hi=(int(cx) + 16, int(top) + 16, 24), bg=3.0)
    shape["values"] = (shape["values"] * np.float32(dx)).astype(np.float32)
    shape["bg"] = np.array([3.0 * dx], np.float32)
    n1 = one.particles_info()[1]
    one.set_grid("KillerSDF", shape)
    one.ParticleEmitter("KillerSDF", 0.5, 0.0, -0.75, seed=5)

    def emit(r, w):
        w.set_grid("KillerSDF", shape)
        w.ParticleEmitter("KillerSDF", 0.5, 0.0, -0.75, seed=5)
    abi.run_ranks(dd, emit)
    assert one.particles_info()[1] > n1 + 1000, "the emitter must add particles"
    check("ParticleEmitter")

    one.FLIP_P2G(dx, 3)
    abi.run_ranks(dd, lambda r, w: w.FLIP_P2G(dx, 3))
    for name in ("Velocity", "LiquidSDF"):
        util.compare_grids(gather_grid(dd, name), one.get_grid(name), f"dd P2G after reseed + emit: {name}", tol=0.0, check_inactive=False)
    for w in dd + [one]:
        w.close()


def test_tension_and_boundary_under_decomposition(gpu_lib):
    """Surface-tension terms and FLIPApplyBoundary on a decomposed world: both read caller-supplied grids by voxel coordinate
    (curvature, moving solid), every rank is given the same grid, nothing is exchanged. Owned results == the single world's."""
    from zeno_b200 import abi
    bounds = [(0, 2), (2, 4)]
    N, side = 128, 32
    pos, vel, dx = scenes.dam_break_points(N, seed=4, random_velocity=True, side=side)
    vel = vel * 0.2
    solid = scenes.box_solid_sdf(N, dx)
    one = abi.World(dx)
    one.set_grid("SolidSDF", solid)
    one.PrimToVDBPointDataGrid(pos, vel)
    dd = make_dd(abi, N, bounds, pos, vel, dx, solid)
    moving = scenes.sphere_sdf(centre=(17.0, 10.0, 12.0), radius=6.0, lo=(0, -8, -8), hi=(40, 32, 32), bg=3.0)
    moving["values"] = (moving["values"] * np.float32(dx)).astype(np.float32)
    moving["bg"] = np.array([3.0 * dx], np.float32)

    def prep(w):
        w.set_grid("KillerSDF", moving)
        w.FLIPApplyBoundary("KillerSDF", False)
        w.FLIP_P2G(dx, 3)
    prep(one)
    abi.run_ranks(dd, lambda r, w: prep(w))
    # the static solid of a rank holds what its particles can reach: compare where the single world's leaves overlap the slab
    ref_s = one.get_grid("SolidSDF")
    for r, w in enumerate(dd):
        lo, hi = w.dd_owned()
        mine = scenes.canonical_grid(w.get_grid("SolidSDF"))
        ref = scenes.canonical_grid(ref_s)
        idx = {tuple(o): i for i, o in enumerate(ref["origins"].tolist())}
        hit = 0
        for i, o in enumerate(mine["origins"].tolist()):
            j = idx.get(tuple(o))
            if j is None:
                continue
            hit += 1
            assert np.array_equal(mine["values"][i], ref["values"][j]), f"rank {r}: static solid leaf {o} differs"
        assert hit > 0
    # curvature = a smooth function of the voxel coordinate on the liquid SDF's leaves (any leaf set works)
    curv = one.get_grid("LiquidSDF")
    o = curv["origins"].astype(np.float32)
    curv = {k: v.copy() for k, v in curv.items()}
    wave = np.sin(0.05 * o[:, 0]).reshape((-1,) + (1,) * (curv["values"].ndim - 1))
    curv["values"] = (wave * np.ones_like(curv["values"]) * np.float32(3.0)).astype(np.float32)
    curv["bg"] = np.array([0.0], np.float32)

    def solve(w):
        w.set_grid("Curvature", curv)
        w.set_surface_tension(1000.0, 0.07)
        w.CutCellWeight()
        w.PushOutLiquidSDF(dx)
        w.FieldAddVector(G[0] * DT, G[1] * DT, G[2] * DT)
        res = w.AssembleSolvePPE(DT, dx)
        w.SubtractPressureGradient(DT, dx, 3)
        return res
    r1 = solve(one)
    rd = abi.run_ranks(dd, lambda r, w: solve(w))
    assert r1["status"] == 0 and all(r["status"] == 0 for r in rd)
    util.compare_grids(gather_grid(dd, "Divergence"), one.get_grid("Divergence"), "dd tension right-hand side", tol=0.0, check_inactive=False)
    util.compare_grids(gather_grid(dd, "Pressure"), one.get_grid("Pressure"), "dd tension Pressure", tol=1e-5, check_inactive=False)
    util.compare_grids(gather_grid(dd, "Velocity"), one.get_grid("Velocity"), "dd tension projected Velocity", tol=1e-5, check_inactive=False)
    for w in dd + [one]:
        w.close()
