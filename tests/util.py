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


# ------------------------------------------------------------------------------------------------
# replay of the REAL-reference fixtures (tests/golden/ref_*.npz, made by tests/golden/make_ref_golden.py)
import math
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_STAGES = [("p2g0", ("Velocity", "PostAdvVelocity", "LiquidSDF")), ("faceweight", ("CellFWeight",)),
              ("pushout", ("LiquidSDF",)), ("addvec", ("Velocity",)), ("ppe", ("Pressure", "Divergence")),
              ("grad", ("Velocity",)), ("g2p", ()), ("p2g1", ("Velocity", "PostAdvVelocity", "LiquidSDF"))]
# relative-L2 bars against the real reference. The reference binary is built with -mfma and GCC's default
# contraction (FF/CMakeLists.txt:56), the oracle and the CUDA path execute the un-contracted op sequence,
# so fp32 values agree to rounding, not bitwise. North-star bar: 1e-5 for transferred velocities.
REF_TOL = {"p2g0": 1e-5, "faceweight": 1e-6, "pushout": 1e-6, "addvec": 1e-6, "ppe": 1e-4, "grad": 1e-5, "p2g1": 1e-5}
G_REF = (0.0, -9.8, 0.0)


def fixture_grid(fx, key):
    return {k: fx[f"{key}.{k}"] for k in ("origins", "masks", "values", "bg")}


def fixture_particles(fx, key):
    return {k: fx[f"{key}.{k}"] for k in ("origins", "voxel_end", "P", "v")}


def run_ref_stage(w, name, dx, dt):
    if name in ("p2g0", "p2g1"):
        w.FLIP_P2G(dx, 3)
    elif name == "faceweight":
        w.CutCellWeight()
    elif name == "pushout":
        w.PushOutLiquidSDF(dx)
    elif name == "addvec":
        w.FieldAddVector(G_REF[0] * dt, G_REF[1] * dt, G_REF[2] * dt)
    elif name == "ppe":
        return w.AssembleSolvePPE(dt, dx)
    elif name == "grad":
        w.SubtractPressureGradient(dt, dx, 3)
    elif name == "g2p":
        w.G2PAdvectorSheetty(dt, dx, 4, 3, 0.03, 0.05, True)
    return None


def particle_code_report(a: np.ndarray, b: np.ndarray):
    """a, b: canonical particle rows [n,9] (voxel xyz, P codes, v half bits) of two implementations of
    the same advect. Pairs particles inside each voxel by sorted order and reports how far the stored
    codes are apart (SURVEY 8d codec caveat: a 1e-7 fp32 difference can move a code by one LSB)."""
    same_voxel = (a[:, :3] == b[:, :3]).all(axis=1)
    dP = np.abs(a[:, 3:6] - b[:, 3:6])[same_voxel]

    def half_to_f(h):
        return h.astype(np.uint16).view(np.float16).astype(np.float64)
    va, vb = half_to_f(a[:, 6:9].astype(np.uint16)), half_to_f(b[:, 6:9].astype(np.uint16))
    dv = np.abs(va - vb)[same_voxel]
    scale = np.maximum(np.abs(vb[same_voxel]), 1e-3)
    return {"same_voxel": float(same_voxel.mean()), "identical": float((a == b).all(axis=1).mean()),
            "P_within_1lsb": float((dP <= 1).all(axis=1).mean()) if dP.size else 1.0,
            "P_max_lsb": int(dP.max()) if dP.size else 0,
            "v_within_1ulp": float((dv <= scale * 2.0 ** -10).all(axis=1).mean()) if dv.size else 1.0}


def replay_ref_chain(fx, world_cls, stages=None, report=None):
    """One-step-synchronised replay: every node starts from the REFERENCE's state of that stage and its
    outputs are compared with the reference's outputs. Returns {stage: metrics}."""
    N, seed = int(fx["meta.N"]), int(fx["meta.seed"])
    dt = float(fx["meta.dt"])
    pos, vel, dx = scenes.dam_break_points(N, seed=seed, random_velocity=True)
    vel = vel * np.float32(fx["meta.vscale"])
    solid = scenes.box_solid_sdf(N, dx)
    report = {} if report is None else report
    have = {k.split(".")[0] for k in fx.keys()}
    w = world_cls(dx)
    w.set_grid("SolidSDF", solid)
    if "bin" in have:
        w.PrimToVDBPointDataGrid(pos, vel)
        compare_particles(w.get_particles(), fixture_particles(fx, "bin.particles"), "bin vs reference")
        report["bin"] = {"particles": int(fx["bin.particles.P"].shape[0])}
    state, pkey = {}, "bin.particles"
    for name, grids in REF_STAGES:
        present = name in have
        if present and (stages is None or name in stages):
            for g, key in state.items():
                w.set_grid(g, fixture_grid(fx, key))
            if f"{pkey}.P" in fx:
                w.set_particles(fixture_particles(fx, pkey))
            r = run_ref_stage(w, name, dx, dt)
            m = {}
            for g in grids:
                m[g] = compare_grids(w.get_grid(g), fixture_grid(fx, f"{name}.{g}"), f"{name}.{g} vs reference",
                                     tol=REF_TOL[name], check_inactive=False)
            if name == "ppe":
                info = w.solver_info()
                ref_it = int(fx["ppe.iterations"])
                assert r["status"] == int(fx["ppe.status"]) == 0, r
                assert info["levels"] == int(fx["ppe.levels"]), (info["levels"], int(fx["ppe.levels"]))
                assert info["num_dof"] == int(fx["ppe.num_dof"]), (info["num_dof"], int(fx["ppe.num_dof"]))
                assert r["iterations"] <= math.ceil(1.1 * ref_it), (r["iterations"], ref_it)
                assert r["rel_residual"] <= 5e-5
                m["iterations"] = (r["iterations"], ref_it)
            if name == "grad" and "cfl.dt" in fx:
                c = w.CFL_dt()
                assert abs(c - float(fx["cfl.dt"])) <= 1e-5 * float(fx["cfl.dt"]), (c, float(fx["cfl.dt"]))
                m["cfl"] = c
            if name == "g2p":
                a = scenes.canonical_particles(w.get_particles())
                b = scenes.canonical_particles(fixture_particles(fx, "g2p.particles"))
                assert a.shape == b.shape, (a.shape, b.shape)
                assert w.dropped() == int(fx["g2p.dropped"])
                m.update(particle_code_report(a, b))
                assert m["same_voxel"] >= 0.999 and m["P_within_1lsb"] >= 0.999 and m["v_within_1ulp"] >= 0.999, m
            report[name] = m
        if present:
            for g in grids:
                state[g] = f"{name}.{g}"
            if name == "g2p":
                pkey = "g2p.particles"
    w.close()
    return report
