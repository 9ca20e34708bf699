"""Regenerates tests/golden/ref_*.npz from the REAL reference (oracle/_ref/libflipref.so =
projects/FastFLIP/{FLIP_vdb,simd_vdb_poisson_uaamg,vdb_velocity_extrapolator,levelset_util}.cpp +
OpenVDB 9.0.1 + TBB, built by oracle/ref/build_ref.sh). Run in the container that has /root/reference:

    bash oracle/ref/build_ref.sh && python tests/golden/make_ref_golden.py

Each fixture holds the state the reference's node chain produces after EVERY node for one substep of
a seeded dam break (inputs are regenerated from zeno_b200/scenes.py, so only outputs are stored):
    bin -> FLIP_P2G -> CutCellWeight -> PushOutLiquidSDF -> FieldAddVector -> AssembleSolvePPE ->
    SubtractPressureGradient -> CFL_dt -> G2PAdvectorSheetty(RK3) -> FLIP_P2G
tests/test_ref_pin_cpu.py replays the chain one-step-synchronised through the oracle restatement and
tests/test_parity_gpu.py through the CUDA library; both compare against these reference outputs.
The reference itself is not bit-reproducible run to run in two places (TBB reduction order in the
solver's dot products; within-voxel particle order), which is why values carry tolerances.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import RefWorld, ref_set_threads  # noqa: E402
from zeno_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
G = (0.0, -9.8, 0.0)

# name -> (N, seed, velocity scale, dt)
CASES = {"ref_chain48": (48, 5, 0.3, 0.008), "ref_chain96": (96, 2, 0.1, 0.004)}
# ref_chain96 exists for the 2-level multigrid solve: only the solver's inputs and outputs are kept
KEEP = {"ref_chain96": ("faceweight.", "pushout.", "addvec.", "ppe.", "grad.", "cfl.", "meta.")}
# stage name -> grids written by the stage
STAGES = [("p2g0", ("Velocity", "PostAdvVelocity", "LiquidSDF")), ("faceweight", ("CellFWeight",)),
          ("pushout", ("LiquidSDF",)), ("addvec", ("Velocity",)), ("ppe", ("Pressure", "Divergence")),
          ("grad", ("Velocity",)), ("g2p", ()), ("p2g1", ("Velocity", "PostAdvVelocity", "LiquidSDF"))]


def put_grid(out, key, g):
    c = scenes.canonical_grid(g, drop_empty=False)
    for k in ("origins", "masks", "values", "bg"):
        out[f"{key}.{k}"] = c[k]


def put_particles(out, key, p):
    for k in ("origins", "voxel_end", "P", "v"):
        out[f"{key}.{k}"] = p[k]


def run_stage(w, name, dx, dt):
    """Executes the node of stage `name` on world w (any of RefWorld / OracleWorld / abi.World)."""
    if name in ("p2g0", "p2g1"):
        w.FLIP_P2G(dx, 3)
    elif name == "faceweight":
        w.CutCellWeight()
    elif name == "pushout":
        w.PushOutLiquidSDF(dx)
    elif name == "addvec":
        w.FieldAddVector(G[0] * dt, G[1] * dt, G[2] * dt)
    elif name == "ppe":
        return w.AssembleSolvePPE(dt, dx)
    elif name == "grad":
        w.SubtractPressureGradient(dt, dx, 3)
    elif name == "g2p":
        w.G2PAdvectorSheetty(dt, dx, 4, 3, 0.03, 0.05, True)
    else:
        raise KeyError(name)
    return None


def make(case):
    N, seed, vscale, dt = CASES[case]
    pos, vel, dx = scenes.dam_break_points(N, seed=seed, random_velocity=True)
    vel = vel * np.float32(vscale)
    w = RefWorld(dx)
    w.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
    w.PrimToVDBPointDataGrid(pos, vel)
    out = {"meta.N": np.int64(N), "meta.seed": np.int64(seed), "meta.vscale": np.float32(vscale), "meta.dt": np.float32(dt)}
    put_particles(out, "bin.particles", w.get_particles())
    for name, grids in STAGES:
        r = run_stage(w, name, dx, dt)
        for g in grids:
            put_grid(out, f"{name}.{g}", w.get_grid(g))
        if name == "ppe":
            info = w.solver_info()
            out["ppe.iterations"] = np.int64(r["iterations"])
            out["ppe.status"] = np.int64(r["status"])
            out["ppe.levels"] = np.int64(info["levels"])
            out["ppe.num_dof"] = np.int64(info["num_dof"])
            out["ppe.history"] = info["history"]
        if name == "grad":
            out["cfl.dt"] = np.float32(w.CFL_dt())
        if name == "g2p":
            put_particles(out, "g2p.particles", w.get_particles())
            out["g2p.dropped"] = np.int64(w.dropped())
    if case in KEEP:
        out = {k: v for k, v in out.items() if k.startswith(KEEP[case])}
    path = os.path.join(HERE, case + ".npz")
    np.savez_compressed(path, **out)
    print(case, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "PCG iterations", int(out["ppe.iterations"]),
          "levels", int(out["ppe.levels"]), "dof", int(out["ppe.num_dof"]))


if __name__ == "__main__":
    print("reference threads:", ref_set_threads(0))
    for c in (sys.argv[1:] or list(CASES)):
        make(c)
