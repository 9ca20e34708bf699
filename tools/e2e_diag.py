"""Where the end-to-end (host buffers through the C ABI) step spends its time: per-call wall clock."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zeno_b200 import abi, scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pos, vel, dx = scenes.dam_break_points(N, seed=1)
solid = scenes.box_solid_sdf(N, dx)
w = abi.World(dx)
w.set_grid("SolidSDF", solid)
w.PrimToVDBPointDataGrid(pos, vel)
w.FLIP_P2G(dx, 3)
GR = ("Velocity", "PostAdvVelocity", "LiquidSDF")
for _ in range(3):
    w.substep(0.004, dx, 4, 3, 0.03, 0.05, (0, -9.8, 0), 3, True)
arena = abi.PinnedArena()
state = {g: {k: (arena.like(v) if k != "bg" else v) for k, v in w.get_grid(g).items()} for g in GR}
pts = {k: arena.like(v) for k, v in w.get_particles().items()}
out_state = {g: {k: (arena.like(v, 1.5) if k != "bg" else v) for k, v in state[g].items()} for g in GR}
out_pts = {k: arena.like(v, 1.5 if k in ("origins", "voxel_end") else 1.0) for k, v in pts.items()}
for it in range(4):
    t = [time.perf_counter()]
    for g in GR:
        w.set_grid(g, state[g]); t.append(time.perf_counter())
    w.set_particles(pts); t.append(time.perf_counter())
    w.substep(0.004, dx, 4, 3, 0.03, 0.05, (0, -9.8, 0), 3, True); t.append(time.perf_counter())
    p = w.get_particles(out=out_pts); t.append(time.perf_counter())
    for g in GR:
        w.get_grid(g, out=out_state[g]); t.append(time.perf_counter())
    d = np.diff(np.array(t)) * 1e3
    print("iter", it, "set_grid x3 %s  set_particles %.2f  substep %.2f  get_particles %.2f  get_grid x3 %s  total %.2f ms" % (
        np.round(d[:3], 2), d[3], d[4], d[5], np.round(d[6:9], 2), d.sum()),
        "pinned-out used:", p["P"].ctypes.data == out_pts["P"].ctypes.data, flush=True)

# the asynchronous sequence bench.py times (downloads begin as soon as a result is final)
GRAVITY = (0.0, -9.8, 0.0)
for it in range(8):
    skip_p = it >= 4   # second half: no particle download, to see what the concurrent copy costs the kernels
    names, t = [], [time.perf_counter()]
    def lap(n):
        names.append(n); t.append(time.perf_counter())
    for g in GR:
        w.set_grid(g, state[g])
    lap("set_grid x3")
    w.set_particles(pts); lap("set_particles")
    dt = float(min(3.0 * w.CFL_dt(), 1.0 / 24.0)); lap("CFL_dt")
    w.G2PAdvectorSheetty(dt, dx, 4, 3, 0.03, 0.05, True); lap("G2P")
    if not skip_p:
        out_p = w.get_particles_begin(out_pts)
    lap("particles_begin")
    w.FLIP_P2G(dx, 3); lap("P2G")
    w.CutCellWeight(); w.PushOutLiquidSDF(dx); lap("weights+pushout")
    og = {g: w.get_grid_begin(g, out_state[g]) for g in ("PostAdvVelocity", "LiquidSDF")}; lap("grids_begin x2")
    w.FieldAddVector(GRAVITY[0] * dt, GRAVITY[1] * dt, GRAVITY[2] * dt); lap("add_vector")
    w.AssembleSolvePPE(dt, dx); lap("solve")
    w.SubtractPressureGradient(dt, dx, 3); lap("gradient")
    og["Velocity"] = w.get_grid_begin("Velocity", out_state["Velocity"]); lap("vel_begin")
    w.download_wait(); lap("download_wait")
    d = np.diff(np.array(t)) * 1e3
    print("async iter", it, "  ".join("%s %.2f" % (n, x) for n, x in zip(names, d)), " total %.2f ms" % d.sum(), flush=True)
