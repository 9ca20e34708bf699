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
