"""Per-node wall clock of one substep through the drop-in's node classes on real OpenVDB objects (oracle/_ref/libflipplugin_gpu.so)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zeno_b200 import scenes
from oracle import pyoracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pos, vel, dx = scenes.dam_break_points(N, seed=1)
pw = pyoracle.PluginGpuWorld(dx)
pw.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
pw.PrimToVDBPointDataGrid(pos, vel)
pw.FLIP_P2G(dx, 3)
G = (0.0, -9.8, 0.0)
for it in range(3):
    names, t = [], [time.perf_counter()]
    def lap(n):
        names.append(n); t.append(time.perf_counter())
    dt = float(min(3.0 * pw.CFL_dt(), 1.0 / 24.0)); lap("CFL_dt")
    pw.G2PAdvectorSheetty(dt, dx, 4, 3, 0.03, 0.05, True); lap("G2P")
    pw.FLIP_P2G(dx, 3); lap("P2G")
    pw.CutCellWeight(); lap("CutCellWeight")
    pw.PushOutLiquidSDF(dx); lap("PushOut")
    pw.FieldAddVector(G[0] * dt, G[1] * dt, G[2] * dt); lap("AddVector")
    pw.AssembleSolvePPE(dt, dx); lap("SolvePPE")
    pw.SubtractPressureGradient(dt, dx, 3); lap("Gradient")
    d = np.diff(np.array(t)) * 1e3
    print("nodes iter", it, "  ".join("%s %.2f" % (n, x) for n, x in zip(names, d)), " total %.2f ms" % d.sum(), flush=True)
