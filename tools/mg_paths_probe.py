"""Debug aid: solve the same PPE through every preconditioner execution path and print the residual histories."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_mg_paths_gpu import _world, _solve

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w, dx = _world(N)
runs = [("cycle", {"FLIPB200_MG_PATH": "cycle"}), ("cycle2", {"FLIPB200_MG_PATH": "cycle"}),
        ("perpass", {"FLIPB200_MG_PATH": "cycle", "FLIPB200_NO_CYCLE_KERNEL": "1"}),
        ("tiles", {"FLIPB200_MG_PATH": "tiles"}), ("tiles2", {"FLIPB200_MG_PATH": "tiles"})]
for extra in sys.argv[2:]:
    k, v = extra.split("=")
    runs.append((extra, {"FLIPB200_MG_PATH": "tiles", k: v}))
out = {}
for name, env in runs:
    os.environ.pop("FLIPB200_NO_CYCLE_KERNEL", None)
    res, hist, p = _solve(w, dx, env, tol=1e-6)
    os.environ.pop("FLIPB200_NO_CYCLE_KERNEL", None)
    out[name] = (hist, p)
    print(name, res, [float.hex(float(h)) for h in hist[:4]])
base = out["perpass"]
for name, (hist, p) in out.items():
    same = np.array_equal(hist, base[0])
    dv = np.abs(p["values"] - base[1]["values"]).max() if p["values"].shape == base[1]["values"].shape else -1
    print(f"{name:10s} history==perpass {same}  max |dp| {dv:.3e}")
