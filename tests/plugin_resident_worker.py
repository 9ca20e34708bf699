"""Worker of tests/test_plugin_cpu.py::test_plugin_resident_mode: two free-running substeps through the plugin's node
classes (oracle behind the C ABI) and a digest of the final state. FLIPB200_RESIDENT is read once per process by the
plugin, hence the separate process."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(world_cls):
    from tests import util
    from zeno_b200 import scenes
    N, dt = 32, 0.006
    pos, vel, dx = scenes.dam_break_points(N, seed=5, random_velocity=True)
    vel *= np.float32(0.25)
    w = world_cls(dx)
    w.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
    w.PrimToVDBPointDataGrid(pos, vel)
    w.FLIP_P2G(dx, 3)
    for _ in range(2):
        for name in ("faceweight", "pushout", "addvec", "ppe", "grad", "g2p", "p2g1"):
            util.run_ref_stage(w, name, dx, dt)
    h = hashlib.sha256()
    h.update(scenes.canonical_particles(w.get_particles()).tobytes())
    for g in ("Velocity", "LiquidSDF", "Pressure"):
        c = scenes.canonical_grid(w.get_grid(g))
        mb = scenes.mask_bits(c["masks"])
        h.update(c["origins"].tobytes()); h.update(c["masks"].tobytes())
        for ch in range(c["values"].shape[1]):
            h.update(np.where(mb, c["values"][:, ch], 0).astype(np.float32).tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    from oracle import pyoracle
    cls = pyoracle.OracleWorld if sys.argv[1:] == ["oracle"] else pyoracle.PluginWorld
    print("DIGEST", run(cls))
