"""Regenerates tests/golden/*.npz.

Provenance: the reference ships no tests or golden vectors for FastFLIP, so
these fixtures are produced by the CPU oracle (oracle/*.cpp, a restatement of the reference
algorithms) on seeded inputs; hand-computable cases are additionally asserted analytically in
tests/test_oracle_cpu.py. When oracle/_ref (the real reference sources compiled by
oracle/ref/build_ref.sh) is available, tests/test_ref_pin.py checks the oracle against it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import OracleWorld, load  # noqa: E402
from zeno_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DT = 0.01
G = (0.0, -9.8, 0.0)


def flat(prefix, g):
    c = scenes.canonical_grid(g)
    return {f"{prefix}_origins": c["origins"], f"{prefix}_masks": c["masks"], f"{prefix}_values": c["values"], f"{prefix}_bg": c["bg"]}


def main():
    lib = load()
    import ctypes as C
    # (i) codec tables
    u = np.arange(65536, dtype=np.uint16)
    fx_dec = np.array([lib.orc_fxpt16_decode(C.c_uint16(int(x))) for x in u], np.float32)
    probe = np.concatenate([np.linspace(-0.6, 0.6, 4001, dtype=np.float32), fx_dec[::97]])
    fx_enc = np.array([lib.orc_fxpt16_encode(C.c_float(float(x))) for x in probe], np.uint16)
    hp = np.concatenate([np.float32([0, -0.0, 1, -1, 65504, 65520, 1e-8, 6e-8, 6.1e-5, 1e5, -1e5, 0.1, 1 / 3]),
                         np.random.default_rng(7).normal(0, 3, 2000).astype(np.float32)])
    h_enc = np.array([lib.orc_half_encode(C.c_float(float(x))) for x in hp], np.uint16)
    np.savez_compressed(os.path.join(HERE, "codecs.npz"), fx_dec=fx_dec, fx_probe=probe, fx_enc=fx_enc, half_probe=hp, half_enc=h_enc)

    # (ii) single-particle P2G at a leaf corner, (iv) dilation of one voxel at (7,7,7)
    dx = 0.1
    w = OracleWorld(dx)
    pos = np.float32([[7.2 * dx, 7.1 * dx, 6.9 * dx]])
    vel = np.float32([[1.0, -2.0, 0.5]])
    w.PrimToVDBPointDataGrid(pos, vel)
    w.FLIP_P2G(dx, 3)
    d = {"pos": pos, "vel": vel, "dx": np.float32(dx)}
    d.update(flat("vel", w.get_grid("Velocity")))
    d.update(flat("post", w.get_grid("PostAdvVelocity")))
    d.update(flat("sdf", w.get_grid("LiquidSDF")))
    np.savez_compressed(os.path.join(HERE, "single_particle_p2g.npz"), **d)

    # (vii) a tiny dam break, two substeps
    N = 16
    pos, vel, dx = scenes.dam_break_points(N, seed=2, random_velocity=True)
    vel *= 0.1
    w = OracleWorld(dx)
    w.set_grid("SolidSDF", scenes.box_solid_sdf(N, dx))
    w.PrimToVDBPointDataGrid(pos, vel)
    w.FLIP_P2G(dx, 3)
    d = {"N": np.int32(N), "dx": np.float32(dx), "pos": pos, "vel": vel, "dt": np.float32(DT)}
    its = []
    for step in range(2):
        w.CutCellWeight()
        w.PushOutLiquidSDF(dx)
        w.FieldAddVector(G[0] * DT, G[1] * DT, G[2] * DT)
        its.append(w.AssembleSolvePPE(DT, dx)["iterations"])
        w.SubtractPressureGradient(DT, dx, 3)
        w.G2PAdvectorSheetty(DT, dx, 4, 3, 0.03, 0.05, True)
        w.FLIP_P2G(dx, 3)
        d[f"particles_{step}"] = scenes.canonical_particles(w.get_particles())
        d.update(flat(f"vel_{step}", w.get_grid("Velocity")))
        d.update(flat(f"sdf_{step}", w.get_grid("LiquidSDF")))
    d["pcg_iterations"] = np.int32(its)
    np.savez_compressed(os.path.join(HERE, "dambreak16.npz"), **d)
    print("golden written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
