"""The DEVICE code of the kernels added after the round's GPU budget was spent (zeno_b200/csrc/next_kernels.cuh: KillParticlesInSDF,
ParticleAddDV, VDBRenormalizeSDF, VDBErodeSDF) executed on the CPU and compared with the oracle bit for bit.

tests/emul/emul_next_kernels.cpp compiles the per-thread kernel bodies with plain g++ (CUDA round-to-nearest intrinsics -> the IEEE
host operations, no contraction) and loops over every (leaf, thread) of a launch; this test builds the device data layout (dense leaf
directory, slot-ordered [leaf][512] arrays, packed particle words, global voxel prefix) with numpy, replays the host-side launch
sequences, and checks the results against the oracle, which is itself pinned to the reference's node classes
(tests/test_ref_pin_cpu.py). Not covered here: launch configuration, shared-memory staging, stream ordering -- tests/test_zz_next_gpu.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import util
from zeno_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    out = os.path.join(ROOT, "tests", "emul", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemul_next.so")
    src = os.path.join(ROOT, "tests", "emul", "emul_next_kernels.cpp")
    hdr = os.path.join(ROOT, "zeno_b200", "csrc", "next_kernels.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-msse2", "-mfpmath=sse", "-fPIC", "-shared", "-w",
                               "-I" + CUDA_INC, src, "-o", so])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Topo:
    """the device topology of a leaf set: slots in lexicographic (x,y,z) order + dense directory over the bounding box"""

    def __init__(self, origins):
        o = np.unique(np.asarray(origins, np.int32).reshape(-1, 3), axis=0)
        order = np.lexsort((o[:, 2], o[:, 1], o[:, 0]))
        self.origins = np.ascontiguousarray(o[order])
        self.n = self.origins.shape[0]
        lc = self.origins >> 3
        self.dmin = np.ascontiguousarray(lc.min(axis=0).astype(np.int32))
        self.ddim = np.ascontiguousarray((lc.max(axis=0) - lc.min(axis=0) + 1).astype(np.int32))
        self.dir = np.full(int(self.ddim.prod()), -1, np.int32)
        rel = lc - self.dmin
        self.dir[(rel[:, 0] * self.ddim[1] + rel[:, 1]) * self.ddim[2] + rel[:, 2]] = np.arange(self.n, dtype=np.int32)
        self.key = {tuple(x): i for i, x in enumerate(self.origins.tolist())}

    def slots(self, origins):
        return np.array([self.key[tuple(x)] for x in np.asarray(origins).tolist()], np.int64)

    def args(self):
        return [C.c_int(self.n), _p(self.dmin), _p(self.ddim), _p(self.dir), _p(self.origins)]


def grid_to_device(g):
    t = Topo(g["origins"])
    s = t.slots(g["origins"])
    val = np.full((t.n, 512), g["bg"][0], np.float32)
    mask = np.zeros((t.n, 8), np.uint64)
    val[s] = g["values"][:, 0]
    mask[s] = g["masks"]
    return t, np.ascontiguousarray(val), np.ascontiguousarray(mask)


def grid_from_device(t, val, mask, bg):
    return {"origins": t.origins.copy(), "masks": mask.copy(), "values": val.reshape(t.n, 1, 512).copy(), "bg": np.asarray(bg, np.float32)}


def particles_to_device(p):
    """packed words + global per-voxel exclusive prefix in slot order (common.cuh 'Particles')"""
    t = Topo(p["origins"])
    s = t.slots(p["origins"])
    ve = p["voxel_end"].astype(np.int64)
    counts_in = np.diff(np.concatenate([np.zeros((ve.shape[0], 1), np.int64), ve], axis=1), axis=1)
    counts = np.zeros((t.n, 512), np.int64)
    counts[s] = counts_in
    start = np.concatenate([[0], np.cumsum(counts.ravel())]).astype(np.uint32)
    leaf_begin_in = np.concatenate([[0], np.cumsum(ve[:, -1])])
    P, v = p["P"].astype(np.uint32), p["v"].astype(np.uint32)
    n = P.shape[0]
    w0, w1, w2 = (np.zeros(n, np.uint32) for _ in range(3))
    for i, slot in enumerate(s):          # a leaf's particles move as a block
        a, b = leaf_begin_in[i], leaf_begin_in[i + 1]
        d = int(start[slot * 512])
        w0[d:d + b - a] = P[a:b, 0] | (P[a:b, 1] << 16)
        w1[d:d + b - a] = P[a:b, 2] | (v[a:b, 0] << 16)
        w2[d:d + b - a] = v[a:b, 1] | (v[a:b, 2] << 16)
    return t, start, w0, w1, w2


def particles_from_device(t, start, w0, w1, w2):
    counts = np.diff(start.astype(np.int64)).reshape(t.n, 512)
    P = np.stack([w0 & 0xffff, w0 >> 16, w1 & 0xffff], axis=1).astype(np.uint16)
    v = np.stack([w1 >> 16, w2 & 0xffff, w2 >> 16], axis=1).astype(np.uint16)
    return {"origins": t.origins.copy(), "voxel_end": np.cumsum(counts, axis=1).astype(np.uint32), "P": P, "v": v}


@pytest.fixture(scope="module")
def scene(oracle_lib):
    from oracle.pyoracle import OracleWorld
    pos, vel, dx = scenes.dam_break_points(32, seed=8, random_velocity=True)
    w = OracleWorld(dx)
    w.PrimToVDBPointDataGrid(pos, vel)
    w.FLIP_P2G(dx, 3)
    return {"dx": dx, "particles": w.get_particles(), "sdf": w.get_grid("LiquidSDF"), "pos": pos, "vel": vel}


def test_renormalize_and_erode_device_code(emul, scene, oracle_lib):
    from oracle.pyoracle import OracleWorld
    dx = scene["dx"]
    ow = OracleWorld(dx)
    ow.set_grid("LiquidSDF", scene["sdf"])
    ow.VDBRenormalizeSDF("LiquidSDF", 4, 0)
    ow.VDBErodeSDF("LiquidSDF", 0.37 * dx)
    # host sequence of renormalize_sdf (stencils.cu): phi0 = copy; three stages; swap
    t, val, mask = grid_to_device(scene["sdf"])
    bg = np.float32(scene["sdf"]["bg"][0])
    h = np.float32(dx)
    dt, inv = np.float32(h * np.float32(1.0)), np.float32(np.float32(1.0) / h)
    a = np.empty_like(val)
    for _ in range(4):
        phi0 = val.copy()
        for cur, out, N, D in ((val, a, 0, 1), (a, val, 3, 4), (val, a, 1, 3)):
            alpha = np.float32(N) / np.float32(D) if N else np.float32(0.0)
            emul.emul_renorm_stage(*t.args(), _p(mask), _p(cur), _p(phi0), _p(out), C.c_float(bg), C.c_float(dt), C.c_float(inv),
                                   C.c_float(alpha), C.c_float(np.float32(1.0) - alpha), C.c_int(1 if N else 0))
        val, a = a, val
    emul.emul_add_active(C.c_int(t.n), _p(mask), _p(val), C.c_float(np.float32(0.37 * dx)))
    util.compare_grids(grid_from_device(t, val, mask, [bg]), ow.get_grid("LiquidSDF"), "device code of renorm_stage + add_active vs oracle", tol=0.0)


def test_add_dv_device_code(emul, scene, oracle_lib):
    from oracle.pyoracle import OracleWorld
    ow = OracleWorld(scene["dx"])
    ow.set_particles(scene["particles"])
    ow.ParticleAddDV(0.013, -0.1633333, 1.0e-4)
    t, start, w0, w1, w2 = particles_to_device(scene["particles"])
    emul.emul_add_dv(_p(w1), _p(w2), C.c_uint64(w0.shape[0]), C.c_double(float(np.float32(0.013))), C.c_double(float(np.float32(-0.1633333))),
                     C.c_double(float(np.float32(1.0e-4))))
    util.compare_particles(particles_from_device(t, start, w0, w1, w2), ow.get_particles(), "device code of add_dv vs oracle")


@pytest.mark.parametrize("keep", [True, False], ids=["KEEP", "DEL"])
def test_kill_keys_device_code(emul, scene, oracle_lib, keep):
    from oracle.pyoracle import OracleWorld
    killer = scenes.sphere_sdf(centre=(3.3, 4.1, 2.7), radius=4.6, lo=(-8, -8, -8), hi=(16, 16, 16), bg=3.0)
    ow = OracleWorld(scene["dx"])
    ow.set_particles(scene["particles"])
    ow.set_grid("KillerSDF", killer)
    ow.KillParticlesInSDF("KillerSDF", keep)
    t, start, w0, w1, w2 = particles_to_device(scene["particles"])
    st, sval, _ = grid_to_device(killer)
    keys = np.zeros(w0.shape[0], np.uint32)
    emul.emul_kill_keys(*t.args(), _p(start), _p(w0), _p(w1), *st.args(), _p(sval), C.c_float(killer["bg"][0]), C.c_int(1 if keep else 0), _p(keys))
    # sort_by_key with cap 0 (particles.cu) on keys that are already in store order = a stable compaction of the survivors
    alive = keys != 0xFFFFFFFF
    assert 0 < alive.sum() < keys.shape[0]
    assert np.all(np.diff(keys[alive].astype(np.int64)) >= 0), "surviving keys must stay sorted (each is the particle's own voxel)"
    counts = np.bincount(keys[alive].astype(np.int64), minlength=t.n * 512)
    new_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    got = particles_from_device(t, new_start, w0[alive], w1[alive], w2[alive])
    util.compare_particles(got, ow.get_particles(), "device code of kill_keys vs oracle")


def test_box_avg_device_code(emul, scene, oracle_lib):
    from oracle.pyoracle import OracleWorld
    ow = OracleWorld(scene["dx"])
    ow.set_grid("LiquidSDF", scene["sdf"])
    ow.VDBSmoothSDF("LiquidSDF", 2, 2)
    # host sequence of smooth_sdf (stencils.cu): per iteration 4 x (X, Z, Y), ping-pong
    t, val, mask = grid_to_device(scene["sdf"])
    bg = np.float32(scene["sdf"]["bg"][0])
    a = np.empty_like(val)
    w, frac = 2, np.float32(1.0) / np.float32(5.0)
    for _ in range(2):
        for _rep in range(4):
            for axis in (0, 2, 1):
                emul.emul_box_avg(*t.args(), _p(mask), _p(val), _p(a), C.c_float(bg), C.c_int(axis), C.c_int(w), C.c_float(frac))
                val, a = a, val
    util.compare_grids(grid_from_device(t, val, mask, [bg]), ow.get_grid("LiquidSDF"), "device code of box_avg vs oracle", tol=0.0)
