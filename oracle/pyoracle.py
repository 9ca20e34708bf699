"""TEST INFRASTRUCTURE ONLY -- ctypes driver of oracle/_build/liboracle.so (the CPU restatement of
the reference's FastFLIP hot path). Imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, as the checker; the product (zeno_b200/) never
imports it. Same method names and exchange formats as zeno_b200.abi.World."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
GRID_IDS = {
    "Velocity": 0, "PostAdvVelocity": 1, "ViscousVelocity": 2, "SolidVelocity": 3, "CellFWeight": 4,
    "LiquidSDF": 5, "SolidSDF": 6, "Pressure": 7, "Divergence": 8, "Curvature": 9, "KillerSDF": 10,
}
VEC_GRIDS = {"Velocity", "PostAdvVelocity", "ViscousVelocity", "SolidVelocity", "CellFWeight"}
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-j8"])
    else:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-j8"])  # incremental, no-op when up to date
    return _LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = C.CDLL(_LIB)
        lib.orc_world_create.restype = C.c_void_p
        lib.orc_cfl.restype = C.c_float
        lib.orc_dropped.restype = C.c_uint64
        lib.orc_fxpt16_encode.restype = C.c_uint16
        lib.orc_fxpt16_decode.restype = C.c_float
        lib.orc_half_encode.restype = C.c_uint16
        lib.orc_half_decode.restype = C.c_float
        lib.orc_fraction_inside2.restype = C.c_float
        lib.orc_fraction_inside4.restype = C.c_float
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class _Prefixed:
    """lib.orc_foo -> getattr(raw, prefix + 'foo'): one harness for liboracle (orc_) and libflipref (ref_)"""

    def __init__(self, raw, prefix):
        self._raw, self._prefix = raw, prefix

    def __getattr__(self, name):
        if name.startswith("orc_"):
            return getattr(self._raw, self._prefix + name[4:])
        return getattr(self._raw, name)


class OracleWorld:
    PREFIX = "orc_"

    @classmethod
    def _load(cls):
        return load()

    def __init__(self, dx: float):
        self.lib = _Prefixed(self._load(), self.PREFIX)
        self.dx = float(dx)
        self.h = C.c_void_p(self.lib.orc_world_create(C.c_float(dx)))

    def close(self):
        if self.h:
            self.lib.orc_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_grid(self, name: str, g: Dict[str, np.ndarray]):
        o = _c(g["origins"], np.int32).reshape(-1, 3)
        m = _c(g["masks"], np.uint64).reshape(-1, 8)
        v = _c(g["values"], np.float32)
        bg = _c(g["bg"], np.float32)
        rc = self.lib.orc_grid_set(self.h, C.c_int(GRID_IDS[name]), C.c_int(o.shape[0]), _p(o), _p(m), _p(v), _p(bg))
        assert rc == 0

    def get_grid(self, name: str) -> Dict[str, np.ndarray]:
        gid = GRID_IDS[name]
        nch = 3 if name in VEC_GRIDS else 1
        n = self.lib.orc_grid_leaf_count(self.h, C.c_int(gid))
        o = np.zeros((n, 3), np.int32)
        m = np.zeros((n, 8), np.uint64)
        v = np.zeros((n, nch, 512), np.float32)
        bg = np.zeros(nch, np.float32)
        assert self.lib.orc_grid_get(self.h, C.c_int(gid), _p(o), _p(m), _p(v), _p(bg)) == 0
        return {"origins": o, "masks": m, "values": v, "bg": bg}

    def set_particles(self, p: Dict[str, np.ndarray]):
        o = _c(p["origins"], np.int32).reshape(-1, 3)
        ve = _c(p["voxel_end"], np.uint32).reshape(-1, 512)
        P = _c(p["P"], np.uint16).reshape(-1, 3)
        v = _c(p["v"], np.uint16).reshape(-1, 3)
        rc = self.lib.orc_particles_set(self.h, C.c_int(o.shape[0]), _p(o), _p(ve), C.c_uint64(P.shape[0]), _p(P), _p(v))
        assert rc == 0, rc

    def particles_info(self):
        nl, n = C.c_int(0), C.c_uint64(0)
        self.lib.orc_particles_info(self.h, C.byref(nl), C.byref(n))
        return nl.value, n.value

    def get_particles(self) -> Dict[str, np.ndarray]:
        nl, n = self.particles_info()
        o = np.zeros((nl, 3), np.int32)
        ve = np.zeros((nl, 512), np.uint32)
        P = np.zeros((n, 3), np.uint16)
        v = np.zeros((n, 3), np.uint16)
        self.lib.orc_particles_get(self.h, _p(o), _p(ve), _p(P), _p(v))
        return {"origins": o, "voxel_end": ve, "P": P, "v": v}

    def PrimToVDBPointDataGrid(self, pos, vel=None):
        pos = _c(pos, np.float32).reshape(-1, 3)
        vel = None if vel is None else _c(vel, np.float32).reshape(-1, 3)
        self.lib.orc_bin_from_points(self.h, _p(pos), _p(vel), C.c_uint64(pos.shape[0]))

    def FLIP_P2G(self, dx=None, VelExtraLayer=3):
        self.lib.orc_p2g(self.h, C.c_float(self.dx if dx is None else dx), C.c_int(VelExtraLayer))

    def G2PAdvectorSheetty(self, dt, dx=None, surface_size=4, RK_ORDER=1, pic_min=0.03, pic_max=0.05, viscous_is_velocity=True):
        self.lib.orc_g2p_advect_sheetty(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(surface_size),
                                        C.c_int(RK_ORDER), C.c_float(pic_min), C.c_float(pic_max), C.c_int(1 if viscous_is_velocity else 0))

    def KillParticlesInSDF(self, sdf_grid: str = "KillerSDF", keep: bool = True):
        rc = self.lib.orc_kill_particles(self.h, C.c_int(GRID_IDS[sdf_grid]), C.c_int(1 if keep else 0))
        if rc != 0:
            raise RuntimeError("KillParticlesInSDF failed")

    def FluidReseed(self, seed: int = 0, leaf_start=None, want_leaf_end: bool = False):
        """leaf_start: optional uint64 array, the draw-sequence start of every particle leaf (store order); default = the seeded
        per-leaf hash. Returns where every leaf's sequence ended when want_leaf_end."""
        nl = self.particles_info()[0]
        ls = None if leaf_start is None else _c(leaf_start, np.uint64)
        le = np.zeros(max(nl, 1), np.uint64) if want_leaf_end else None
        rc = self.lib.orc_fluid_reseed(self.h, C.c_uint32(seed & 0xffffffff), None if ls is None else _p(ls), None if le is None else _p(le))
        if rc != 0:
            raise RuntimeError("FluidReseed failed: " + self._err())
        return le[:nl] if want_leaf_end else None

    def ParticleEmitter(self, shape_grid: str = "KillerSDF", vx=0.0, vy=0.0, vz=0.0, seed: int = 0, leaf_start=None, want_leaf_end: bool = False, max_leaves: int = 0):
        """leaf_start / the returned ends are indexed by the leaves of the RESULT; ~0 marks leaves the shape does not touch."""
        ls = None if leaf_start is None else _c(leaf_start, np.uint64)
        le = np.full(max(max_leaves, 1), np.uint64(0), np.uint64) if want_leaf_end else None
        rc = self.lib.orc_emit_liquid(self.h, C.c_int(GRID_IDS[shape_grid]), C.c_float(vx), C.c_float(vy), C.c_float(vz), C.c_uint32(seed & 0xffffffff),
                                      None if ls is None else _p(ls), None if le is None else _p(le))
        if rc != 0:
            raise RuntimeError("ParticleEmitter failed: " + self._err())
        return le[:self.particles_info()[0]] if want_leaf_end else None

    def FLIPApplyBoundary(self, moving_grid: str = "KillerSDF", moving_vertex_centred: bool = False):
        rc = self.lib.orc_apply_boundary(self.h, C.c_int(GRID_IDS[moving_grid]), C.c_int(1 if moving_vertex_centred else 0))
        if rc != 0:
            raise RuntimeError("FLIPApplyBoundary failed: " + self._err())

    def set_surface_tension(self, density: float = 1000.0, coef: float = 0.0):
        """the Density / SurfaceTension sockets of AssembleSolvePPE and SubtractPressureGradient (coef > 0 enables the tension terms)"""
        rc = self.lib.orc_set_surface_tension(self.h, C.c_float(density), C.c_float(coef))
        if rc != 0:
            raise RuntimeError("set_surface_tension failed")

    def VDBPointsToPrimitive(self):
        """world positions and velocities of all particles, store order: (pos [N,3] f32, vel [N,3] f32)"""
        n = self.particles_info()[1]
        pos, vel = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        rc = self.lib.orc_particles_to_points(self.h, _p(pos), _p(vel))
        if rc != 0:
            raise RuntimeError("VDBPointsToPrimitive failed: " + self._err())
        return pos, vel

    def _err(self):
        try:
            return (self.lib.orc_last_error() or b"").decode()
        except Exception:
            return ""

    def ParticleAddDV(self, x, y, z):
        rc = self.lib.orc_particles_add_dv(self.h, C.c_float(x), C.c_float(y), C.c_float(z))
        if rc != 0:
            raise RuntimeError("ParticleAddDV failed")

    def G2P_Advector(self, dt, dx=None, RK_ORDER=1, pic_smoothness=0.02):
        rc = self.lib.orc_g2p_advect(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(RK_ORDER), C.c_float(pic_smoothness))
        if rc != 0:
            raise RuntimeError("G2P_Advector failed")

    def VDBRenormalizeSDF(self, grid: str = "LiquidSDF", iterations: int = 4, dilateIters: int = 0):
        rc = self.lib.orc_renormalize_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_int(iterations), C.c_int(dilateIters))
        if rc != 0:
            raise RuntimeError("VDBRenormalizeSDF failed")

    def VDBErodeSDF(self, grid: str, depth: float):
        rc = self.lib.orc_erode_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_float(depth))
        if rc != 0:
            raise RuntimeError("VDBErodeSDF failed")

    def VDBSmoothSDF(self, grid: str, width: int = 1, iterations: int = 1):
        rc = self.lib.orc_smooth_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_int(width), C.c_int(iterations))
        if rc != 0:
            raise RuntimeError("VDBSmoothSDF failed")

    def dropped(self) -> int:
        return int(self.lib.orc_dropped(self.h))

    def capture_precodec(self, on: bool):
        self.lib.orc_capture_precodec(self.h, C.c_int(1 if on else 0))

    def get_precodec(self, n: int):
        pos = np.zeros((n, 3), np.float32)
        vel = np.zeros((n, 3), np.float32)
        alive = np.zeros(n, np.uint8)
        self.lib.orc_get_precodec(self.h, _p(pos), _p(vel), _p(alive))
        return pos, vel, alive

    def CutCellWeight(self):
        self.lib.orc_face_weights(self.h)

    def PushOutLiquidSDF(self, dx=None):
        self.lib.orc_pushout_sdf(self.h, C.c_float(self.dx if dx is None else dx))

    def FieldAddVector(self, x, y, z):
        self.lib.orc_add_vector(self.h, C.c_float(x), C.c_float(y), C.c_float(z))

    def CFL_dt(self) -> float:
        return float(self.lib.orc_cfl(self.h))

    def AssembleSolvePPE(self, dt, dx=None, rel_tol=None, max_iter=100):
        it, res, st = C.c_int(0), C.c_float(0), C.c_int(0)
        d = C.c_float(self.dx if dx is None else dx)
        if rel_tol is None and max_iter == 100:
            self.lib.orc_solve_ppe(self.h, C.c_float(dt), d, C.byref(it), C.byref(res), C.byref(st))
        else:
            self.lib.orc_solve_ppe_ex(self.h, C.c_float(dt), d, C.c_float(5e-5 if rel_tol is None else rel_tol), C.c_int(max_iter),
                                      C.byref(it), C.byref(res), C.byref(st))
        return {"iterations": it.value, "rel_residual": res.value, "status": st.value}

    def solver_info(self):
        lv, nd, nh = C.c_int(0), C.c_int(0), C.c_int(0)
        self.lib.orc_solver_info(self.h, C.byref(lv), C.byref(nd), C.byref(nh))
        hist = np.zeros(nh.value, np.float32)
        if nh.value:
            self.lib.orc_residual_history(self.h, _p(hist))
        return {"levels": lv.value, "num_dof": nd.value, "history": hist}

    def SubtractPressureGradient(self, dt, dx=None, VelExtraLayer=3):
        self.lib.orc_subtract_grad(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(VelExtraLayer))

    def substep(self, dt, dx=None, surface_size=4, RK_ORDER=3, pic_min=0.03, pic_max=0.05, gravity=(0.0, -9.8, 0.0),
                VelExtraLayer=3, viscous_is_velocity=True, want_stage_ms=False):
        secs = (C.c_double * 5)()
        self.lib.orc_substep(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(surface_size), C.c_int(RK_ORDER),
                             C.c_float(pic_min), C.c_float(pic_max), C.c_float(gravity[0]), C.c_float(gravity[1]), C.c_float(gravity[2]),
                             C.c_int(VelExtraLayer), C.c_int(1 if viscous_is_velocity else 0), secs)
        return [s * 1e3 for s in secs] if want_stage_ms else None


# ---------------------------------------------------------------------------------------------
# oracle/_ref: the REAL reference sources (FLIP_vdb.cpp, simd_vdb_poisson_uaamg.cpp, ...) compiled by
# oracle/ref/build_ref.sh behind the same flat API (prefix ref_). Used to pin the restatement and as the
# `--impl reference` CPU arm of bench.py.
_REF_LIB = os.path.join(_HERE, "_ref", "libflipref.so")
_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_LIB)


def load_ref() -> C.CDLL:
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_LIB):
            raise FileNotFoundError(f"{_REF_LIB} not built (run oracle/ref/build_ref.sh where /root/reference exists)")
        lib = C.CDLL(_REF_LIB)
        lib.ref_world_create.restype = C.c_void_p
        lib.ref_cfl.restype = C.c_float
        lib.ref_dropped.restype = C.c_uint64
        lib.ref_fraction_inside2.restype = C.c_float
        lib.ref_fraction_inside4.restype = C.c_float
        lib.ref_build_info.restype = C.c_char_p
        lib.ref_set_threads.restype = C.c_int
        _ref = lib
    return _ref


def ref_set_threads(n: int = 0) -> int:
    return int(load_ref().ref_set_threads(C.c_int(n)))


class RefWorld(OracleWorld):
    """Same interface as OracleWorld, executed by the reference's own code."""
    PREFIX = "ref_"

    @classmethod
    def _load(cls):
        return load_ref()


class PluginWorld(OracleWorld):
    """Same interface again, executed by the NODES of the Zeno-side drop-in (zeno_b200/plugin/flipb200_nodes.cpp) on real
    OpenVDB objects, with the CPU oracle behind the C ABI they call (oracle/ref/plugin_nodes_test.cpp, prefix pn_)."""
    PREFIX = "pn_"

    @classmethod
    def _load(cls):
        lib = load_ref()
        if not getattr(lib, "_pn_ready", False):
            build()
            lib.pn_world_create.restype = C.c_void_p
            lib.pn_cfl.restype = C.c_float
            lib.pn_dropped.restype = C.c_uint64
            lib.pn_last_error.restype = C.c_char_p
            if lib.pn_backend(_LIB.encode()) != 0:
                raise RuntimeError("pn_backend: " + (lib.pn_last_error() or b"").decode())
            lib._pn_ready = True
        return lib

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("plugin node failed: " + (self.lib.orc_last_error() or b"").decode())

    def FLIP_P2G(self, dx=None, VelExtraLayer=3):
        self._ck(self.lib.orc_p2g(self.h, C.c_float(self.dx if dx is None else dx), C.c_int(VelExtraLayer)))

    def G2PAdvectorSheetty(self, dt, dx=None, surface_size=4, RK_ORDER=1, pic_min=0.03, pic_max=0.05, viscous_is_velocity=True):
        self._ck(self.lib.orc_g2p_advect_sheetty(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(surface_size),
                                                 C.c_int(RK_ORDER), C.c_float(pic_min), C.c_float(pic_max), C.c_int(1 if viscous_is_velocity else 0)))

    def CutCellWeight(self):
        self._ck(self.lib.orc_face_weights(self.h))

    def PushOutLiquidSDF(self, dx=None):
        self._ck(self.lib.orc_pushout_sdf(self.h, C.c_float(self.dx if dx is None else dx)))

    def FieldAddVector(self, x, y, z):
        self._ck(self.lib.orc_add_vector(self.h, C.c_float(x), C.c_float(y), C.c_float(z)))

    def AssembleSolvePPE(self, dt, dx=None, rel_tol=None, max_iter=100):
        it, res, st = C.c_int(0), C.c_float(0), C.c_int(0)
        self._ck(self.lib.orc_solve_ppe(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.byref(it), C.byref(res), C.byref(st)))
        return {"iterations": it.value, "rel_residual": res.value, "status": st.value}

    def SubtractPressureGradient(self, dt, dx=None, VelExtraLayer=3):
        self._ck(self.lib.orc_subtract_grad(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(VelExtraLayer)))

    def KillParticlesInSDF(self, sdf_grid: str = "KillerSDF", keep: bool = True):
        self._ck(self.lib.orc_kill_particles(self.h, C.c_int(GRID_IDS[sdf_grid]), C.c_int(1 if keep else 0)))

    def ParticleAddDV(self, x, y, z):
        self._ck(self.lib.orc_particles_add_dv(self.h, C.c_float(x), C.c_float(y), C.c_float(z)))

    def VDBRenormalizeSDF(self, grid: str = "LiquidSDF", iterations: int = 4, dilateIters: int = 0):
        self._ck(self.lib.orc_renormalize_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_int(iterations), C.c_int(dilateIters)))

    def VDBErodeSDF(self, grid: str, depth: float):
        self._ck(self.lib.orc_erode_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_float(depth)))

    def VDBSmoothSDF(self, grid: str, width: int = 1, iterations: int = 1):
        self._ck(self.lib.orc_smooth_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_int(width), C.c_int(iterations)))

    def G2P_Advector(self, dt, dx=None, RK_ORDER=1, pic_smoothness=0.02):
        self._ck(self.lib.orc_g2p_advect(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(RK_ORDER), C.c_float(pic_smoothness)))


class RefNodeWorld(PluginWorld):
    """Same interface, executed by the REFERENCE's own node classes (projects/FastFLIP/nosys/*.cpp compiled unmodified into
    oracle/_ref, oracle/ref/ref_nodes_test.cpp, prefix rn_) on real OpenVDB objects."""
    PREFIX = "rn_"

    @classmethod
    def _load(cls):
        lib = load_ref()
        if not getattr(lib, "_rn_ready", False):
            lib.rn_world_create.restype = C.c_void_p
            lib.rn_cfl.restype = C.c_float
            lib.rn_dropped.restype = C.c_uint64
            lib.rn_last_error.restype = C.c_char_p
            lib._rn_ready = True
        return lib

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("reference node failed: " + (self.lib.orc_last_error() or b"").decode())


_PLUGIN_GPU_LIB = os.path.join(_HERE, "_ref", "libflipplugin_gpu.so")
_plugin_gpu = None


def plugin_gpu_available() -> bool:
    return os.path.exists(_PLUGIN_GPU_LIB)


class PluginGpuWorld(PluginWorld):
    """The drop-in's NODE CLASSES on real OpenVDB objects with the PRODUCT behind them: oracle/_ref/libflipplugin_gpu.so
    (oracle/ref/plugin_nodes_gpu.cpp, prefix pg_) links zeno_b200/libflipb200.so. Needs a B200."""
    PREFIX = "pg_"

    @classmethod
    def _load(cls):
        global _plugin_gpu
        if _plugin_gpu is None:
            load_ref()   # libflipref.so first: the harness takes its RefWorld helpers from it
            lib = C.CDLL(_PLUGIN_GPU_LIB)
            lib.pg_world_create.restype = C.c_void_p
            lib.pg_cfl.restype = C.c_float
            lib.pg_dropped.restype = C.c_uint64
            lib.pg_last_error.restype = C.c_char_p
            _plugin_gpu = lib
        return _plugin_gpu
