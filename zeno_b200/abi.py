"""ctypes binding of libflipb200 (include/flipb200.h) -- the host-side driver used by the
parity tests and bench.py. It mirrors the reference's node interface one-to-one (same node
names as projects/FastFLIP/nosys/*.cpp, same argument meaning) and does no computing of its
own: every method is one C-ABI call. There is no CPU fallback; a missing library or a missing
sm_100 device raises.

Grid exchange format (numpy), identical to the OpenVDB leaf layout (see flipb200.h):
    {"origins": int32[n,3], "masks": uint64[n,8], "values": float32[n,C,512], "bg": float32[C]}
Particle exchange format (the reference's PointDataGrid layout, FF/FLIP_vdb.h:28-37):
    {"origins": int32[n,3], "voxel_end": uint32[n,512], "P": uint16[N,3], "v": uint16[N,3]}
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

GRID_IDS = {
    "Velocity": 0, "PostAdvVelocity": 1, "ViscousVelocity": 2, "SolidVelocity": 3, "CellFWeight": 4,
    "LiquidSDF": 5, "SolidSDF": 6, "Pressure": 7, "Divergence": 8, "Curvature": 9, "KillerSDF": 10,
}
VEC_GRIDS = {"Velocity", "PostAdvVelocity", "ViscousVelocity", "SolidVelocity", "CellFWeight"}
SOA, AOS = 0, 1

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libflipb200.so")

# every symbol include/flipb200.h declares (the CPU-side loader test checks this list against the header)
EXPORTS = [
    "flipb200_last_error", "flipb200_build_info", "flipb200_abi_version", "flipb200_device_count",
    "flipb200_world_create", "flipb200_world_destroy", "flipb200_host_alloc", "flipb200_host_free", "flipb200_grid_upload", "flipb200_grid_leaf_count",
    "flipb200_grid_download", "flipb200_particles_upload", "flipb200_particles_info",
    "flipb200_particles_download", "flipb200_bin_from_points", "flipb200_p2g",
    "flipb200_g2p_advect_sheetty", "flipb200_dropped", "flipb200_capture_precodec", "flipb200_get_precodec",
    "flipb200_face_weights", "flipb200_pushout_sdf", "flipb200_add_vector", "flipb200_cfl",
    "flipb200_solve_ppe", "flipb200_solve_ppe_ex", "flipb200_solver_info", "flipb200_residual_history",
    "flipb200_subtract_grad", "flipb200_substep", "flipb200_launch_count", "flipb200_profile_enable",
    "flipb200_profile_reset", "flipb200_profile_get", "flipb200_stream", "flipb200_comm_unique_id",
    "flipb200_comm_init", "flipb200_comm_init_local", "flipb200_comm_abort", "flipb200_dd_set_slab", "flipb200_dd_owned",
    "flipb200_dd_owned_particles",
    "flipb200_sync_count", "flipb200_smooth_sdf", "flipb200_renormalize_sdf", "flipb200_erode_sdf", "flipb200_g2p_advect", "flipb200_kill_particles_in_sdf", "flipb200_fluid_reseed", "flipb200_emit_liquid", "flipb200_apply_boundary", "flipb200_particles_to_points", "flipb200_set_surface_tension", "flipb200_particles_add_dv", "flipb200_particles_download_begin", "flipb200_grid_download_begin", "flipb200_download_wait",
]


class FlipB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libflipb200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen libflipb200.so (built in-tree by zeno_b200/csrc/build.sh). Fails loudly."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    lib.flipb200_last_error.restype = C.c_char_p
    lib.flipb200_build_info.restype = C.c_char_p
    lib.flipb200_abi_version.restype = C.c_int
    lib.flipb200_device_count.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class PinnedArena:
    """Page-locked host arrays (flipb200_host_alloc) for staging grids / particles across PCIe."""

    def __init__(self, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        self._ptrs = []

    def empty(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        rc = self.lib.flipb200_host_alloc(C.c_size_t(max(n, 1)), C.byref(ptr))
        if rc != 0:
            raise FlipB200Error(rc, (self.lib.flipb200_last_error() or b"").decode())
        self._ptrs.append(ptr)
        buf = (C.c_byte * max(n, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def like(self, a: np.ndarray, slack: float = 1.0) -> np.ndarray:
        """pinned copy of a; slack > 1 reserves extra leading-dimension capacity"""
        shape = list(a.shape)
        cap = int(np.ceil(shape[0] * slack)) if shape else 0
        out = self.empty([cap] + shape[1:], a.dtype)
        out[:shape[0]] = a
        return out

    def close(self):
        for p in self._ptrs:
            self.lib.flipb200_host_free(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class World:
    """One FLIP world on one GPU (SetFLIPWorld, FF/nosys/FLIP_Creator.cpp)."""

    def __init__(self, dx: float, device: int = 0, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        self.dx = float(dx)
        self.h = C.c_void_p()
        self._ck(self.lib.flipb200_world_create(C.c_int(device), C.c_float(dx), C.byref(self.h)))

    # -- plumbing ------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise FlipB200Error(rc, (self.lib.flipb200_last_error() or b"").decode())

    def close(self):
        if self.h:
            self.lib.flipb200_world_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- marshalling ---------------------------------------------------------------
    def set_grid(self, name: str, g: Dict[str, np.ndarray], layout: int = SOA):
        gid = GRID_IDS[name]
        o = _c(g["origins"], np.int32).reshape(-1, 3)
        m = _c(g["masks"], np.uint64).reshape(-1, 8)
        v = _c(g["values"], np.float32)
        bg = _c(g["bg"], np.float32)
        n = o.shape[0]
        self._ck(self.lib.flipb200_grid_upload(self.h, C.c_int(gid), C.c_int(n), _p(o), _p(m), _p(v), C.c_int(layout), _p(bg)))

    def get_grid(self, name: str, layout: int = SOA, out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """out: optional pre-allocated (e.g. pinned) arrays with leading capacity >= the leaf count; views are returned"""
        gid = GRID_IDS[name]
        nch = 3 if name in VEC_GRIDS else 1
        n = C.c_int(0)
        self._ck(self.lib.flipb200_grid_leaf_count(self.h, C.c_int(gid), C.byref(n)))
        n = n.value
        if out is not None and out["origins"].shape[0] >= n:
            o, m, v = out["origins"][:n], out["masks"][:n], out["values"][:n]
        else:
            o = np.zeros((n, 3), np.int32)
            m = np.zeros((n, 8), np.uint64)
            v = np.zeros((n, nch, 512) if layout == SOA else (n, 512, nch), np.float32)
        bg = np.zeros(nch, np.float32)
        self._ck(self.lib.flipb200_grid_download(self.h, C.c_int(gid), _p(o), _p(m), _p(v), C.c_int(layout), _p(bg)))
        return {"origins": o, "masks": m, "values": v, "bg": bg}

    def set_particles(self, p: Dict[str, np.ndarray]):
        o = _c(p["origins"], np.int32).reshape(-1, 3)
        ve = _c(p["voxel_end"], np.uint32).reshape(-1, 512)
        P = _c(p["P"], np.uint16).reshape(-1, 3)
        v = _c(p["v"], np.uint16).reshape(-1, 3)
        self._ck(self.lib.flipb200_particles_upload(self.h, C.c_int(o.shape[0]), _p(o), _p(ve), C.c_uint64(P.shape[0]), _p(P), _p(v)))

    def particles_info(self):
        nl, n = C.c_int(0), C.c_uint64(0)
        self._ck(self.lib.flipb200_particles_info(self.h, C.byref(nl), C.byref(n)))
        return nl.value, n.value

    def get_particles(self, out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        nl, n = self.particles_info()
        if out is not None and out["origins"].shape[0] >= nl and out["P"].shape[0] >= n:
            o, ve, P, v = out["origins"][:nl], out["voxel_end"][:nl], out["P"][:n], out["v"][:n]
        else:
            o = np.zeros((nl, 3), np.int32)
            ve = np.zeros((nl, 512), np.uint32)
            P = np.zeros((n, 3), np.uint16)
            v = np.zeros((n, 3), np.uint16)
        self._ck(self.lib.flipb200_particles_download(self.h, _p(o), _p(ve), _p(P), _p(v)))
        return {"origins": o, "voxel_end": ve, "P": P, "v": v}

    # -- asynchronous downloads into caller-provided (pinned) buffers; valid after download_wait() -------------
    def get_particles_begin(self, out: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        nl, n = C.c_int(0), C.c_uint64(0)
        self._ck(self.lib.flipb200_particles_download_begin(self.h, C.c_int(out["origins"].shape[0]), C.c_uint64(out["P"].shape[0]),
                                                            _p(out["origins"]), _p(out["voxel_end"]), _p(out["P"]), _p(out["v"]),
                                                            C.byref(nl), C.byref(n)))
        return {"origins": out["origins"][:nl.value], "voxel_end": out["voxel_end"][:nl.value], "P": out["P"][:n.value], "v": out["v"][:n.value]}

    def get_grid_begin(self, name: str, out: Dict[str, np.ndarray], layout: int = SOA) -> Dict[str, np.ndarray]:
        gid = GRID_IDS[name]
        nch = 3 if name in VEC_GRIDS else 1
        n = C.c_int(0)
        bg = np.zeros(nch, np.float32)
        self._ck(self.lib.flipb200_grid_download_begin(self.h, C.c_int(gid), C.c_int(out["origins"].shape[0]), _p(out["origins"]), _p(out["masks"]), _p(out["values"]),
                                                       C.c_int(layout), _p(bg), C.byref(n)))
        return {"origins": out["origins"][:n.value], "masks": out["masks"][:n.value], "values": out["values"][:n.value], "bg": bg}

    def download_wait(self):
        self._ck(self.lib.flipb200_download_wait(self.h))

    # -- nodes (names follow the reference's ZENDEFNODE names) ------------------------
    def PrimToVDBPointDataGrid(self, pos: np.ndarray, vel: Optional[np.ndarray] = None):
        pos = _c(pos, np.float32).reshape(-1, 3)
        vel = None if vel is None else _c(vel, np.float32).reshape(-1, 3)
        self._ck(self.lib.flipb200_bin_from_points(self.h, _p(pos), _p(vel), C.c_uint64(pos.shape[0])))

    def FLIP_P2G(self, dx: Optional[float] = None, VelExtraLayer: int = 3):
        self._ck(self.lib.flipb200_p2g(self.h, C.c_float(self.dx if dx is None else dx), C.c_int(VelExtraLayer)))

    def G2PAdvectorSheetty(self, dt: float, dx: Optional[float] = None, surface_size: int = 4, RK_ORDER: int = 1,
                           pic_min: float = 0.03, pic_max: float = 0.05, viscous_is_velocity: bool = True):
        self._ck(self.lib.flipb200_g2p_advect_sheetty(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx),
                                                      C.c_int(surface_size), C.c_int(RK_ORDER), C.c_float(pic_min),
                                                      C.c_float(pic_max), C.c_int(1 if viscous_is_velocity else 0)))

    def FluidReseed(self, seed: int = 0):
        self._ck(self.lib.flipb200_fluid_reseed(self.h, C.c_uint32(seed & 0xffffffff)))

    def ParticleEmitter(self, shape_grid: str = "KillerSDF", vx=0.0, vy=0.0, vz=0.0, seed: int = 0):
        self._ck(self.lib.flipb200_emit_liquid(self.h, C.c_int(GRID_IDS[shape_grid]), C.c_float(vx), C.c_float(vy), C.c_float(vz), C.c_uint32(seed & 0xffffffff)))

    def FLIPApplyBoundary(self, moving_grid: str = "KillerSDF", moving_vertex_centred: bool = False):
        self._ck(self.lib.flipb200_apply_boundary(self.h, C.c_int(GRID_IDS[moving_grid]), C.c_int(1 if moving_vertex_centred else 0)))

    def set_surface_tension(self, density: float = 1000.0, coef: float = 0.0):
        self._ck(self.lib.flipb200_set_surface_tension(self.h, C.c_float(density), C.c_float(coef)))

    def VDBPointsToPrimitive(self):
        n = self.particles_info()[1]
        pos, vel = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        self._ck(self.lib.flipb200_particles_to_points(self.h, pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p)))
        return pos, vel

    def KillParticlesInSDF(self, sdf_grid: str = "KillerSDF", keep: bool = True):
        self._ck(self.lib.flipb200_kill_particles_in_sdf(self.h, C.c_int(GRID_IDS[sdf_grid]), C.c_int(1 if keep else 0)))

    def ParticleAddDV(self, x: float, y: float, z: float):
        self._ck(self.lib.flipb200_particles_add_dv(self.h, C.c_float(x), C.c_float(y), C.c_float(z)))

    def G2P_Advector(self, dt: float, dx: Optional[float] = None, RK_ORDER: int = 1, pic_smoothness: float = 0.02):
        self._ck(self.lib.flipb200_g2p_advect(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(RK_ORDER), C.c_float(pic_smoothness)))

    def VDBRenormalizeSDF(self, grid: str = "LiquidSDF", iterations: int = 4, dilateIters: int = 0):
        self._ck(self.lib.flipb200_renormalize_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_int(iterations), C.c_int(dilateIters)))

    def VDBErodeSDF(self, grid: str, depth: float):
        self._ck(self.lib.flipb200_erode_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_float(depth)))

    def VDBSmoothSDF(self, grid: str, width: int = 1, iterations: int = 1):
        self._ck(self.lib.flipb200_smooth_sdf(self.h, C.c_int(GRID_IDS[grid]), C.c_int(width), C.c_int(iterations)))

    def dropped(self) -> int:
        n = C.c_uint64(0)
        self._ck(self.lib.flipb200_dropped(self.h, C.byref(n)))
        return n.value

    def capture_precodec(self, on: bool):
        self._ck(self.lib.flipb200_capture_precodec(self.h, C.c_int(1 if on else 0)))

    def get_precodec(self, n: int):
        pos = np.zeros((n, 3), np.float32)
        vel = np.zeros((n, 3), np.float32)
        alive = np.zeros(n, np.uint8)
        self._ck(self.lib.flipb200_get_precodec(self.h, _p(pos), _p(vel), _p(alive)))
        return pos, vel, alive

    def CutCellWeight(self):
        self._ck(self.lib.flipb200_face_weights(self.h))

    def PushOutLiquidSDF(self, dx: Optional[float] = None):
        self._ck(self.lib.flipb200_pushout_sdf(self.h, C.c_float(self.dx if dx is None else dx)))

    def FieldAddVector(self, x: float, y: float, z: float):
        self._ck(self.lib.flipb200_add_vector(self.h, C.c_float(x), C.c_float(y), C.c_float(z)))

    def CFL_dt(self) -> float:
        out = C.c_float(0)
        self._ck(self.lib.flipb200_cfl(self.h, C.byref(out)))
        return out.value

    def AssembleSolvePPE(self, dt: float, dx: Optional[float] = None, rel_tol: Optional[float] = None, max_iter: int = 100):
        it, res, st = C.c_int(0), C.c_float(0), C.c_int(0)
        d = C.c_float(self.dx if dx is None else dx)
        if rel_tol is None:
            self._ck(self.lib.flipb200_solve_ppe(self.h, C.c_float(dt), d, C.byref(it), C.byref(res), C.byref(st)))
        else:
            self._ck(self.lib.flipb200_solve_ppe_ex(self.h, C.c_float(dt), d, C.c_float(rel_tol), C.c_int(max_iter),
                                                    C.byref(it), C.byref(res), C.byref(st)))
        return {"iterations": it.value, "rel_residual": res.value, "status": st.value}

    def solver_info(self):
        lv, nd, nh = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.flipb200_solver_info(self.h, C.byref(lv), C.byref(nd), C.byref(nh)))
        hist = np.zeros(nh.value, np.float32)
        if nh.value:
            self._ck(self.lib.flipb200_residual_history(self.h, _p(hist)))
        return {"levels": lv.value, "num_dof": nd.value, "history": hist}

    def SubtractPressureGradient(self, dt: float, dx: Optional[float] = None, VelExtraLayer: int = 3):
        self._ck(self.lib.flipb200_subtract_grad(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(VelExtraLayer)))

    def substep(self, dt: float, dx: Optional[float] = None, surface_size: int = 4, RK_ORDER: int = 3, pic_min: float = 0.03,
                pic_max: float = 0.05, gravity=(0.0, -9.8, 0.0), VelExtraLayer: int = 3, viscous_is_velocity: bool = True,
                want_stage_ms: bool = False):
        ms = (C.c_float * 5)()
        self._ck(self.lib.flipb200_substep(self.h, C.c_float(dt), C.c_float(self.dx if dx is None else dx), C.c_int(surface_size),
                                           C.c_int(RK_ORDER), C.c_float(pic_min), C.c_float(pic_max), C.c_float(gravity[0]),
                                           C.c_float(gravity[1]), C.c_float(gravity[2]), C.c_int(VelExtraLayer),
                                           C.c_int(1 if viscous_is_velocity else 0), ms if want_stage_ms else None))
        return list(ms) if want_stage_ms else None

    # -- multi-GPU: slab decomposition along x (include/flipb200.h "multi-GPU") --------------
    def comm_init_nccl(self, rank: int, world: int, unique_id: bytes):
        """one process per GPU; unique_id = the 128 bytes rank 0 got from comm_unique_id()"""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.flipb200_comm_init(self.h, C.c_int(rank), C.c_int(world), buf))

    def comm_abort(self):
        self.lib.flipb200_comm_abort(self.h)

    def dd_set_slab(self, leaf_lo: int, leaf_hi: int):
        self._ck(self.lib.flipb200_dd_set_slab(self.h, C.c_int(leaf_lo), C.c_int(leaf_hi)))

    def dd_owned(self):
        lo, hi = C.c_int(0), C.c_int(0)
        self._ck(self.lib.flipb200_dd_owned(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def dd_owned_particles(self) -> int:
        n = C.c_uint64(0)
        self._ck(self.lib.flipb200_dd_owned_particles(self.h, C.byref(n)))
        return n.value

    # -- measurement hooks -----------------------------------------------------------
    def launch_count(self) -> int:
        n = C.c_uint64(0)
        self._ck(self.lib.flipb200_launch_count(self.h, C.byref(n)))
        return n.value

    def sync_count(self) -> int:
        n = C.c_uint64(0)
        self._ck(self.lib.flipb200_sync_count(self.h, C.byref(n)))
        return n.value

    def profile_enable(self, on: bool):
        self._ck(self.lib.flipb200_profile_enable(self.h, C.c_int(1 if on else 0)))

    def profile_reset(self):
        self._ck(self.lib.flipb200_profile_reset(self.h))

    def profile_get(self):
        cap = 128
        names = C.create_string_buffer(8192)
        ms = (C.c_float * cap)()
        ln = (C.c_uint64 * cap)()
        by = (C.c_uint64 * cap)()
        n = C.c_int(0)
        self._ck(self.lib.flipb200_profile_get(self.h, names, C.c_size_t(8192), ms, ln, by, C.c_int(cap), C.byref(n)))
        keys = names.value.decode().split(";") if n.value else []
        return {k: {"ms": ms[i], "launches": ln[i], "bytes": by[i]} for i, k in enumerate(keys)}

    def stream(self) -> int:
        s = C.c_void_p()
        self._ck(self.lib.flipb200_stream(self.h, C.byref(s)))
        return s.value or 0


def comm_unique_id(lib: Optional[C.CDLL] = None) -> bytes:
    lib = lib or load_library()
    buf = (C.c_uint8 * 128)()
    rc = lib.flipb200_comm_unique_id(buf)
    if rc != 0:
        raise FlipB200Error(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable? import torch first)")
    return bytes(buf)


def comm_init_local(worlds) -> None:
    """Connect several worlds of THIS process (one host thread each) with the in-process communicator."""
    lib = worlds[0].lib
    arr = (C.c_void_p * len(worlds))(*[w.h for w in worlds])
    rc = lib.flipb200_comm_init_local(arr, C.c_int(len(worlds)))
    if rc != 0:
        raise FlipB200Error(rc, "comm_init_local failed")


def run_ranks(worlds, fn):
    """Run fn(rank, world) for every world on its own host thread (ctypes releases the GIL inside library calls);
    a failing rank aborts the in-process communicator so its peers do not wait forever. Returns the results by rank."""
    import threading
    res, err = [None] * len(worlds), [None] * len(worlds)

    def body(r):
        try:
            res[r] = fn(r, worlds[r])
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            for w in worlds:
                w.comm_abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(len(worlds))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return res
