"""Deterministic synthetic inputs for the FastFLIP hot path: a dam-break water
cube in a box tank, seeded with the reference's own hash frand (FF/FLIP_vdb.h:10-17).
Used by tests/ and bench.py; pure numpy, no device code."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def frand(i: np.ndarray) -> np.ndarray:
    """frand (projects/FastFLIP/FLIP_vdb.h:10-17), vectorised over uint32."""
    i = np.asarray(i, dtype=np.uint32)
    with np.errstate(over="ignore"):
        value = (i ^ np.uint32(61)) ^ (i >> np.uint32(16))
        value = value * np.uint32(9)
        value = value ^ (value << np.uint32(4))
        value = value * np.uint32(0x27D4EB2D)
        value = value ^ (value >> np.uint32(15))
    return value.astype(np.float32) / np.float32(4294967296.0)


def dam_break_points(N: int, seed: int = 1, ppc: int = 8, W: int = 0, side: int | None = None,
                     random_velocity: bool = False, box=None) -> Tuple[np.ndarray, np.ndarray, float]:
    """Water cube [W, W+side)^3 voxels in an N^3 tank, dx = 1/N, `ppc` particles per cell placed per
    octant: p_local = +-0.25 + 0.25*(frand(seed + 3*pid + c) - 0.5). Returns world pos, vel, dx.
    box = ((x0, x1), (y0, y1), (z0, z1)) voxel ranges replaces the cube (one rank's part of a larger block)."""
    dx = np.float32(1.0 / N)
    side = N // 4 if side is None else side
    ii = np.arange(W, W + side, dtype=np.int64)
    if box is None:
        gx, gy, gz = np.meshgrid(ii, ii, ii, indexing="ij")
    else:
        gx, gy, gz = np.meshgrid(*[np.arange(a, b, dtype=np.int64) for a, b in box], indexing="ij")
    cells = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)  # [V,3]
    V = cells.shape[0]
    reps = (ppc + 7) // 8
    octs = np.arange(ppc) % 8
    sign = np.stack([(octs >> 2) & 1, (octs >> 1) & 1, octs & 1], axis=1).astype(np.float32) * 2 - 1  # [ppc,3]
    pid = (np.arange(V, dtype=np.uint64)[:, None] * np.uint64(ppc) + np.arange(ppc, dtype=np.uint64)[None, :]).astype(np.uint32)
    jit = np.empty((V, ppc, 3), np.float32)
    for c in range(3):
        jit[:, :, c] = frand((np.uint32(seed) + np.uint32(3) * pid + np.uint32(c)).astype(np.uint32))
    scale = np.float32(0.25) if reps == 1 else np.float32(0.2)
    local = sign[None, :, :] * np.float32(0.25) + scale * (jit - np.float32(0.5))
    idx = cells[:, None, :].astype(np.float64) + local.astype(np.float64)
    pos = (idx * np.float64(dx)).astype(np.float32).reshape(-1, 3)
    if random_velocity:
        vel = np.empty((V, ppc, 3), np.float32)
        for c in range(3):
            vel[:, :, c] = 2.0 * frand((np.uint32(seed * 7919 + 17) + np.uint32(3) * pid + np.uint32(c)).astype(np.uint32)) - 1.0
        vel = vel.reshape(-1, 3)
    else:
        vel = np.zeros_like(pos)
    return pos, vel, float(dx)


def box_solid_sdf(N: int, dx: float, band: int = 3, inset: float = 0.02) -> Dict[str, np.ndarray]:
    """Analytic tank: solid outside the box, vertex-centred samples (FF/nosys/FLIP_Creator.cpp:95-98:
    solid voxel (i,j,k) sits at index position ijk - 0.5). phi = (min_c min(i_c, N - i_c) - inset) * dx,
    clamped to +-band*dx; only leaves intersecting the band are stored, background = 3dx."""
    lo, hi = -8, N + 8
    nl = (hi - lo) // 8
    # per-axis min over the 8 voxels of a leaf of (min(i, N - i) - inset); a leaf is kept when it
    # comes within band+1 voxels of a wall surface (interior leaves hold only background)
    ax = lo + 8 * np.arange(nl)
    vox = ax[:, None] + np.arange(8)[None, :]
    dax = (np.minimum(vox, N - vox).astype(np.float32) - np.float32(inset)).min(axis=1)  # [nl]
    near = dax < band + 1
    keep = near[:, None, None] | near[None, :, None] | near[None, None, :]
    lx, ly, lz = np.nonzero(keep)
    origins = np.stack([ax[lx], ax[ly], ax[lz]], axis=1).astype(np.int32)
    n = origins.shape[0]
    r = np.arange(8)
    X = origins[:, 0, None] + r[None, :]
    Y = origins[:, 1, None] + r[None, :]
    Z = origins[:, 2, None] + r[None, :]
    dX = np.minimum(X, N - X).astype(np.float32) - np.float32(inset)
    dY = np.minimum(Y, N - Y).astype(np.float32) - np.float32(inset)
    dZ = np.minimum(Z, N - Z).astype(np.float32) - np.float32(inset)
    phi = np.minimum(np.minimum(dX[:, :, None, None], dY[:, None, :, None]), dZ[:, None, None, :])
    phi = np.clip(phi, -band, band).astype(np.float32) * np.float32(dx)
    values = phi.reshape(n, 1, 512)
    active = (np.abs(phi) < np.float32(band * dx)).reshape(n, 8, 64)
    weights = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    masks = (active.astype(np.uint64) * weights[None, None, :]).sum(axis=2).astype(np.uint64)
    return {"origins": origins, "masks": masks, "values": values, "bg": np.array([3.0 * dx], np.float32)}


def empty_grid(nch: int, bg) -> Dict[str, np.ndarray]:
    return {"origins": np.zeros((0, 3), np.int32), "masks": np.zeros((0, 8), np.uint64),
            "values": np.zeros((0, nch, 512), np.float32), "bg": np.asarray(bg, np.float32).reshape(nch)}


def canonical_grid(g: Dict[str, np.ndarray], drop_empty: bool = True) -> Dict[str, np.ndarray]:
    """Sort leaves by origin and (optionally) drop leaves without active voxels, so two grids can
    be compared leaf by leaf. An all-inactive leaf holding only background equals no leaf."""
    o, m, v = g["origins"], g["masks"], g["values"]
    if drop_empty and o.shape[0]:
        keep = m.any(axis=1)
        o, m, v = o[keep], m[keep], v[keep]
    if o.shape[0]:
        order = np.lexsort((o[:, 2], o[:, 1], o[:, 0]))
        o, m, v = o[order], m[order], v[order]
    return {"origins": o, "masks": m, "values": v, "bg": g["bg"]}


def mask_bits(masks: np.ndarray) -> np.ndarray:
    """uint64[n,8] -> bool[n,512] in voxel-offset order."""
    n = masks.shape[0]
    b = np.unpackbits(masks.view(np.uint8).reshape(n, 64), axis=1, bitorder="little")
    return b.astype(bool)


def canonical_particles(p: Dict[str, np.ndarray]) -> np.ndarray:
    """Order-independent representation: one row (x,y,z voxel, P[3], v[3]) per particle, rows sorted.
    Within-voxel order is not defined by the reference (non-determinism)."""
    o, ve = p["origins"], p["voxel_end"].astype(np.int64)
    n = p["P"].shape[0]
    if n == 0:
        return np.zeros((0, 9), np.int64)
    counts = np.diff(np.concatenate([np.zeros((ve.shape[0], 1), np.int64), ve], axis=1), axis=1)  # [nl,512]
    off = np.arange(512)
    vx = o[:, 0, None] + (off >> 6)[None, :]
    vy = o[:, 1, None] + ((off >> 3) & 7)[None, :]
    vz = o[:, 2, None] + (off & 7)[None, :]
    c = counts.ravel()
    rows = np.stack([np.repeat(vx.ravel(), c), np.repeat(vy.ravel(), c), np.repeat(vz.ravel(), c)], axis=1)
    full = np.concatenate([rows, p["P"].astype(np.int64), p["v"].astype(np.int64)], axis=1)
    order = np.lexsort(tuple(full[:, k] for k in range(8, -1, -1)))
    return full[order]


def sphere_sdf(centre, radius: float, lo, hi, bg: float = 3.0) -> Dict[str, np.ndarray]:
    """A float grid in flat leaf layout holding (|ijk - centre| - radius) on every leaf that overlaps the voxel box [lo, hi)
    (index space, unit = voxels), clamped to +-bg, all voxels active. A generic "killer" / sink shape for tests."""
    lo = (np.asarray(lo) // 8) * 8
    hi = -((-np.asarray(hi)) // 8) * 8
    ax = [np.arange(lo[a], hi[a], 8) for a in range(3)]
    ox, oy, oz = np.meshgrid(*ax, indexing="ij")
    origins = np.stack([ox.ravel(), oy.ravel(), oz.ravel()], axis=1).astype(np.int32)
    n = origins.shape[0]
    r = np.arange(8)
    X = (origins[:, 0, None] + r[None, :]).astype(np.float32) - np.float32(centre[0])
    Y = (origins[:, 1, None] + r[None, :]).astype(np.float32) - np.float32(centre[1])
    Z = (origins[:, 2, None] + r[None, :]).astype(np.float32) - np.float32(centre[2])
    d = np.sqrt(X[:, :, None, None] ** 2 + Y[:, None, :, None] ** 2 + Z[:, None, None, :] ** 2) - np.float32(radius)
    values = np.clip(d, -bg, bg).astype(np.float32).reshape(n, 1, 512)
    masks = np.full((n, 8), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    return {"origins": origins, "masks": masks, "values": values, "bg": np.array([bg], np.float32)}
