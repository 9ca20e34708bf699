"""CPU model of the exchange schedule of the sharded MGPCG (zeno_b200/csrc/poisson.cu dd_cycle0).

Claim: because whole ghost LEAVES (8 voxels deep) are exchanged, eight red/black colour passes fit between two exchanges.
What a rank computes in its ghost layer is wrong only within k voxels of the layer's far face after k passes, and its owned
unknowns read only the nearest ghost plane. The model runs the reference's red-black SOR (uaamg.cpp:1109-1150: colour by
(x+y+z)&1, x <- fma(x, 1-w, ((b - offdiag) * invdiag) * w), w = 1.2) on a 7-point Laplacian, once on the whole grid and once
on two overlapping halves with garbage beyond the ghost layer, and compares the owned values bit for bit.
"""
import numpy as np

W = np.float32(1.2)


def colour_pass(x, b, colour, x_lo_global=0):
    """one colour of RBGS on a dense block with zero Dirichlet values outside; in place"""
    nx, ny, nz = x.shape
    xp = np.zeros((nx + 2, ny + 2, nz + 2), np.float32)
    xp[1:-1, 1:-1, 1:-1] = x
    c = np.float32(-1.0)
    fx = xp[2:, 1:-1, 1:-1] * c + xp[:-2, 1:-1, 1:-1] * c
    fy = xp[1:-1, 2:, 1:-1] * c + xp[1:-1, :-2, 1:-1] * c
    fz = xp[1:-1, 1:-1, 2:] * c + xp[1:-1, 1:-1, :-2] * c
    off = (fx + fy) + fz
    inv = np.float32(1.0) / np.float32(6.0)
    t = ((b - off) * inv) * W
    new = x * (np.float32(1.0) - W) + t     # numpy has no fp32 fma; both sides of the comparison use this same expression
    ii, jj, kk = np.meshgrid(np.arange(nx) + x_lo_global, np.arange(ny), np.arange(nz), indexing="ij")
    m = ((ii + jj + kk) & 1) == colour
    x[m] = new[m]


def test_eight_colour_passes_fit_between_two_ghost_leaf_exchanges():
    rng = np.random.default_rng(11)
    nx, ny, nz = 48, 8, 8          # six leaf layers along x; rank 0 owns [0,24), rank 1 owns [24,48)
    b = rng.standard_normal((nx, ny, nz)).astype(np.float32)
    x = rng.standard_normal((nx, ny, nz)).astype(np.float32)
    ref = x.copy()
    for k in range(8):
        colour_pass(ref, b, k & 1)
    cut, ghost, ring = 24, 8, 8
    # rank 0: owned [0,24) + ghost leaf [24,32) exact after the exchange + one ring leaf [32,40) holding garbage
    lo0, hi0 = 0, cut + ghost + ring
    x0 = x[lo0:hi0].copy()
    x0[cut + ghost:] = 1.0e3
    b0 = b[lo0:hi0].copy()
    b0[cut + ghost:] = -7.0
    # rank 1: ring [8,16) garbage, ghost [16,24) exact, owned [24,48)
    lo1 = cut - ghost - ring
    x1 = x[lo1:].copy()
    x1[:ring] = -1.0e3
    b1 = b[lo1:].copy()
    b1[:ring] = 5.0
    for k in range(8):
        colour_pass(x0, b0, k & 1, lo0)
        colour_pass(x1, b1, k & 1, lo1)
    assert np.array_equal(x0[:cut], ref[:cut]), "rank 0: owned values differ after 8 passes"
    assert np.array_equal(x1[ring + ghost:], ref[cut:]), "rank 1: owned values differ after 8 passes"
    # ... and a ninth pass without a new exchange is NOT safe (the bound is tight)
    colour_pass(ref, b, 0)
    colour_pass(x0, b0, 0, lo0)
    assert not np.array_equal(x0[:cut], ref[:cut])


def test_balanced_slabs_properties():
    from zeno_b200 import dist_util
    rng = np.random.default_rng(3)
    for world in (2, 3, 4, 8):
        for _ in range(20):
            n_layers = int(rng.integers(2 * world, 6 * world))
            counts = np.zeros(n_layers + 6, np.int64)
            counts[3:3 + n_layers] = rng.integers(1, 1000, size=n_layers)
            b = dist_util.balanced_slabs(counts, world)
            assert b[0][0] == 3 and b[-1][1] == 3 + n_layers
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            assert all(hi - lo >= 2 for lo, hi in b)
