"""Leaves per x layer of the P2G outputs on every rank of a two-rank in-process decomposition against the single world (a quick look at what the ghost refresh delivered). usage: python tools/dd_layers_probe.py"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from zeno_b200 import abi, scenes
from tests.test_dd_gpu import make_dd, owned_part
from tests import util
N=128; side=32; bounds=[(0,2),(2,4)]
pos, vel, dx = scenes.dam_break_points(N, seed=3, random_velocity=True, side=side)
vel = vel*0.2
solid = scenes.box_solid_sdf(N, dx)
one = abi.World(dx); one.set_grid("SolidSDF", solid); one.PrimToVDBPointDataGrid(pos, vel); one.FLIP_P2G(dx,3)
dd = make_dd(abi, N, bounds, pos, vel, dx, solid)
abi.run_ranks(dd, lambda r,w: w.FLIP_P2G(dx,3))
for name in ("Velocity","PostAdvVelocity","LiquidSDF"):
    ref = scenes.canonical_grid(one.get_grid(name))
    for r,w in enumerate(dd):
        lo,hi = w.dd_owned()
        g = scenes.canonical_grid(w.get_grid(name))
        lx = g["origins"][:,0]>>3
        cnt = {int(l): int((lx==l).sum()) for l in np.unique(lx)}
        rlx = ref["origins"][:,0]>>3
        rcnt = {int(l): int((rlx==l).sum()) for l in np.unique(rlx)}
        print(name, "rank", r, "owned", (lo,hi), "leaves per layer", cnt, "ref", rcnt)
