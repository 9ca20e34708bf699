"""torchrun worker of tests/test_nccl_gpu.py: one rank per GPU, the library's NCCL communicator, the same checks as
tests/test_dd_gpu.py (owned results of the decomposed run against a single world run on rank 0)."""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_check(rank, world, local, dist, box=None):
    """The decomposed chain on `world` ranks against ONE world on rank 0, same tank. Returns the report line on rank 0
    (raises on a mismatch). box = voxel extents (y, z) of the water block; its x extent is 16 voxels (two leaf layers) per rank.
    Used by tests/test_nccl_gpu.py (cube) and by bench.py --gpus N before the timed region (a thin block, a few seconds)."""
    import torch
    from tests import util
    from tests.test_dd_gpu import DT, G, owned_part, owned_particles, merge_grids, split_points
    from zeno_b200 import abi, scenes

    N, side = 128, 16 * world
    bounds = [(2 * r, 2 * r + 2) for r in range(world)]
    if box is None:
        pos, vel, dx = scenes.dam_break_points(N, seed=3, random_velocity=True, side=side)
    else:
        pos, vel, dx = scenes.dam_break_points(max(N, side + 16), seed=3, random_velocity=True, box=((0, side), (0, box[0]), (0, box[1])))
        N = max(N, side + 16)
    vel = vel * 0.2
    solid = scenes.box_solid_sdf(N, dx)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(abi.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    w = abi.World(dx, device=local)
    w.comm_init_nccl(rank, world, uid.cpu().numpy().tobytes())
    w.dd_set_slab(*bounds[rank])
    w.set_grid("SolidSDF", solid)
    w.PrimToVDBPointDataGrid(*split_points(pos, vel, dx, bounds)[rank])
    out = {}
    lo, hi = w.dd_owned()
    out["binned"] = owned_particles(w.get_particles(), lo, hi)
    w.FLIP_P2G(dx, 3)
    out["p2g_Velocity"] = owned_part(w.get_grid("Velocity"), lo, hi)
    out["p2g_LiquidSDF"] = owned_part(w.get_grid("LiquidSDF"), lo, hi)
    w.CutCellWeight()
    w.PushOutLiquidSDF(dx)
    w.FieldAddVector(G[0] * DT, G[1] * DT, G[2] * DT)
    out["ppe"] = w.AssembleSolvePPE(DT, dx)
    out["Pressure"] = owned_part(w.get_grid("Pressure"), lo, hi)
    w.SubtractPressureGradient(DT, dx, 3)
    out["Velocity"] = owned_part(w.get_grid("Velocity"), lo, hi)
    w.G2PAdvectorSheetty(DT, dx, 4, 3, 0.03, 0.05, True)
    out["advected"] = owned_particles(w.get_particles(), lo, hi)
    out["dt"] = w.CFL_dt()
    w.substep(DT, dx, 4, 3, 0.03, 0.05, G, 3, True)
    out["n_after"] = w.dd_owned_particles()
    gathered = [None] * world
    dist.all_gather_object(gathered, pickle.dumps(out))
    if rank == 0:
        parts = [pickle.loads(b) for b in gathered]
        one = abi.World(dx, device=local)
        one.set_grid("SolidSDF", solid)
        one.PrimToVDBPointDataGrid(pos, vel)

        def rows(key):
            b = np.concatenate([p[key] for p in parts])
            return b[np.lexsort(tuple(b[:, k] for k in range(8, -1, -1)))]
        assert np.array_equal(rows("binned"), scenes.canonical_particles(one.get_particles()))
        one.FLIP_P2G(dx, 3)
        for name in ("Velocity", "LiquidSDF"):
            util.compare_grids(merge_grids([p["p2g_" + name] for p in parts]), one.get_grid(name), f"nccl P2G {name}", tol=0.0, check_inactive=False)
        one.CutCellWeight()
        one.PushOutLiquidSDF(dx)
        one.FieldAddVector(G[0] * DT, G[1] * DT, G[2] * DT)
        r1 = one.AssembleSolvePPE(DT, dx)
        assert all(p["ppe"]["status"] == 0 for p in parts) and abs(parts[0]["ppe"]["iterations"] - r1["iterations"]) <= 1, (r1, [p["ppe"] for p in parts])
        e_p = util.compare_grids(merge_grids([p["Pressure"] for p in parts]), one.get_grid("Pressure"), "nccl Pressure", tol=1e-5, check_inactive=False)
        one.SubtractPressureGradient(DT, dx, 3)
        e_v = util.compare_grids(merge_grids([p["Velocity"] for p in parts]), one.get_grid("Velocity"), "nccl Velocity", tol=1e-5, check_inactive=False)
        one.G2PAdvectorSheetty(DT, dx, 4, 3, 0.03, 0.05, True)
        a, b = scenes.canonical_particles(one.get_particles()), rows("advected")
        assert a.shape == b.shape
        same = (a[:, :3] == b[:, :3]).all(axis=1).mean()
        assert same > 0.999, same
        assert abs(parts[0]["dt"] - one.CFL_dt()) <= 1e-5 * parts[0]["dt"]
        one.substep(DT, dx, 4, 3, 0.03, 0.05, G, 3, True)
        assert sum(p["n_after"] for p in parts) == one.particles_info()[1]
        report = (f"NCCL_DD_OK ranks={world} particles={a.shape[0]} iterations={parts[0]['ppe']['iterations']}/{r1['iterations']} "
                  f"pressure_relL2={e_p:.2e} velocity_relL2={e_v:.2e} same_voxel={same:.6f}")
        one.close()
    else:
        report = None
    dist.barrier()
    w.close()
    return report


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    report = run_check(rank, world, local, dist)
    if rank == 0:
        print(report)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
