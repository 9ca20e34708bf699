"""NCCL transport and all-reduce time between the GPUs of the box (torchrun --nproc-per-node N tools/nccl_probe.py)."""
import os, time, torch, torch.distributed as dist
dist.init_process_group("nccl")
r = dist.get_rank(); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
for mb in (1, 11, 44):
    x = torch.ones(mb * 262144, device="cuda")
    for _ in range(3): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter()
    for _ in range(10): dist.all_reduce(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    if r == 0: print(f"all_reduce {mb} MB: {dt*1e6:.0f} us  ({mb/1e3/dt:.1f} GB/s algorithmic)", flush=True)
