"""The slab decomposition over the real transport: one process per GPU (torchrun), the library's NCCL communicator.
Needs two GPUs on the box; tests/test_dd_gpu.py covers the same logic on one GPU through the in-process communicator."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_over_nccl(gpu_lib):
    n = gpu_lib.flipb200_device_count()
    if n < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dd_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_DD_OK" in r.stdout, r.stdout[-3000:]
    print([ln for ln in r.stdout.splitlines() if "NCCL_DD_OK" in ln][0])
