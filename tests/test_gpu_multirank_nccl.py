"""GPU, >= 2 devices: the real multi-rank path — CUDA partial MSM per rank + NCCL all-gather of the 96-byte Jacobian partials +
device fold (util/msm.rs:322-336 across GPUs), and the pairing batch sharded across ranks — under torchrun, world size 2, against
the CPU oracle.  (tests/test_multirank_gloo.py covers the host-side sharding logic on CPU; this one runs the kernels and NCCL.)"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_cuda_partials_nccl_allgather_fold(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
