"""2-rank NCCL run of the sharded detector (tools/check_multi_gpu.py): on every rank the all-gathered result must equal
what one engine computes for the whole global batch.  Needs two GPUs on the box (`gpurun --gpus 2`); skipped otherwise.
The host-side shard / gather logic is covered on CPU with gloo in tests/test_host_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_detector_two_ranks_nccl():
    port = str(29600 + os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("sharded result == single-engine result") == 2 and "False" not in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_native_training_under_ddp_two_ranks():
    port = str(29900 + os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "check_ddp_train.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("parameters identical across ranks after 6 DDP steps: True") == 2
