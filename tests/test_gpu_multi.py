"""Launches tests/multi_gpu_check.py under torchrun when the box has at least two GPUs."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_sharded_paths():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(HERE, "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(res.stdout[-3000:])
    assert res.returncode == 0 and "MULTI_GPU_CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
