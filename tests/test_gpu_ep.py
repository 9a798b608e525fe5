"""GPU (>= 2 devices): expert-parallel layer over NCCL equals the single-GPU layer bit for bit."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2])
def test_ep_matches_single_gpu(world, lib_built):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(HERE, "ep_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "EP_WORKER_RESULT OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
