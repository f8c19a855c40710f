"""Multi-GPU rollout over real NCCL: runs scripts/nccl_check.py under torchrun on two GPUs of the
box when it has them (the single-GPU test box skips; the gloo world-size-2 tests cover the host logic on CPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_rollout_over_nccl_equals_single_gpu():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "nccl_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "NCCL_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
