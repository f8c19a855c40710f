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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_bench_line_carries_a_green_exchange_check():
    """bench.py under torchrun on two GPUs (the driver's scaling launch): the line must carry the pre-timing self-check
    (peer-memory result == NCCL result == single-GPU result) and the post-timing consistency check of the two-stream loop."""
    import json
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "6", "--warmup", "3",
                        "--no-cpu-baseline", "--no-train", "--no-extras"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    assert line["n_gpus"] == 2 and line["exchange"]["mode"] == "p2p"
    chk = line["exchange"]["check"]
    assert chk["ok"] and chk["p2p"] and chk["nccl"] and chk["timed_last_step"]["ok"], chk
    assert line["value"] > 1.5 * 50e6                      # two GPUs, well above one GPU's single-stream rate
