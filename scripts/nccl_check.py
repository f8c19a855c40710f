"""Multi-GPU check of parallel.sharded_rollout on real hardware (launch with torchrun, >= 2 ranks): the transitions gathered
from the shards -- through the peer-memory push (csrc/peer.cu) and through the NCCL padded all-gather -- must equal what a
single GPU produces for the same global start states (parallel.self_check), with real terminations and a real filter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from mobody_b200 import parallel as P
from helpers import cuda_agent, cuda_dynamics

S, A = 17, 6
dyn, _ = cuda_dynamics(S, A, 21, "walker2d", 5.0, precision="bf16x2", h0=1.0, t3_gain=6.0)
dyn.seed = 5
ag, _ = cuda_agent(S, A, 21, env_filter=0.78)
ag.dynamics = dyn
res = P.self_check(ag, 10_007, rounds=6)                      # 10 007: not divisible by the world size
if rank == 0:
    print("NCCL_CHECK", "OK" if res["ok"] else "FAILED", res)
dist.destroy_process_group()
sys.exit(0 if res["ok"] else 1)
