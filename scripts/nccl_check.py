"""Multi-GPU check of parallel.sharded_rollout over real NCCL (launch with torchrun, >= 2 ranks):
the all-gathered transitions of the sharded rollout must equal, as a set and rank-major in order, what a single GPU
produces for the same global start states (Philox is keyed on the global row id), for T = 1 and T = 3, through the
compact (gather=True) and the padded-slab (gather="padded") exchanges."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from mobody_b200 import parallel as P
from helpers import cuda_agent, cuda_dynamics

S, A, B = 17, 6, 10_007                                       # not divisible by the world size
dyn, _ = cuda_dynamics(S, A, 21, "walker2d", 5.0, precision="bf16x2", h0=1.0, t3_gain=6.0)
dyn.seed = 5
ag, _ = cuda_agent(S, A, 21, env_filter=0.78)
ag.dynamics = dyn
rng = np.random.default_rng(0)
obs = torch.from_numpy((np.r_[1.25, np.zeros(S - 1)][None] + 0.15 * rng.standard_normal((B, S))).astype(np.float32)).cuda()
ok = True
for T in (1, 3):
    full, fi = ag.rollout_device(obs, T)                       # every rank: the single-GPU answer for all rows
    out, info = P.sharded_rollout(ag, obs, T)                  # compact all-gather-v
    ok &= info["num_transitions"] == fi["num_transitions"] and info["kept"] == fi["kept"]
    ok &= abs(info["reward_mean"] - fi["reward_mean"]) <= 1e-5 * abs(fi["reward_mean"]) + 1e-6
    cat = lambda d: torch.cat([d[k] for k in P.KEYS], 1)       # noqa: E731
    a, b = cat(out), cat(full)
    ok &= a.shape == b.shape
    key = lambda m: m[np.lexsort(m.cpu().numpy().T[::-1])] if m.numel() else m    # noqa: E731
    ok &= bool(torch.equal(key(a), key(b)))                    # same multiset of transitions
    if T == 1:                                                 # one step: rank-major order == global row order
        ok &= bool(torch.equal(a, b))
    (slabs, counts_dev, widths), pinfo = P.sharded_rollout(ag, obs, T, gather="padded")
    cnt = [int(c) for c in counts_dev.cpu()]
    ok &= sum(cnt) == fi["kept"]
    ok &= bool(torch.equal(torch.cat([slabs[r, :cnt[r]] for r in range(world)], 0), a))
    assert 0 < fi["kept"] < fi["num_transitions"] and (T == 1 or fi["num_transitions"] < B * T), "test case must filter and terminate"
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL_CHECK", "OK" if flag.item() == 1.0 else "FAILED", "world", world)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
