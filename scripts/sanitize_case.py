"""Smallest end-to-end case for compute-sanitizer: one fused step per precision, one 2-step rollout, one train step,
one classifier step, small, ragged and large (32-row-tile) batches.  Run plain (prints finite losses) or under
compute-sanitizer --tool memcheck where the pool allows it."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_dynamics, cuda_agent
import mobody_b200 as mb

S, A, B = 17, 6, 300
rng = np.random.default_rng(0)
obs = torch.from_numpy((np.r_[1.25, np.zeros(S - 1)][None] + 0.1 * rng.standard_normal((B, S))).astype(np.float32)).cuda()
act = torch.from_numpy(rng.uniform(-1, 1, (B, A)).astype(np.float32)).cuda()
for prec in ("fp32", "bf16x2", "bf16", "fp16"):
    dyn, _ = cuda_dynamics(S, A, 5, "walker2d", 5.0, precision=prec)
    dyn.step(obs, act)
    ag, _ = cuda_agent(S, A, 5, env_filter=1e9)
    ag.dynamics = dyn
    out, info = ag.rollout_device(obs, 2)
    print(prec, info["rows_per_step"], info["kept"])
src, tar = mb.ReplayBuffer(S, A, "cuda", max_size=4000), mb.ReplayBuffer(S, A, "cuda", max_size=1000)
for b, n in ((src, 4000), (tar, 1000), (ag.fake_replay_buffer, 2000)):
    b.add_batch({"obss": rng.standard_normal((n, S)).astype(np.float32), "actions": rng.uniform(-1, 1, (n, A)).astype(np.float32),
                 "next_obss": rng.standard_normal((n, S)).astype(np.float32), "rewards": rng.standard_normal((n, 1)).astype(np.float32),
                 "terminals": np.zeros((n, 1), np.float32)})
ag.total_it = 1
ag.train(src, tar, 64)
ag.train(src, tar, 2048)            # 5 120 rows: 32-row tiles (two CTAs per SM), tensor-core weight gradients, 27 row splits
ag.train(src, tar, 31)              # ragged: 77 rows
ag.update_classifier(src, tar, 64)
ag.update_classifier(src, tar, 2500)
ag.dara_relabel(src)
torch.cuda.synchronize()
print("ok", ag.loss_scalars()["q_loss"])
