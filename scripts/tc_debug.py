"""Debug aid: run one fused step in every precision mode and print error statistics vs the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_dynamics, cuda_agent, rel_err
from oracle import mobody_oracle as M

S, A = int(os.environ.get("S", 17)), int(os.environ.get("A", 6))
B = int(os.environ.get("B", 200))
rng = np.random.default_rng(0)
obs = (np.r_[1.25, np.zeros(S - 1)][None] + 0.2 * rng.standard_normal((B, S))).astype(np.float32)
act = rng.uniform(-1, 1, (B, A)).astype(np.float32)
eps = rng.standard_normal((7, B, S)).astype(np.float32)
idx = rng.integers(0, 7, B)
for prec in sys.argv[1:] or ["fp32", "bf16x2", "bf16"]:
    dyn, p = cuda_dynamics(S, A, 41, "walker2d", 5.0, precision=prec, h0=0.95, t3_gain=4.0)
    ref = M.step(p, torch.from_numpy(obs), torch.from_numpy(act), torch.from_numpy(eps), idx, 3, 5.0)
    nobs, rew, term, info = dyn.step(torch.from_numpy(obs).cuda(), torch.from_numpy(act).cuda(), eps=eps, idx=idx)
    torch.cuda.synchronize()
    print(prec, "mean", rel_err(info["samples"].cpu().numpy(), ref["mean"].numpy()),
          "next", rel_err(nobs.cpu().numpy(), ref["next_obs"].numpy()),
          "raw", rel_err(info["raw_reward"].cpu().numpy(), ref["raw_reward"].numpy()),
          "pen", rel_err(info["penalty"].cpu().numpy(), ref["penalty"].numpy()),
          "term_mismatch", int((term != ref["terminal"]).sum()), flush=True)
    if os.environ.get("VERBOSE"):
        print(info["samples"][0, :2].cpu().numpy(), ref["mean"][0, :2].numpy())
    # fused policy path
    from mobody_b200.dynamics import StepWorkspace
    ag, st = cuda_agent(S, A, 41)
    ws = StepWorkspace(B, S, A, "cuda", want_act=True)
    o = torch.from_numpy(obs).cuda()
    dyn.launch_step(o, None, ws, policy=ag.policy.network, max_action=1.0, eps=torch.from_numpy(eps).cuda(), idx=torch.from_numpy(idx).cuda())
    torch.cuda.synchronize()
    act_ref = M.policy_forward(st.policy, torch.from_numpy(obs), 1.0)
    ref2 = M.step(p, torch.from_numpy(obs), act_ref, torch.from_numpy(eps), idx, 3, 5.0)
    print(prec, "policy-fused act", rel_err(ws.act.cpu().numpy(), act_ref.numpy()), "next", rel_err(ws.next_obs.cpu().numpy(), ref2["next_obs"].numpy()),
          "rew", rel_err(ws.reward.cpu().numpy(), ref2["reward"].numpy()), flush=True)
