"""Debug aid: error statistics of every precision mode of the fused step against the reference goldens."""
import sys, os, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_dynamics, rel_err
from mobody_b200 import _ffi

for prec in _ffi.ENABLED_PRECISIONS:
    worst = {}
    for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "step_*.npz"))):
        g = np.load(f)
        env, S, A, seed = str(g["env"]), int(g["S"]), int(g["A"]), int(g["seed"])
        dyn, p = cuda_dynamics(S, A, seed, env, float(g["coef"]), precision=prec)
        members = p["elites"].numpy()[g["idx"]]
        nobs, rew, term, info = dyn.step(torch.from_numpy(g["obs"]).cuda(), torch.from_numpy(g["act"]).cuda(), True, bool(g["use_trg"]),
                                         eps=g["eps"], idx=members)
        for k, got, want in (("next_obs", nobs, g["next_obs"]), ("reward", rew, g["reward"]), ("raw_reward", info["raw_reward"], g["raw_reward"]),
                             ("penalty", info["penalty"], g["penalty"]), ("mean", info["samples"], g["mean"])):
            worst[k] = max(worst.get(k, 0.0), rel_err(got.cpu().numpy(), want))
        worst["mask_flips"] = worst.get("mask_flips", 0) + int((term != g["terminal"]).sum())
    print(prec, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in worst.items()})
