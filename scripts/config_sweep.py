"""BASELINE.json configs as concrete synthetic runs: one line of JSON per config with the
transitions/s of MOBODY.rollout_device (device resident, nothing read back inside the timed loop), rows per step and
kept fraction.  Single GPU; the sharded configs are run at their per-GPU share."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_dynamics, cuda_agent, healthy

CONFIGS = [
    ("C1 walker2d-friction S17/A6 T=1", "walker2d", 17, 6, 50_000, 1),
    ("C2 halfcheetah-gravity S17/A6 T=1", "halfcheetah", 17, 6, 100_000, 1),
    ("C3 hopper-kinematic S11/A3 T=5 (1M start states / 8 GPUs)", "hopper", 11, 3, 125_000, 5),
    ("C3 hopper-kinematic S11/A3 T=5 (1M start states on one GPU)", "hopper", 11, 3, 1_000_000, 5),
    ("C4 ant-friction S27/A8 T=5", "ant", 27, 8, 50_000, 5),
] + [(f"C5 antmaze-umaze S29/A8 T={T} B={B}", "ant", 29, 8, B, T) for T in (1, 5) for B in (10_000, 30_000, 100_000, 300_000, 1_000_000, 4_000_000)]

sel = sys.argv[1:]                      # optional substrings selecting configs
for name, env, S, A, B, T in CONFIGS:
    if sel and not any(x in name for x in sel):
        continue
    torch.cuda.empty_cache()
    dyn, _ = cuda_dynamics(S, A, 1, env, 5.0, precision="bf16x2", t3_gain=0.3)
    ag, _ = cuda_agent(S, A, 1, env_filter=10.0)
    ag.dynamics = dyn
    rng = np.random.default_rng(0)
    obs = torch.from_numpy((healthy(env, S)[None] + 0.05 * rng.standard_normal((B, S))).astype(np.float32)).cuda()
    packed = torch.empty(T * B, 2 * S + A + 3, dtype=torch.float32, device="cuda")   # result slab, allocated once (no cudaMalloc in the timed loop)
    out, info = ag.rollout_device(obs, T, out_packed=packed)     # warm-up + host-visible counts
    iters = 5 if B * T <= 1_000_000 else 2
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ag.rollout_device(obs, T, sync=False, out_packed=packed)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"config": name, "start_states": B, "T": T, "rows_per_step": info["rows_per_step"], "transitions": info["num_transitions"],
                      "kept": info["kept"], "ms": round(ms, 3), "transitions_per_s": round(info["num_transitions"] / (ms * 1e-3))}), flush=True)
    del ag, dyn, obs, out, packed
    ag = dyn = None
