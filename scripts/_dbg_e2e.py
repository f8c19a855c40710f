import sys, os, time, io, contextlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench as Bn
import mobody_b200 as mb
from mobody_b200 import _ffi
from mobody_b200.dynamics import StepWorkspace
dev = torch.device("cuda", 0)
dyn = Bn.build_dynamics(mb, 17, 6, "bf16x2", dev); ag = Bn.build_agent(mb, 17, 6, dev); ag.dynamics = dyn
obs_host = torch.from_numpy(Bn.synth_obs(100000, 100)).pin_memory()
obs_dev = obs_host.to(dev)
def e2e(tag):
    for _ in range(2):
        with contextlib.redirect_stdout(io.StringIO()): ag.rollout(obs_host, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        with contextlib.redirect_stdout(io.StringIO()): tr, info = ag.rollout(obs_host, 1)
    torch.cuda.synchronize(); print(tag, "e2e ms/step", (time.perf_counter() - t0) / 20 * 1e3, "allocs", getattr(ag, "_slab_allocs", 0), flush=True)
e2e("fresh")
for _ in range(23): ag.rollout_device(obs_dev, 1, sync=False)
torch.cuda.synchronize(); e2e("after device loop")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ws = StepWorkspace(100000, 17, 6, dev, want_act=True)
for it in range(23):
    flush.zero_(); dyn.launch_step(obs_dev, None, ws, policy=ag.policy.network, max_action=1.0, step=it)
torch.cuda.synchronize(); e2e("after kernel-only loop")
s = Bn.ClockSampler(0); s.start(); e2e("with sampler 5ms"); s.period = 0.05; e2e("with sampler 50ms"); print(s.stop())
e2e("sampler stopped")
