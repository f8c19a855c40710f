"""A few dynamics fitting steps at the reference's batch size (for ncu launch lists / captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
import mobody_b200 as mb
dev = torch.device("cuda:0")
dyn = bench.build_dynamics(mb, bench.S, bench.A, "bf16x2", dev)
data = [torch.from_numpy(x).to(dev) for x in bench.fit_batch_data(256)]
for _ in range(4):
    sc = dyn.fit_batch(True, *data)
torch.cuda.synchronize(); print("ok", sc.cpu().tolist())
