"""A few steady-state train steps (for ncu launch lists)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import mobody_b200 as mb
from helpers import cuda_agent
S, A = bench.S, bench.A
ag, _ = cuda_agent(S, A, 2, penalty_type="none")
src, tar = mb.ReplayBuffer(S, A, "cuda"), mb.ReplayBuffer(S, A, "cuda")
src.convert_D4RL(bench.synth_buffer_dict(200_000, 1)); tar.convert_D4RL(bench.synth_buffer_dict(20_000, 2))
ag.fake_replay_buffer.convert_D4RL(bench.synth_buffer_dict(50_000, 3))
ag.total_it = 1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for _ in range(6):
    ag.train(src, tar, B)
torch.cuda.synchronize()
print("ok", ag.loss_scalars())
