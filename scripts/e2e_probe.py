"""Debug aid: where the end-to-end time of MOBODY.rollout() goes (copies vs kernels vs host)."""
import sys, os, time, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_dynamics, cuda_agent

S, A, B = 17, 6, 100_000
dyn, _ = cuda_dynamics(S, A, 1, "halfcheetah", 5.0, precision="bf16x2")
ag, _ = cuda_agent(S, A, 1, env_filter=10.0)
ag.dynamics = dyn
obs_host = torch.from_numpy((0.3 * np.random.default_rng(0).standard_normal((B, S))).astype(np.float32)).pin_memory()
dev = torch.device("cuda")
W = 2 * S + A + 3
slab_d = torch.empty(B, W, device=dev); slab_h = torch.empty(B, W).pin_memory()

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print("H2D 6.8MB  ms", timeit(lambda: obs_host.to(dev, non_blocking=True)))
print("D2H 17.2MB ms", timeit(lambda: slab_h.copy_(slab_d, non_blocking=True)))
x = obs_host.to(dev)
print("rollout_device sync=False ms", timeit(lambda: ag.rollout_device(x, 1, sync=False)))
print("rollout_device sync=True  ms", timeit(lambda: ag.rollout_device(x, 1)))
def quiet(fn):
    def g():
        with contextlib.redirect_stdout(io.StringIO()): fn()
    return g
print("rollout(host) pipelined   ms", timeit(quiet(lambda: ag.rollout(obs_host, 1))))
for pr in (18944, 37888, 56832):
    ag.PIPE_ROWS = pr
    print(f"rollout(host) PIPE_ROWS={pr} ms", timeit(quiet(lambda: ag.rollout(obs_host, 1))))
ag.PIPE_ROWS = 10**9
print("rollout(host) unpipelined ms", timeit(quiet(lambda: ag.rollout(obs_host, 1))))

# host-side cost of one asynchronous rollout_device call (Python + ctypes + launches), GPU idle in between
import cProfile, pstats
ag.PIPE_ROWS = 37888
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    ag.rollout_device(x, 1, sync=False)
host = (time.perf_counter() - t0) / 50 * 1e3
torch.cuda.synchronize()
print("host time per rollout_device(sync=False) call ms", host)
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    ag.rollout_device(x, 1, sync=False)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
