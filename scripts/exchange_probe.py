"""Where the per-step cost of the peer-memory exchange goes (2+ GPUs, torchrun): CUDA-event timings of
 (a) a local rollout packing into ordinary device memory, (b) the same rollout packing into this rank's slot of the symmetric
 receive buffer, no push, (c) rollout + push back to back on one stream (no overlap), (d) the pipelined loop of bench.py."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import mobody_b200 as mb
from mobody_b200 import parallel as P
dyn = bench.build_dynamics(mb, bench.S, bench.A, "bf16x2", dev)
ag = bench.build_agent(mb, bench.S, bench.A, dev); ag.dynamics = dyn
Bn = 100_000
x = torch.from_numpy(bench.synth_obs(Bn, 100 + rank)).to(dev)

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    host = (time.perf_counter() - t0) / n * 1e3
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4), round(host, 4)

res = {}
res["a local rollout (ordinary memory)"] = timed(lambda: ag.rollout_device(x, 1, row0=rank * Bn, sync=False))
def p2p(overlap):
    def f():
        h = P.sharded_rollout(ag, x, 1, sharded_input=True, gather="p2p")
        h.wait()
    return f
for ex_overlap in ("0", "1"):
    os.environ["MOBODY_PUSH_OVERLAP"] = ex_overlap
    ag.__dict__.pop("_peer_exchanges", None)
    res[f"c rollout + push + wait, one stream, push {'on side stream' if ex_overlap == '1' else 'inline'}"] = timed(p2p(ex_overlap))
ex = list(ag._peer_exchanges.values())[0]
# (b): pack into the symmetric slot without any push: drive the rollout desc directly
import ctypes as C
from mobody_b200 import _ffi
ws = ag._rollout_workspace(1, Bn, bench.S, bench.A, 7)
rows, _ = ex.views(2)
d, keep = ag._rollout_desc(x, 1, True, ws, rows[ex.rank], row0=rank * Bn)
res["b local rollout packing into the symmetric buffer, no push"] = timed(lambda: _ffi.check(_ffi.lib().mobody_rollout(C.byref(d), _ffi.stream_ptr(dev))))
if rank == 0:
    for k, v in res.items():
        print(f"{k:90s} device ms/step {v[0]:8.4f}   host enqueue ms/step {v[1]:8.4f}", flush=True)
dist.barrier(); dist.destroy_process_group()
