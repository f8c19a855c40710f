"""Debug aid: per-layer timeline (clock64) of CTA 0 of the tensor-core step kernel."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import cuda_dynamics, cuda_agent
from mobody_b200 import _ffi
from mobody_b200.dynamics import StepWorkspace

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128 * 148
S, A = 17, 6
dyn, p = cuda_dynamics(S, A, 1, "halfcheetah", 5.0, precision=prec)
ag, _ = cuda_agent(S, A, 1)
obs = torch.randn(B, S, device="cuda") * 0.3
ws = StepWorkspace(B, S, A, "cuda", want_act=True)
trace = torch.zeros(80 * 8, dtype=torch.int64, device="cuda")
L = _ffi.lib()
for it in range(3):
    dyn.launch_step(obs, None, ws, policy=ag.policy.network, max_action=1.0, step=it)
torch.cuda.synchronize()
L.mobody_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
dyn.launch_step(obs, None, ws, policy=ag.policy.network, max_action=1.0, step=9)
torch.cuda.synchronize()
L.mobody_debug_set_trace(None)
t = trace.cpu().numpy().reshape(80, 8)
n = 73
t0 = t[0, 0]
names = ["P1", "P2", "P3"] + sum([[f"{e}.zs1", f"{e}.zs2", f"{e}.zs3", f"{e}.za1", f"{e}.za2", f"{e}.t1", f"{e}.t2", f"{e}.t3"] for e in range(7)], []) + sum([[f"{e}.r1", f"{e}.r2"] for e in range(7)], [])
print("layer      mma_start first_issue mma_end | g0_dfull g0_end g1_dfull g1_end | mma_dur epi0_dur  dfull-mma_end  period")
prev = t0
for i in range(n):
    r = t[i] - t0
    print(f"{names[i]:8s} {r[0]:9d} {r[1]:9d} {r[2]:9d} | {r[3]:9d} {r[4]:9d} {r[5]:9d} {r[6]:9d} | {r[2]-r[0]:7d} {r[4]-r[3]:7d} {r[3]-r[2]:7d} {r[4]-(prev-t0):7d}")
    prev = t[i, 4]
print("total cycles", t[n - 1, 4] - t0)
print("stats phase stamps (rel. to t3(6) epilogue end):", (t[79, :7] - t[58, 4]).tolist())

print("row-iteration stamps (rel. to its start):", (t[78, :6] - t[78, 0]).tolist(), "iter start rel. loop start", int(t[78,0]-t[79,1]))
