import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_agent
from oracle import mobody_oracle as M
from mobody_b200 import _ffi
S, A, N, n_true = 27, 8, 1500, 1200
CFG = dict(gamma=0.99, tau=0.005, actor_lr=3e-4, critic_lr=3e-4, weight=2.5, bc_coef=1.0, max_action=1.0)
rng = np.random.default_rng(N)
ag, st = cuda_agent(S, A, 55)
RW = _ffi.lib().mobody_row_width(S, A)
s, a = rng.standard_normal((N, S)).astype(np.float32), rng.uniform(-1, 1, (N, A)).astype(np.float32)
s2, r = rng.standard_normal((N, S)).astype(np.float32), rng.standard_normal((N, 1)).astype(np.float32)
nd = (rng.random((N, 1)) > 0.1).astype(np.float32)
rows = np.zeros((N, RW), np.float32)
rows[:, :S], rows[:, S:S + A], rows[:, S + A:2 * S + A], rows[:, 2 * S + A:2 * S + A + 1], rows[:, 2 * S + A + 1:2 * S + A + 2] = s, a, s2, r, nd
rows_d = torch.from_numpy(rows).cuda()
batch = tuple(torch.from_numpy(x) for x in (s, a, s2, r, nd))
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for it in range(nsteps):
    M.train_step(st, batch, n_true, CFG); ag.train_on_rows(rows_d, n_true)
for grp, mod, ref in (("pi", ag.policy, st.policy), ("q", ag.q_funcs, st.q), ("qt", ag.target_q_funcs, st.q_target)):
    for k, v in mod.state_dict().items():
        d = np.abs(v.cpu().numpy() - ref[k].numpy()).reshape(-1)
        print(grp, k, "max abs", d.max(), "n>3e-5", int((d > 3e-5).sum()), "n>3e-6", int((d > 3e-6).sum()), "of", d.size)
