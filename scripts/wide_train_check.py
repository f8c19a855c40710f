"""Debug aid: where the post-Adam policy layer-0 weights differ from the oracle at wide observations."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_agent
from oracle import mobody_oracle as M
from mobody_b200 import _ffi
from test_gpu_train import CFG
S, A, N, n_true = (int(x) for x in sys.argv[1:5])
steps = int(sys.argv[5])
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
w0 = ag.policy.state_dict()["network.network.0.weight"].cpu().numpy().copy()
for it in range(steps):
    M.train_step(st, batch, n_true, CFG); ag.train_on_rows(rows_d, n_true)
for grp, mod, ref in (("pi", ag.policy, st.policy), ("q", ag.q_funcs, st.q)):
    for k, v in mod.state_dict().items():
        g, w = v.cpu().numpy().astype(np.float64), ref[k].numpy().astype(np.float64)
        d = np.abs(g - w); bad = d > 1e-4 * (np.abs(w) + np.mean(np.abs(w)))
        print(grp, k, g.shape, "bad frac", float(bad.mean()), "max abs diff", float(d.max()))
        if k == "network.network.0.weight" and grp == "pi":
            print("   bad per input column (first 24):", np.round(bad.mean(0)[:24], 2), " columns >= 64:", np.round(bad.mean(0)[64:72], 2))
            print("   bad per unit: n units with any bad", int((bad.sum(1) > 0).sum()), "of", bad.shape[0], "; units fully bad", int((bad.mean(1) > 0.9).sum()))
            dg, dw = g - w0, w - w0
            print("   update sign disagreement frac", float(np.mean(np.sign(dg) != np.sign(dw))), "mean |update| gpu/oracle", float(np.abs(dg).mean()), float(np.abs(dw).mean()))
