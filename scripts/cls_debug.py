"""Debug aid: large-batch classifier step vs the oracle for several batch sizes (tile-size paths)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import cuda_agent
from oracle import mobody_oracle as M
from mobody_b200 import _ffi
S, A, seed, std, lr = 17, 6, 21, 0.5, 3e-4
for N in (int(x) for x in sys.argv[1:]):
    rng = np.random.default_rng(5)
    ag, _ = cuda_agent(S, A, seed, penalty_type="dara", penalty_coef=1.0, gaussian_noise_std=std, actor_lr=lr, penalize_fake=0)
    cl = M.ClassifierState(S, A, seed)
    ag.classifier.load_state_dict(cl.params)
    s, a = rng.standard_normal((N, S)).astype(np.float32), rng.uniform(-1, 1, (N, A)).astype(np.float32)
    s2 = (s + 0.3 * rng.standard_normal((N, S))).astype(np.float32)
    label = (rng.random(N) < 0.5).astype(np.int64)
    s2[label == 1] += 0.25
    RW = _ffi.lib().mobody_row_width(S, A)
    rows = np.zeros((N, RW), np.float32)
    rows[:, :S], rows[:, S:S + A], rows[:, S + A:2 * S + A] = s, a, s2
    rows_d, label_d = torch.from_numpy(rows).cuda(), torch.from_numpy(label.astype(np.int32)).cuda()
    for it in range(2):
        n_sas, n_sa = rng.standard_normal((N, 2 * S + A)).astype(np.float32), rng.standard_normal((N, S + A)).astype(np.float32)
        want = M.classifier_update(cl, torch.from_numpy(s), torch.from_numpy(a), torch.from_numpy(s2), torch.from_numpy(label),
                                   torch.from_numpy(n_sas), torch.from_numpy(n_sa), std, lr)
        got = ag.classifier_step_on_rows(rows_d, label_d, noise_sas=n_sas, noise_sa=n_sa).cpu().numpy()
        print(N, it, got, want)
    for k, v in ag.classifier.state_dict().items():
        g, w = v.detach().cpu().numpy().reshape(-1), cl.params[k].numpy().reshape(-1)
        print("   ", k, "max abs diff", float(np.abs(g - w).max()), "frac > 1e-6", float(np.mean(np.abs(g - w) > 1e-6)))
