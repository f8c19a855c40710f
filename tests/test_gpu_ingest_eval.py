"""Dataset ingest -> device-resident buffers, and the evaluators with the CUDA policy / dynamics."""
import numpy as np
import pytest
import torch

from helpers import cuda_agent, cuda_dynamics
from test_host_logic import _VecEnv

pytestmark = pytest.mark.gpu


def test_ingest_to_device_buffers():
    import mobody_b200 as mb
    rng = np.random.default_rng(0)
    n, S, A = 1001, 17, 6
    raw = {"observations": rng.standard_normal((n, S)), "actions": rng.uniform(-1, 1, (n, A)), "rewards": rng.standard_normal((n, 1)),
           "terminals": rng.random(n) < 0.05, "timeouts": rng.random(n) < 0.01}
    tr = mb.transitions_from_raw(raw)
    src, tar = mb.load_buffers(tr, tr, S, A, "cuda")
    assert src.size == tar.size == n - 1
    assert np.array_equal(src.state[:n - 1].cpu().numpy(), tr["observations"])
    assert np.array_equal(src.next_state[:n - 1].cpu().numpy(), raw["observations"][1:].astype(np.float32))
    assert np.array_equal(src.reward[:n - 1, 0].cpu().numpy(), tr["rewards"])
    assert np.array_equal(src.not_done[:n - 1, 0].cpu().numpy(), 1.0 - tr["terminals"].astype(np.float32))
    a, b = src.sample(64), tar.sample(64)
    assert not torch.equal(a[0], b[0])                       # independent index streams per buffer


def test_evaluators_with_cuda_policy_and_model_check(capsys):
    import mobody_b200 as mb
    S, A = 5, 2
    ag, st = cuda_agent(S, A, 3)
    dyn, _ = cuda_dynamics(S, A, 3, "halfcheetah", 1.0)
    lengths = [4, 9, 6]
    got = mb.eval_policy_batch(ag, _VecEnv(lengths, S), ag.policy, len(lengths), eval_cnt=1, dynamics=dyn, eval_trg=True)
    out = capsys.readouterr().out
    assert np.isfinite(got) and "reward mse" in out and "obs mse" in out and "Evaluation on target over 3 episodes" in out
    # the same episodes one at a time through eval_policy: a single-copy env per episode
    class One:
        def __init__(self, k): self.v = _VecEnv([lengths[k]], S)
        def reset(self): return self.v.reset()[0]
        def step(self, a):
            ns, r, d, i = self.v.step(np.asarray(a).reshape(1, -1))
            return ns[0], float(r[0]), bool(d[0]), i
    r0 = mb.eval_policy(ag, One(0), ag.policy, 2, eval_cnt=0)
    assert np.isfinite(r0)
