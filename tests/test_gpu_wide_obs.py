"""Wide observations (gym Ant-v3 / D4RL ant-*: obs 111, act 8 — SURVEY.md section 8: "kernels must take S, A at run
time, S <= 128").  The tensor-core kernel takes S <= 64; precision 'auto' routes wider observations to the fp32
CUDA-core kernel on 32-row tiles.  Parity against the oracle: step, multi-step rollout, one train step."""
import numpy as np
import pytest
import torch

from helpers import cuda_agent, cuda_dynamics, rel_err
from oracle import mobody_oracle as M

pytestmark = pytest.mark.gpu
S, A = 111, 8


def test_wide_step_rollout_and_train_match_oracle():
    rng = np.random.default_rng(3)
    dyn, p = cuda_dynamics(S, A, 5, "ant", 1.0, precision=None)
    assert dyn.precision == "fp32"                               # 'auto' resolved to the CUDA-core kernel
    B = 150                                                      # 32-row tiles: 4 full + 1 ragged
    obs = (np.r_[0.6, np.zeros(S - 1)][None] + 0.1 * rng.standard_normal((B, S))).astype(np.float32)
    act = rng.uniform(-1, 1, (B, A)).astype(np.float32)
    eps = rng.standard_normal((7, B, S)).astype(np.float32)
    idx = rng.integers(0, 7, B)
    ref = M.step(p, torch.from_numpy(obs), torch.from_numpy(act), torch.from_numpy(eps), idx, 4, 1.0)
    nobs, rew, term, info = dyn.step(torch.from_numpy(obs).cuda(), torch.from_numpy(act).cuda(), eps=eps, idx=idx)
    assert rel_err(nobs.cpu().numpy(), ref["next_obs"].numpy()) < 1e-4
    assert rel_err(rew.cpu().numpy(), ref["reward"].numpy()) < 1e-4
    assert rel_err(info["penalty"].cpu().numpy(), ref["penalty"].numpy()) < 1e-4
    assert rel_err(info["samples"].cpu().numpy(), ref["mean"].numpy()) < 1e-4
    assert np.array_equal(term, ref["terminal"])
    # rollout through mobody_rollout (policy fused, compaction, filter) against the oracle with the same Philox draws
    ag, st = cuda_agent(S, A, 5, env_filter=1e9)
    ag.dynamics = dyn
    T = 2
    eps_t = rng.standard_normal((T, 7, B, S)).astype(np.float32)
    idx_t = rng.integers(0, 7, (T, B))
    tr, inf = ag.rollout(torch.from_numpy(obs).cuda(), T, True, eps=eps_t, idx=idx_t)
    want, winf = M.rollout(p, st.policy, 1.0, torch.from_numpy(obs), T, [lambda n, t=t: torch.from_numpy(eps_t[t][:, :n]) for t in range(T)],
                           [lambda n, t=t: idx_t[t][:n] for t in range(T)], 4, 1.0, 1e9, True)
    assert inf["num_transitions"] == winf["num_transitions"]
    for k in ("obss", "actions", "next_obss", "rewards", "terminals"):
        assert tuple(tr[k].shape) == tuple(want[k].shape), k
        assert rel_err(tr[k].numpy(), want[k].numpy()) < 1e-4, k
    # one train step on wide rows
    from mobody_b200 import _ffi
    N, n_true = 96, 64
    RW = _ffi.lib().mobody_row_width(S, A)
    s, a = rng.standard_normal((N, S)).astype(np.float32), rng.uniform(-1, 1, (N, A)).astype(np.float32)
    s2, r = rng.standard_normal((N, S)).astype(np.float32), rng.standard_normal((N, 1)).astype(np.float32)
    nd = (rng.random((N, 1)) > 0.1).astype(np.float32)
    rows = np.zeros((N, RW), np.float32)
    rows[:, :S], rows[:, S:S + A], rows[:, S + A:2 * S + A], rows[:, 2 * S + A:2 * S + A + 1], rows[:, 2 * S + A + 1:2 * S + A + 2] = s, a, s2, r, nd
    cfg = dict(gamma=0.99, tau=0.005, actor_lr=3e-4, critic_lr=3e-4, weight=2.5, bc_coef=1.0, max_action=1.0)
    wantl = M.train_step(st, tuple(torch.from_numpy(x) for x in (s, a, s2, r, nd)), n_true, cfg)
    ag.train_on_rows(torch.from_numpy(rows).cuda(), n_true)
    got = ag.loss_scalars()
    for k in ("q_loss", "pi_loss", "bc_loss", "q1_mean", "q_policy", "w_mean"):
        assert abs(got[k] - wantl[k]) <= 1e-4 * (abs(wantl[k]) + 1e-2), (k, got[k], wantl[k])
