"""GPU parity of the fused step against (a) reference-generated golden vectors and (b) the CPU oracle.

Tolerances (north_star): termination masks and indices bit-exact; fp32 outputs <= 1e-4 relative
(fp32 and bf16x2 modes); single-pass fp16 mode: within the stated looser bound 5e-3.
rel = |a-b| / (|b| + 1e-3).
"""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import cuda_dynamics, rel_err, rel_err_strict
from oracle import mobody_oracle as M

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16x2": 1e-4, "fp16": 5e-3}   # single-pass fp16 GEMMs meet the stated looser bound 5e-3


def precisions():
    from mobody_b200 import _ffi
    return list(getattr(_ffi, "ENABLED_PRECISIONS", ("fp32",)))


def _check(prec, out, ref_next, ref_rew, ref_raw, ref_pen, ref_mean, ref_term, near=None):
    nobs, rew, term, info = out
    tol = TOL[prec]
    assert rel_err(info["samples"].cpu().numpy(), ref_mean) < tol
    assert rel_err(nobs.cpu().numpy(), ref_next) < tol
    assert rel_err(info["raw_reward"].cpu().numpy(), ref_raw) < tol
    assert rel_err(info["penalty"].cpu().numpy(), ref_pen) < tol
    assert rel_err(rew.cpu().numpy(), ref_rew) < tol
    if prec in ("fp32", "bf16x2"):      # the fp32-parity modes: 1e-4 under the strict metric too
        for got, want in ((info["samples"], ref_mean), (nobs, ref_next), (info["raw_reward"], ref_raw), (info["penalty"], ref_pen), (rew, ref_rew)):
            assert rel_err_strict(got.cpu().numpy(), want) < 1e-4
    assert term.dtype == np.bool_ and term.shape == ref_term.shape
    if prec == "fp32" or near is None:
        assert np.array_equal(term, ref_term)
    else:   # lower precision may flip a mask only for rows within tol of a threshold; count them
        bad = np.flatnonzero(term[:, 0] != ref_term[:, 0])
        assert all(near(ref_next[i]) for i in bad), bad


@pytest.mark.parametrize("prec", precisions())
def test_step_matches_reference_golden(golden_dir, prec):
    files = sorted(glob.glob(os.path.join(golden_dir, "step_*.npz")))
    assert len(files) >= 6
    for f in files:
        g = np.load(f)
        env, S, A, seed = str(g["env"]), int(g["S"]), int(g["A"]), int(g["seed"])
        dyn, p = cuda_dynamics(S, A, seed, env, float(g["coef"]), precision=prec)
        members = p["elites"].numpy()[g["idx"]]
        out = dyn.step(torch.from_numpy(g["obs"]).cuda(), torch.from_numpy(g["act"]).cuda(), True, bool(g["use_trg"]),
                       eps=g["eps"], idx=members)
        _check(prec, out, g["next_obs"], g["reward"], g["raw_reward"], g["penalty"], g["mean"], g["terminal"])


@pytest.mark.parametrize("prec", precisions())
@pytest.mark.parametrize("B", [1, 63, 64, 65, 129, 1000])
def test_step_matches_oracle_ragged_batches(prec, B):
    S, A, seed = 17, 6, 41
    rng = np.random.default_rng(B)
    dyn, p = cuda_dynamics(S, A, seed, "walker2d", 5.0, precision=prec, h0=0.95, t3_gain=4.0)
    obs = (np.r_[1.25, np.zeros(S - 1)][None] + 0.2 * rng.standard_normal((B, S))).astype(np.float32)
    act = rng.uniform(-1, 1, (B, A)).astype(np.float32)
    eps = rng.standard_normal((7, B, S)).astype(np.float32)
    idx = rng.integers(0, 7, B)                      # any member id is legal for injected idx
    ref = M.step(p, torch.from_numpy(obs), torch.from_numpy(act), torch.from_numpy(eps), idx, 3, 5.0)
    out = dyn.step(torch.from_numpy(obs).cuda(), torch.from_numpy(act).cuda(), eps=eps, idx=idx)
    near = lambda x: min(abs(x[0] - 0.8), abs(x[0] - 2.0), abs(abs(x[1]) - 1.0)) < 1e-3   # noqa: E731
    _check(prec, out, ref["next_obs"].numpy(), ref["reward"].numpy(), ref["raw_reward"].numpy(), ref["penalty"].numpy(),
           ref["mean"].numpy(), ref["terminal"], near)
    if B >= 129:
        assert 0 < ref["terminal"].sum() < B          # the mask test is not vacuous


def test_step_flags_and_quirks():
    """use_penalty=False / coef=0 leave reward raw; use_trg switches the action encoder; empty batch."""
    S, A = 11, 3
    rng = np.random.default_rng(5)
    dyn, p = cuda_dynamics(S, A, 9, "hopper", 2.0)
    B = 70
    obs = torch.from_numpy((np.r_[1.25, np.zeros(S - 1)][None] + 0.1 * rng.standard_normal((B, S))).astype(np.float32)).cuda()
    act = torch.from_numpy(rng.uniform(-1, 1, (B, A)).astype(np.float32)).cuda()
    eps = rng.standard_normal((7, B, S)).astype(np.float32); idx = rng.integers(0, 5, B)
    n1, r1, t1, i1 = dyn.step(obs, act, True, True, eps=eps, idx=idx)
    n2, r2, t2, i2 = dyn.step(obs, act, False, True, eps=eps, idx=idx)
    assert torch.equal(n1, n2) and torch.equal(r2, i2["raw_reward"]) and not torch.equal(r1, r2)
    assert torch.allclose(r1, i1["raw_reward"] - 2.0 * i1["penalty"], rtol=1e-6, atol=1e-6)
    n3, _, _, i3 = dyn.step(obs, act, True, False, eps=eps, idx=idx)
    ref = M.step(p, obs.cpu(), act.cpu(), torch.from_numpy(eps), idx, 2, 2.0, True, False)
    assert rel_err(n3.cpu().numpy(), ref["next_obs"].numpy()) < 1e-4 and not torch.allclose(n3, n1)
    dyn._penalty_coef = 0.0
    _, r4, _, i4 = dyn.step(obs, act, True, True, eps=eps, idx=idx)
    assert torch.equal(r4, i4["raw_reward"])
    n0, r0, t0, i0 = dyn.step(obs[:0], act[:0])
    assert n0.shape == (0, S) and r0.shape == (0, 1) and t0.shape == (0, 1) and i0["samples"].shape == (7, 0, S)


def test_step_production_mode_is_philox_and_shard_invariant():
    """Without injected draws the kernel's noise/picks must equal the numpy Philox oracle, and a
    row's result must not depend on which shard (row0 offset / batch position) computed it."""
    from oracle.philox import rollout_noise, rollout_elite_slot
    from mobody_b200.dynamics import StepWorkspace
    S, A, B = 17, 6, 300
    rng = np.random.default_rng(8)
    dyn, p = cuda_dynamics(S, A, 3, "halfcheetah", 1.0)
    dyn.seed = 1234
    obs = torch.from_numpy(rng.standard_normal((B, S)).astype(np.float32)).cuda()
    act = torch.from_numpy(rng.uniform(-1, 1, (B, A)).astype(np.float32)).cuda()
    ws = StepWorkspace(B, S, A, obs.device)
    dyn.launch_step(obs, act, ws, step=7, row0=1000)
    torch.cuda.synchronize()
    rows = np.arange(1000, 1000 + B)
    eps_sel = rollout_noise(1234, 7, rows, S)                       # noise of the picked member
    members = p["elites"].numpy()[rollout_elite_slot(1234, 7, rows, 5)]
    eps = np.zeros((7, B, S), np.float32); eps[members, np.arange(B)] = eps_sel
    ref = M.step(p, obs.cpu(), act.cpu(), torch.from_numpy(eps), members, 1, 1.0)
    assert rel_err(ws.next_obs.cpu().numpy(), ref["next_obs"].numpy()) < 1e-4
    # second half computed as its own shard
    h = B // 2
    ws2 = StepWorkspace(B - h, S, A, obs.device)
    dyn.launch_step(obs[h:].contiguous(), act[h:].contiguous(), ws2, step=7, row0=1000 + h)
    torch.cuda.synchronize()
    assert torch.equal(ws2.next_obs, ws.next_obs[h:]) and torch.equal(ws2.reward, ws.reward[h:])


def test_policy_forward_matches_oracle():
    from helpers import cuda_agent
    for S, A, B in ((17, 6, 1), (11, 3, 257), (27, 8, 64)):
        ag, st = cuda_agent(S, A, 77)
        x = torch.randn(B, S, generator=torch.Generator().manual_seed(B))
        got = ag.policy(x.cuda()).cpu().numpy()
        want = M.policy_forward(st.policy, x, 1.0).numpy()
        assert got.shape == (B, A) and rel_err(got, want) < 1e-4
        a1 = ag.select_action(x[:1].numpy(), ag.policy)            # numpy in, numpy out, squeezed (mobody.py:138-144)
        assert isinstance(a1, np.ndarray) and a1.shape == (A,)


def test_single_pass_bf16_is_not_a_product_precision():
    """Single-pass bf16 (measured 1.6e-2) is outside north_star's bounds (1e-4; stated looser bound 5e-3): not selectable."""
    import mobody_b200 as mb
    from mobody_b200 import _ffi
    assert "bf16" not in _ffi.PREC and "bf16" not in _ffi.ENABLED_PRECISIONS
    with pytest.raises(ValueError, match="precision"):
        cuda_dynamics(17, 6, 1, "halfcheetah", 1.0, precision="bf16")
    assert mb.MOBODYEnsembleDynamics  # (import check)


def test_step_reads_packed_buffer_rows_in_place():
    """obs / act given as column views of packed replay-buffer rows (row stride RW): same result as dense copies."""
    import mobody_b200 as mb
    from mobody_b200.dynamics import StepWorkspace
    S, A, B = 17, 6, 300
    rng = np.random.default_rng(2)
    buf = mb.ReplayBuffer(S, A, "cuda", max_size=B)
    buf.convert_D4RL(dict(observations=rng.standard_normal((B, S)).astype(np.float32), actions=rng.uniform(-1, 1, (B, A)).astype(np.float32),
                          next_observations=rng.standard_normal((B, S)).astype(np.float32), rewards=np.zeros(B, np.float32),
                          terminals=np.zeros(B, bool)))
    rows = buf._rows
    for prec in precisions():
        dyn, _ = cuda_dynamics(S, A, 3, "halfcheetah", 1.0, precision=prec)
        w1, w2 = StepWorkspace(B, S, A, "cuda"), StepWorkspace(B, S, A, "cuda")
        dyn.launch_step(rows[:, :S], rows[:, S:S + A], w1, step=3)
        dyn.launch_step(rows[:, :S].contiguous(), rows[:, S:S + A].contiguous(), w2, step=3)
        torch.cuda.synchronize()
        assert torch.equal(w1.next_obs, w2.next_obs) and torch.equal(w1.reward, w2.reward) and torch.equal(w1.penalty, w2.penalty), prec
