"""Pin the oracle against vectors produced by the reference itself (oracle/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle as O
from oracle import mobody_oracle as M
from oracle.make_golden import HEALTHY, healthy

torch.set_num_threads(1)
FTOL = dict(rtol=1e-6, atol=1e-6)   # same torch build -> normally bit-identical; tolerate BLAS blocking differences


def _step_files(golden_dir):
    fs = sorted(glob.glob(os.path.join(golden_dir, "step_*.npz")))
    assert len(fs) >= 6
    return fs


def test_step_matches_reference(golden_dir):
    for f in _step_files(golden_dir):
        g = np.load(f)
        env, S, A, seed = str(g["env"]), int(g["S"]), int(g["A"]), int(g["seed"])
        p = M.make_dynamics_params(S, A, seed, healthy_state=HEALTHY[env](S))
        kind = M.termination_kind(env)
        elites = p["elites"].numpy()
        r = M.step(p, torch.from_numpy(g["obs"]), torch.from_numpy(g["act"]), torch.from_numpy(g["eps"]),
                   elites[g["idx"]], kind, float(g["coef"]), True, bool(g["use_trg"]))
        np.testing.assert_allclose(r["mean"].numpy(), g["mean"], **FTOL)
        np.testing.assert_allclose(r["next_obs"].numpy(), g["next_obs"], **FTOL)
        np.testing.assert_allclose(r["raw_reward"].numpy(), g["raw_reward"], **FTOL)
        np.testing.assert_allclose(r["penalty"].numpy(), g["penalty"], **FTOL)
        np.testing.assert_allclose(r["reward"].numpy(), g["reward"], **FTOL)
        assert r["terminal"].dtype == np.bool_ and r["terminal"].shape == g["terminal"].shape
        assert np.array_equal(r["terminal"], g["terminal"])


def test_quirks_visible_in_golden(golden_dir):
    """A.5: penalty ignores the last state dim; reward averages all 7 members."""
    g = np.load(_step_files(golden_dir)[0])
    mean = g["mean"]
    d = mean[..., :-1] - mean[..., :-1].mean(0)
    pen = np.sqrt((d ** 2).sum(-1)).max(0)
    np.testing.assert_allclose(pen, g["penalty"][:, 0], rtol=1e-5)
    d_full = mean - mean.mean(0)
    assert not np.allclose(np.sqrt((d_full ** 2).sum(-1)).max(0), g["penalty"][:, 0], rtol=1e-5)
    np.testing.assert_allclose(g["reward"], g["raw_reward"] - g["coef"] * g["penalty"], rtol=1e-6, atol=1e-6)


def test_termination_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "termination.npz"))
    for env in ("halfcheetah", "hopper", "walker2d", "ant"):
        with np.errstate(invalid="ignore"):
            done = M.termination(M.termination_kind(env), g[env + "_x"])
        assert done.dtype == np.bool_ and np.array_equal(done, g[env + "_done"]), env
        assert 0 < done.sum() < done.size
    ref_name = {"termination_fn_halfcheetahveljump": "never", "termination_fn_halfcheetah": "halfcheetah",
                "termination_fn_hopper": "hopper", "termination_fn_antangle": "ant", "termination_fn_ant": "ant",
                "termination_fn_walker2d": "walker2d", "termination_fn_pendulum": "never",
                "termination_fn_humanoid": "humanoid", "termination_fn_pen": "pen", "terminaltion_fn_door": "never"}
    for name, fn in zip(g["dispatch_names"], g["dispatch_fn"]):
        assert M.termination_kind(str(name)) == M.TERM_KINDS[ref_name[str(fn)]], name
    with pytest.raises(TypeError):
        M.termination_kind("unknown-task")


@pytest.mark.parametrize("name", ["rollout_walker2d_T3.npz", "rollout_hopper_T5.npz"])
def test_rollout_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    env, S, A, seed, T = str(g["env"]), int(g["S"]), int(g["A"]), int(g["seed"]), int(g["T"])
    p = M.make_dynamics_params(S, A, seed, healthy_state=healthy(env, S, float(g["h0"])), t3_gain=float(g["t3_gain"]))
    ag = M.AgentState(S, A, seed)
    elites = p["elites"].numpy()
    eps, idx = g["eps"], g["idx"]
    tr, info = M.rollout(p, ag.policy, 1.0, torch.from_numpy(g["obs"]), T,
                         [(lambda n, t=t: torch.from_numpy(eps[t][:, :n].copy())) for t in range(T)],
                         [(lambda n, t=t: elites[idx[t][:n]]) for t in range(T)],
                         M.termination_kind(env), float(g["coef"]), float(g["env_filter"]), True)
    assert info["num_transitions"] == int(g["num_transitions"])
    assert info["num_transitions"] < int(g["B"]) * T          # some rows terminated
    assert len(tr["obss"]) < info["num_transitions"]          # the penalty filter dropped rows
    np.testing.assert_allclose(info["reward_mean"], float(g["reward_mean"]), rtol=1e-5)
    for k in ("obss", "next_obss", "actions", "rewards", "terminals", "penalty"):
        assert tr[k].shape == g["out_" + k].shape, k
        np.testing.assert_allclose(tr[k].numpy(), g["out_" + k], err_msg=k, **FTOL)
    assert np.array_equal(tr["terminals"].numpy(), g["out_terminals"])


def test_rollout_zero_length():
    assert M.rollout(None, None, 1.0, None, 0, [], [], 0, 0.0) == (None, None)


def test_ring_buffer_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "buffer.npz"))
    buf = M.RingBuffer(int(g["S"]), int(g["A"]), int(g["cap"]))
    buf.add_batch(None)
    for i in range(int(g["n_batches"])):
        buf.add_batch({k: torch.from_numpy(g[f"b{i}_{k}"]) for k in ("obss", "next_obss", "actions", "rewards", "terminals")})
        assert (buf.ptr, buf.size) == (int(g[f"after{i}_ptr"]), int(g[f"after{i}_size"]))
        for f in ("state", "action", "next_state", "reward", "not_done"):
            assert np.array_equal(getattr(buf, f).numpy(), g[f"after{i}_{f}"]), (i, f)
    smp = buf.gather(g["ind"])
    for f, v in zip(("state", "action", "next_state", "reward", "not_done"), smp):
        assert np.array_equal(v.numpy(), g["sample_" + f])


@pytest.mark.parametrize("name", ["train_S17A6_B32.npz", "train_S11A3_B16.npz"])
def test_train_step_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    S, A, B, seed, n_steps = (int(g[k]) for k in ("S", "A", "B", "seed", "n_steps"))
    ag = M.AgentState(S, A, seed)
    bufs = {}
    for nm in ("src", "tar", "fake"):
        n = int(g["n_" + nm])
        b = M.RingBuffer(S, A, n)
        for f in ("state", "action", "next_state", "reward", "not_done"):
            getattr(b, f)[:] = torch.from_numpy(g[f"{nm}_{f}"])
        b.size = n
        bufs[nm] = b
    cfg = dict(gamma=0.99, tau=0.005, actor_lr=3e-4, critic_lr=3e-4, weight=2.5, bc_coef=1.0, max_action=1.0)
    for it in range(n_steps):
        parts = [bufs[nm].gather(g[f"ind{3 * it + j}"]) for j, nm in enumerate(("src", "tar", "fake"))]
        batch = tuple(torch.cat([p[c] for p in parts], 0) for c in range(5))
        out = M.train_step(ag, batch, 2 * B, cfg)
        assert np.isfinite(out["q_loss"]) and np.isfinite(out["pi_loss"])
    for grp, sd in (("pi", ag.policy), ("q", ag.q), ("qt", ag.q_target)):
        for k, v in sd.items():
            flat = v.numpy().reshape(-1)
            np.testing.assert_allclose(flat[::37], g[f"post_{grp}_{k}_sub"], rtol=2e-5, atol=2e-7, err_msg=f"{grp}.{k}")
            np.testing.assert_allclose(flat.astype(np.float64).sum(), float(g[f"post_{grp}_{k}_sum"]), rtol=1e-4, atol=1e-5)


def classifier_batch(g, it):
    """The concatenated + permuted batch update_classifier builds (mobody.py:147-167) from the scripted draws."""
    B = int(g["B"])
    si, ti, pm = g[f"ind{2 * it}"], g[f"ind{2 * it + 1}"], g[f"perm{it}"]
    cat = lambda f: np.concatenate([g["src_" + f][si], g["tar_" + f][ti]], 0)[pm]          # noqa: E731
    label = np.concatenate([np.zeros(B, np.int64), np.ones(B, np.int64)])[pm]
    return cat("state"), cat("action"), cat("next_state"), label, g[f"noise{2 * it}"], g[f"noise{2 * it + 1}"]


def test_classifier_update_and_relabel_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "classifier_S17A6_B32.npz"))
    S, A, seed, n_steps = (int(g[k]) for k in ("S", "A", "seed", "n_steps"))
    cl = M.ClassifierState(S, A, seed)
    for it in range(n_steps):
        s, a, s2, label, nz_sas, nz_sa = (torch.from_numpy(x) for x in classifier_batch(g, it))
        lsa, lsas = M.classifier_update(cl, s, a, s2, label, nz_sas, nz_sa, float(g["std"]), float(g["lr"]))
        np.testing.assert_allclose([lsa, lsas], g["losses"][it], rtol=2e-6)
    for k, v in cl.params.items():
        flat = v.numpy().reshape(-1)
        np.testing.assert_allclose(flat[::29], g[f"post_{k}_sub"], rtol=2e-5, atol=2e-7, err_msg=k)
        np.testing.assert_allclose(flat.astype(np.float64).sum(), float(g[f"post_{k}_sum"]), rtol=1e-4, atol=1e-5)
    pen = M.dara_reward_penalty(cl.params, *(torch.from_numpy(g["src_" + f]) for f in ("state", "action", "next_state")))
    np.testing.assert_allclose(pen.numpy(), g["reward_penalty"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(g["src_reward"] + float(g["penalty_coef"]) * pen.numpy(), g["new_reward"], rtol=1e-6, atol=1e-6)


def test_philox_known_answers():
    """Random123 published KATs for philox4x32-10."""
    z = O.philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = O.philox4x32_10(np.full(4, 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))
    assert [hex(int(v)) for v in f] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    p = O.philox4x32_10(np.array([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], np.uint32),
                        np.array([0xa4093822, 0x299f31d0], np.uint32))
    assert [hex(int(v)) for v in p] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_normals_are_normal():
    from oracle.philox import rollout_noise, rollout_elite_slot, buffer_indices
    n = rollout_noise(0, 0, np.arange(20000), 17)
    assert n.shape == (20000, 17) and abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01
    s = rollout_elite_slot(0, 0, np.arange(50000), 5)
    assert s.min() == 0 and s.max() == 4 and np.all(np.abs(np.bincount(s) / 50000 - 0.2) < 0.01)
    i = buffer_indices(1, 2, 10000, 777)
    assert i.min() >= 0 and i.max() < 777


def fit_replay(g, step_fn):
    """Drive ``step_fn(call, use_trg, batch, eps_latent, eps_next, t)`` through the calls of a dynfit fixture; t = per-layer
    Adam step counts AFTER the call.  Shared by the oracle test here and the CUDA test (tests/test_gpu_dynfit.py)."""
    counts = {}
    for c, use_trg in enumerate(g["calls"]):
        for n in M.fit_trained_layers(bool(use_trg)):
            counts[n] = counts.get(n, 0) + 1
        batch = [g[f"c{c}_{k}"] for k in ("obs", "act", "next_obs", "reward")]
        step_fn(c, bool(use_trg), batch, g[f"c{c}_eps_latent"], g[f"c{c}_eps_next"], dict(counts))
    return counts


@pytest.mark.parametrize("name", ["dynfit_S11A3_B24", "dynfit_S17A6_B40"])
def test_dynamics_fitting_step_matches_reference(golden_dir, name):
    """oracle.fit_step == the reference's learn() (mobody_dynamics.py:594-653) under the same scripted randn_like draws:
    returned loss scalars per call, Adam state and parameters after the last call (sampled elements)."""
    from oracle.make_golden import fit_sample_index
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    S, A = int(g["S"]), int(g["A"])
    p = M.make_dynamics_params(S, A, int(g["seed"]), healthy_state=HEALTHY["hopper" if S == 11 else "walker2d"](S))
    m, v, first_m = {}, {}, {}

    def step(c, use_trg, batch, el, en, t):
        for n in M.fit_trained_layers(use_trg):
            for sfx in (".weight", ".bias"):
                m.setdefault(n + sfx, torch.zeros_like(p[n + sfx])); v.setdefault(n + sfx, torch.zeros_like(p[n + sfx]))
        losses, _ = M.fit_step(p, m, v, t, *(torch.from_numpy(x) for x in batch), torch.from_numpy(el), torch.from_numpy(en), use_trg)
        np.testing.assert_allclose(losses[:5], g[f"c{c}_losses"], rtol=2e-5)
        if c == 0:
            first_m.update({k: x.clone() for k, x in m.items()})
    counts = fit_replay(g, step)
    for n in M.dynamics_layer_shapes(S, A):
        for sfx in (".weight", ".bias"):
            k = n + sfx
            idx = fit_sample_index(p[k].numel())
            assert bool(g["has_state_" + k]) == (k in m)
            np.testing.assert_allclose(p[k].numpy().ravel()[idx], g["p_" + k], rtol=1e-5, atol=2e-6)
            if k in m:
                assert int(g["t_" + k]) == counts[n]
                np.testing.assert_allclose(m[k].numpy().ravel()[idx], g["m_" + k], rtol=2e-4, atol=1e-7)
                np.testing.assert_allclose(v[k].numpy().ravel()[idx], g["v_" + k], rtol=4e-4, atol=1e-12)
            if "first_m_" + k in g.files:
                np.testing.assert_allclose(first_m[k].numpy().ravel()[idx], g["first_m_" + k], rtol=2e-4, atol=1e-7)
