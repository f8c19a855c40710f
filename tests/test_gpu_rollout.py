"""GPU parity of the fused multi-step rollout against reference-generated goldens."""
import os

import numpy as np
import pytest
import torch

from helpers import cuda_agent, cuda_dynamics, rel_err, rel_err_strict
from oracle import mobody_oracle as M

pytestmark = pytest.mark.gpu


def _setup(g, precision="fp32"):
    env, S, A, seed = str(g["env"]), int(g["S"]), int(g["A"]), int(g["seed"])
    dyn, p = cuda_dynamics(S, A, seed, env, float(g["coef"]), precision=precision, h0=float(g["h0"]), t3_gain=float(g["t3_gain"]))
    ag, _ = cuda_agent(S, A, seed, env_filter=float(g["env_filter"]))
    ag.dynamics = dyn
    return ag, p


@pytest.mark.parametrize("precision", ["fp32", "bf16x2"])
@pytest.mark.parametrize("name", ["rollout_walker2d_T3.npz", "rollout_hopper_T5.npz"])
def test_rollout_matches_reference_golden(golden_dir, name, precision, capsys):
    """Multi-step rollouts with real terminations and a penalty filter that drops rows, in the fp32 CUDA-core mode AND in
    the benchmarked tensor-core mode (bf16 hi+lo split): same row count, same order, masks bit-exact, values <= 1e-4."""
    g = np.load(os.path.join(golden_dir, name))
    ag, p = _setup(g, precision)
    members = p["elites"].numpy()[g["idx"]]
    tr, info = ag.rollout(torch.from_numpy(g["obs"]).cuda(), int(g["T"]), True, eps=g["eps"], idx=members)
    assert "filtered rollout" in capsys.readouterr().out                 # quirk 8
    assert info["num_transitions"] == int(g["num_transitions"])
    assert abs(info["reward_mean"] - float(g["reward_mean"])) < 1e-4 * abs(float(g["reward_mean"])) + 1e-6
    for k in ("obss", "next_obss", "actions", "rewards", "terminals", "penalty"):
        assert not tr[k].is_cuda and tr[k].shape == g["out_" + k].shape, k   # CPU tensors, post-filter row count and order
        assert rel_err(tr[k].numpy(), g["out_" + k]) < 1e-4, k
        assert rel_err_strict(tr[k].numpy(), g["out_" + k]) < 1e-4, k
    assert np.array_equal(tr["terminals"].numpy(), g["out_terminals"])      # masks bit-exact
    assert tr["rewards"].shape[1] == 1 and tr["terminals"].dtype == torch.float32


@pytest.mark.parametrize("name", ["rollout_walker2d_T3.npz", "rollout_hopper_T5.npz"])
def test_rollout_fp16_mode_within_its_stated_bound(golden_dir, name):
    """Single-pass fp16 mode (stated looser bound 5e-3): a row within rounding of a termination / filter threshold may land
    on the other side, which changes which rows later steps hold.  Rows are therefore lined up by their start state (the
    first-step rows carry it bit-exactly): first-step transitions present in both results agree within 5e-3, and the
    transition count differs from the reference's by a few rows at most."""
    g = np.load(os.path.join(golden_dir, name))
    ag, p = _setup(g, "fp16")
    members = p["elites"].numpy()[g["idx"]]
    tr, info = ag.rollout(torch.from_numpy(g["obs"]).cuda(), int(g["T"]), True, eps=g["eps"], idx=members)
    assert abs(info["num_transitions"] - int(g["num_transitions"])) <= 0.05 * int(g["num_transitions"])
    key = {row.tobytes(): i for i, row in enumerate(g["obs"])}
    got = {key[r.tobytes()]: j for j, r in enumerate(tr["obss"].numpy()) if r.tobytes() in key}
    want = {key[r.tobytes()]: j for j, r in enumerate(g["out_obss"]) if r.tobytes() in key}
    both = sorted(set(got) & set(want))
    assert len(both) >= 0.9 * len(want) > 10
    gi, wi = [got[i] for i in both], [want[i] for i in both]
    for k in ("next_obss", "actions", "rewards", "penalty"):
        assert rel_err(tr[k].numpy()[gi], g["out_" + k][wi]) < 5e-3, k


def test_rollout_edge_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "rollout_walker2d_T3.npz"))
    ag, p = _setup(g)
    assert ag.rollout(torch.from_numpy(g["obs"]).cuda(), 0) == (None, None)          # quirk 9
    ag.fake_replay_buffer.add_batch(None)
    # all rows terminate at step 0 -> later steps contribute nothing (early break, quirk 11)
    bad = torch.zeros(10, int(g["S"]), device="cuda")
    ag.dynamics.model.transition3.bias.data.zero_()
    ag.dynamics.model.transition3.weight.data.zero_()
    ag.config["filter_bad_rollout"] = 0
    tr, info = ag.rollout(bad, 4)
    assert info["num_transitions"] == 10 and tr["obss"].shape[0] == 10 and float(tr["terminals"].sum()) == 10.0
    # single-row batch (squeeze/reshape quirk 10)
    tr1, info1 = ag.rollout(bad[:1], 2)
    assert tr1["actions"].shape == (1, int(g["A"]))


def test_rollout_production_mode_shard_invariant(golden_dir):
    """Philox draws keyed on the global row id: rolling two halves separately gives the same multiset."""
    g = np.load(os.path.join(golden_dir, "rollout_walker2d_T3.npz"))
    ag, _ = _setup(g)
    ag.config["filter_bad_rollout"] = 0
    obs = torch.from_numpy(g["obs"]).cuda()
    full, fi = ag.rollout_device(obs, 3, step0=0)
    h = obs.shape[0] // 2
    a, ai = ag.rollout_device(obs[:h], 3, row0=0, step0=0)
    b, bi = ag.rollout_device(obs[h:], 3, row0=h, step0=0)
    assert fi["num_transitions"] == ai["num_transitions"] + bi["num_transitions"]
    assert fi["rows_per_step"][0] == obs.shape[0] and 0 < fi["rows_per_step"][2] < obs.shape[0]

    def canon(d):
        m = torch.cat([d[k] for k in ("obss", "actions", "next_obss", "rewards", "terminals")], 1).cpu().numpy()
        return m[np.lexsort(m.T[::-1])]
    both = {k: torch.cat([a[k], b[k]], 0) for k in a}
    assert np.array_equal(canon(full), canon(both))


@pytest.mark.parametrize("precision", ["bf16x2", "fp32"])
def test_pipelined_host_rollout_equals_single_pass(precision, capsys):
    """rollout() with HOST start states pipelines one-step rollouts in chunks on side streams; the returned CPU
    tensors must be bit-identical (values and row order) to the unchunked device rollout, filter included."""
    S, A = 11, 3
    dyn, _ = cuda_dynamics(S, A, 3, "hopper", 5.0, precision=precision)
    ag, _ = cuda_agent(S, A, 3, env_filter=1e9)
    ag.dynamics = dyn
    ag.PIPE_ROWS = 4096 if precision == "fp32" else ag.PIPE_ROWS
    B = 2 * ag.PIPE_ROWS + 1500                             # three chunks (a short tail is folded only below 1/4 chunk)
    rng = np.random.default_rng(5)
    obs = (np.r_[1.25, np.zeros(S - 1)][None] + 0.2 * rng.standard_normal((B, S))).astype(np.float32)
    host_in = torch.from_numpy(obs).pin_memory()
    probe, _ = ag.rollout_device(torch.from_numpy(obs).cuda(), 1)
    ag.config["env_filter"] = float(probe["penalty"].median())     # a filter that really drops rows
    draw = dyn._draw
    tr, info = ag.rollout(host_in, 1)                            # consumes Philox draw `draw`
    ref, ri = ag.rollout_device(torch.from_numpy(obs).cuda(), 1, step0=draw)
    assert 0 < ri["kept"] < B and info["num_transitions"] == ri["num_transitions"] == B
    assert abs(info["reward_mean"] - ri["reward_mean"]) < 1e-5 * abs(ri["reward_mean"]) + 1e-7
    for k in ("obss", "actions", "next_obss", "rewards", "terminals", "penalty"):
        assert not tr[k].is_cuda and tuple(tr[k].shape) == tuple(ref[k].shape), k
        assert torch.equal(tr[k], ref[k].cpu()), k


def test_full_size_rollout_properties_and_sampled_oracle_parity():
    """BASELINE configs[1] size (100 000 start states, obs 17 / act 6, one step, production Philox draws): size-independent
    properties — run-to-run bit determinism, independence from how the rows are sharded (three shards with their global
    row offsets == one pass), output ranges and counts — and parity of a random sample of rows against the CPU oracle fed
    the same Philox draws (noise from oracle.philox, member picks from the elite slots)."""
    from oracle.philox import rollout_noise, rollout_elite_slot
    S, A, B = 17, 6, 100_000
    dyn, p = cuda_dynamics(S, A, 1, "halfcheetah", 5.0, precision="bf16x2")
    dyn.seed = 77
    ag, st = cuda_agent(S, A, 1, env_filter=1e9)
    ag.dynamics = dyn
    rng = np.random.default_rng(100)
    obs_np = (0.3 * rng.standard_normal((B, S))).astype(np.float32)
    obs = torch.from_numpy(obs_np).cuda()
    a, ia = ag.rollout_device(obs, 1, step0=0)
    pa = ia["packed"][:ia["kept"]].clone()
    b, ib = ag.rollout_device(obs, 1, step0=0)
    assert ia["kept"] == ib["kept"] == B and ia["num_transitions"] == B
    assert torch.equal(pa, ib["packed"][:B])                                   # bit-deterministic
    cuts = [0, 33_333, 70_001, B]
    parts = [ag.rollout_device(obs[lo:hi].contiguous(), 1, row0=lo, step0=0)[1] for lo, hi in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat([q["packed"][:q["kept"]] for q in parts], 0), pa)    # shard invariant, order preserved
    term, pen = pa[:, 2 * S + A + 1], pa[:, 2 * S + A + 2]
    assert bool(((term == 0) | (term == 1)).all()) and bool((pen >= 0).all()) and bool(torch.isfinite(pa).all())
    assert torch.equal(pa[:, :S], obs)                                         # the obs columns are the start states, in order
    rows = np.sort(rng.choice(B, 256, replace=False))
    eps_sel = rollout_noise(77, 0, rows, S)
    members = p["elites"].numpy()[rollout_elite_slot(77, 0, rows, 5)]
    eps = np.zeros((7, len(rows), S), np.float32); eps[members, np.arange(len(rows))] = eps_sel
    so = torch.from_numpy(obs_np[rows])
    act = M.policy_forward(st.policy, so, 1.0)
    ref = M.step(p, so, act, torch.from_numpy(eps), members, 1, 5.0)
    got = pa[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert rel_err(got[:, S:S + A], act.numpy()) < 1e-4
    assert rel_err(got[:, S + A:2 * S + A], ref["next_obs"].numpy()) < 1e-4
    assert rel_err(got[:, 2 * S + A:2 * S + A + 1], ref["reward"].numpy()) < 1e-4
    assert rel_err(got[:, 2 * S + A + 2:2 * S + A + 3], ref["penalty"].numpy()) < 1e-4
    assert np.array_equal(got[:, 2 * S + A + 1] != 0, ref["terminal"][:, 0])


@pytest.mark.parametrize("precision", ["bf16x2", "fp16"])
def test_packed_weight_image_follows_data_writes(precision):
    """The tensor-core kernels read a packed 16-bit image of the weights.  The reference restores weights with
    ``param.data.copy_()`` (MOBODYModule.load_save, mobody_module.py:407-408), which bumps no autograd version counter: the
    image must follow such writes (device-side checksum), and must NOT be rebuilt when nothing changed."""
    S, A = 17, 6
    dyn, p = cuda_dynamics(S, A, 7, "halfcheetah", 1.0, precision=precision)
    ag, _ = cuda_agent(S, A, 7, env_filter=1e9)
    ag.dynamics = dyn
    obs = torch.randn(300, S, device="cuda", generator=torch.Generator("cuda").manual_seed(2)) * 0.3
    base, _ = ag.rollout_device(obs, 1, step0=0)
    n0, a0 = base["next_obss"].clone(), base["actions"].clone()
    again, _ = ag.rollout_device(obs, 1, step0=0)
    assert torch.equal(again["next_obss"], n0)
    blob, state = dyn._packs[("dyn", precision)]
    ck = state.clone()
    # (1) .data write to the dynamics weights (version counter untouched)
    w = dyn.model.transition3.bias
    v0 = w._version
    w.data.copy_(w.data + 0.25)
    assert w._version == v0
    moved, _ = ag.rollout_device(obs, 1, step0=0)
    assert float((moved["next_obss"] - n0).abs().min()) > 0.2        # every next-state component moved by ~0.25
    assert not torch.equal(dyn._packs[("dyn", precision)][1][:1], ck[:1])      # image re-packed: stored checksum changed
    # (2) load_save-style restore of one member from the saved copy
    dyn.model.transition3.bias.data.copy_(w.data - 0.25)
    back, _ = ag.rollout_device(obs, 1, step0=0)
    assert torch.allclose(back["next_obss"], n0, atol=1e-5)
    # (3) policy weights written through .data (a checkpoint restore): actions change
    ag.policy.network.network[4].bias.data.add_(0.3)
    pol, _ = ag.rollout_device(obs, 1, step0=0)
    assert float((pol["actions"] - a0).abs().max()) > 0.05
    # (4) an optimiser step of the fused train kernel changes the policy in place: the next rollout sees it
    from mobody_b200 import _ffi
    RW = _ffi.lib().mobody_row_width(S, A)
    rows = torch.randn(320, RW, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    ag.train_on_rows(rows, 256)
    trained, _ = ag.rollout_device(obs, 1, step0=0)
    assert not torch.equal(trained["actions"], pol["actions"])


def test_rollout_results_stay_valid_while_referenced():
    """rollout() returns CPU tensors the caller owns (mobody.py:641-657): keeping several results alive must not let a
    later call overwrite them; dropped results let their pinned staging slab be reused."""
    S, A = 11, 3
    dyn, _ = cuda_dynamics(S, A, 3, "hopper", 1.0)
    ag, _ = cuda_agent(S, A, 3, env_filter=1e9)
    ag.dynamics = dyn
    rng = np.random.default_rng(0)
    keep, copies = [], []
    for i in range(4):
        obs = torch.from_numpy((np.r_[1.25, np.zeros(S - 1)][None] + 0.1 * rng.standard_normal((64 + i, S))).astype(np.float32)).cuda()
        tr, _ = ag.rollout(obs, 1)
        keep.append(tr); copies.append({k: v.clone() for k, v in tr.items()})
    for tr, cp in zip(keep, copies):
        for k in tr:
            assert torch.equal(tr[k], cp[k]), k
    n_slabs = len(ag._host_slabs)
    del keep, tr
    for _ in range(4):
        ag.rollout(obs, 1)
    assert len(ag._host_slabs) <= max(n_slabs, 2)                   # idle slabs are reused, not accumulated
