"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no
/root/reference):   python -m oracle.make_golden

RNG injection: the reference draws noise with torch.normal (mobody_dynamics.py:220),
member indices with np.random.choice (mobody_module.py:356) and buffer indices with
np.random.randint (utils.py:128).  We patch those three callables around the reference
call so the recorded eps / idx / ind are exactly what the reference consumed.
Weights come from oracle.philox.recipe_fill (named by seed), so fixtures stay small.
"""
import os
import sys
import contextlib

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

from oracle import mobody_oracle as O  # noqa: E402

HEALTHY = {  # state means inside each env's healthy set
    "walker2d": lambda S: np.r_[1.25, 0.0, np.zeros(S - 2)],
    "hopper": lambda S: np.r_[1.25, 0.0, np.zeros(S - 2)],
    "halfcheetah": lambda S: np.zeros(S),
    "ant": lambda S: np.r_[0.6, np.zeros(S - 1)],
}
ENV_NAME = {"walker2d": "walker2d-medium-v2", "hopper": "hopper-medium-v2",
            "halfcheetah": "halfcheetah-medium-v2", "ant": "ant-medium-v2"}


def _import_reference():
    sys.path.insert(0, REF)
    from algo.dynamics.mobody_module import MOBODYModule
    from algo.dynamics.mobody_dynamics import MOBODYEnsembleDynamics
    from algo.mb_utils.terminal_funs import get_termination_fn
    from algo.offline_offline.mobody import MOBODY
    from algo import utils as ref_utils
    return MOBODYModule, MOBODYEnsembleDynamics, get_termination_fn, MOBODY, ref_utils


class Inject:
    """Patch torch.normal / np.random.choice / np.random.randint with scripted draws."""
    def __init__(self, eps_fn=None, idx_fn=None, ind_list=None):
        self.eps_fn, self.idx_fn, self.ind_list = eps_fn, idx_fn, list(ind_list or [])
        self.calls = {"normal": 0, "choice": 0, "randint": 0}

    def __enter__(self):
        self._n, self._c, self._r = torch.normal, np.random.choice, np.random.randint
        me = self

        def normal(mean=None, std=None, **kw):
            eps = me.eps_fn(me.calls["normal"], tuple(std.shape)); me.calls["normal"] += 1
            return mean + eps * std

        def choice(a, size=None, **kw):
            idx = me.idx_fn(me.calls["choice"], int(size)); me.calls["choice"] += 1
            return np.asarray(a)[idx]          # idx are *slots* into elites

        def randint(low, high=None, size=None, **kw):
            ind = me.ind_list[me.calls["randint"]]; me.calls["randint"] += 1
            assert len(ind) == size and ind.max() < high
            return ind
        torch.normal, np.random.choice, np.random.randint = normal, choice, randint
        return self

    def __exit__(self, *a):
        torch.normal, np.random.choice, np.random.randint = self._n, self._c, self._r


def healthy(env, S, h0=None):
    hs = HEALTHY[env](S).astype(np.float64)
    if h0 is not None:
        hs[0] = h0
    return hs


def build_reference_dynamics(S, A, seed, env, coef, h0=None, t3_gain=1.0):
    MOBODYModule, MOBODYEnsembleDynamics, get_termination_fn, _, _ = _import_reference()
    cfg = {"mopo": 0, "latent_reward": 0, "encoder_loss_coef": 1, "domain_loss_coef": 0, "cycle_loss_coef": 0}
    with contextlib.redirect_stdout(None):
        m = MOBODYModule(S, A, hidden_dims=256, num_ensemble=7, num_elites=5, device="cpu", config=cfg)
    p = O.make_dynamics_params(S, A, seed, healthy_state=healthy(env, S, h0), t3_gain=t3_gain)
    with torch.no_grad():
        for name in O.dynamics_layer_shapes(S, A):
            getattr(m, name).weight.copy_(p[name + ".weight"])
            getattr(m, name).bias.copy_(p[name + ".bias"])
    dyn = MOBODYEnsembleDynamics(cfg, m, None, None, get_termination_fn(ENV_NAME[env]), penalty_coef=coef)
    return dyn, p


def synth_obs(env, S, B, rng, spread=0.15):
    return (HEALTHY[env](S)[None, :] + spread * rng.standard_normal((B, S))).astype(np.float32)


def gen_step(env, S, A, B, seed, coef, use_trg=True):
    rng = np.random.default_rng(seed)
    dyn, _ = build_reference_dynamics(S, A, seed, env, coef)
    obs = synth_obs(env, S, B, rng); act = rng.uniform(-1, 1, (B, A)).astype(np.float32)
    eps = rng.standard_normal((7, B, S)).astype(np.float32)
    slot = rng.integers(0, 5, B)
    with Inject(lambda c, shp: torch.from_numpy(eps), lambda c, n: slot):
        nobs, rew, term, info = dyn.step(torch.from_numpy(obs), torch.from_numpy(act), True, use_trg)
    return dict(env=env, S=S, A=A, seed=seed, coef=np.float32(coef), use_trg=use_trg, obs=obs, act=act, eps=eps,
                idx=slot.astype(np.int64), next_obs=nobs.numpy(), reward=rew.numpy(), terminal=term,
                mean=info["samples"].numpy(), raw_reward=info["raw_reward"].numpy(), penalty=info["penalty"].numpy())


def gen_termination():
    _, _, get_termination_fn, _, _ = _import_reference()
    rng = np.random.default_rng(7)
    out = {}
    for env, S in (("halfcheetah", 17), ("hopper", 11), ("walker2d", 17), ("ant", 27)):
        x = synth_obs(env, S, 256, rng, 0.4)
        # edge rows: exact thresholds, +-inf, nan, big magnitudes, far-negative dims
        edges = [(0, 0, 0.7), (1, 0, 0.8), (2, 0, 2.0), (3, 0, 0.2), (4, 0, 1.0), (5, 1, 0.2), (6, 1, -0.2),
                 (7, 1, 1.0), (8, 1, -1.0), (9, 2, 100.0), (10, 2, -100.0), (11, 3, np.inf), (12, 3, -np.inf),
                 (13, 4, np.nan), (14, 2, -500.0), (15, 0, np.nextafter(np.float32(0.8), np.float32(1))),
                 (16, 0, np.nextafter(np.float32(0.7), np.float32(1))), (17, S - 1, 150.0), (18, 0, np.nan)]
        for r, c, v in edges:
            x[r, c] = v
        fn = get_termination_fn(ENV_NAME[env])
        with np.errstate(invalid="ignore"):
            out[env + "_x"] = x
            out[env + "_done"] = fn(x, x[:, :1], x)
    names = ["halfcheetahvel-x", "halfcheetah-medium-v2", "hopper-medium-v2", "antangle-v0", "ant-medium-v2",
             "antmaze-umaze-v0", "walker2d-medium-v2", "pendulum", "humanoid-v2", "pen-human-v0", "door-human-v0"]
    out["dispatch_names"] = np.array(names)
    out["dispatch_fn"] = np.array([get_termination_fn(n).__name__ for n in names])
    return out


def build_reference_agent(S, A, seed, overrides=None):
    _, _, _, MOBODY, _ = _import_reference()
    cfg = dict(state_dim=S, action_dim=A, max_action=1.0, hidden_sizes=256, gamma=0.99, tau=0.005,
               update_interval=2, actor_lr=3e-4, critic_lr=3e-4, gaussian_noise_std=1.0, weight=2.5,
               penalty_type="none", penalty_coef=0.1, mopo=0, latent_reward=0, encoder_loss_coef=1,
               domain_loss_coef=0, cycle_loss_coef=0, advantage=0, q_weighted=1, scale_Q=1, bc_coef=1.0,
               fake_batch_scale=0.5, src_ratio=1, trg_ratio=1, filter_bad_rollout=1, env_filter=10.0,
               src_rollout_length=1, trg_rollout_length=1, use_src_sa_to_get_target_next_state=1,
               rollout_from_src=0, penalize_fake=0)
    cfg.update(overrides or {})
    pol = MOBODY(cfg, torch.device("cpu"))
    ag = O.AgentState(S, A, seed)
    pol.policy.load_state_dict(ag.policy)
    pol.q_funcs.load_state_dict(ag.q)
    pol.target_q_funcs.load_state_dict(ag.q_target)
    return pol, ag, cfg


def gen_rollout(env, S, A, B, T, seed, coef, env_filter, h0, t3_gain):
    rng = np.random.default_rng(seed)
    dyn, _ = build_reference_dynamics(S, A, seed, env, coef, h0, t3_gain)
    pol, _, _ = build_reference_agent(S, A, seed, dict(env_filter=env_filter))
    pol.dynamics = dyn
    obs = synth_obs(env, S, B, rng, 0.2)
    eps_full = rng.standard_normal((T, 7, B, S)).astype(np.float32)   # step t uses the first B_t rows
    slot_full = rng.integers(0, 5, (T, B))
    with Inject(lambda c, shp: torch.from_numpy(eps_full[c][:, :shp[1]].copy()), lambda c, n: slot_full[c][:n]), \
            contextlib.redirect_stdout(None):
        tr, info = pol.rollout(torch.from_numpy(obs), T, True)
    d = dict(env=env, S=S, A=A, B=B, T=T, seed=seed, coef=np.float32(coef), env_filter=np.float32(env_filter),
             h0=np.float64(h0), t3_gain=np.float64(t3_gain),
             obs=obs, eps=eps_full, idx=slot_full.astype(np.int64), num_transitions=info["num_transitions"],
             reward_mean=np.float64(info["reward_mean"]))
    for k, v in tr.items():
        d["out_" + k] = v.numpy()
    return d


def gen_buffer():
    _, _, _, _, ref_utils = _import_reference()
    rng = np.random.default_rng(3)
    S, A, cap = 5, 2, 50
    buf = ref_utils.ReplayBuffer(S, A, "cpu", max_size=cap)
    batches = []
    for M in (20, 25, 17, 50, 3):       # 20,45, wrap (12 spill), full-size batch, small
        b = dict(obss=torch.from_numpy(rng.standard_normal((M, S)).astype(np.float32)),
                 next_obss=torch.from_numpy(rng.standard_normal((M, S)).astype(np.float32)),
                 actions=torch.from_numpy(rng.standard_normal((M, A)).astype(np.float32)),
                 rewards=torch.from_numpy(rng.standard_normal((M, 1)).astype(np.float32)),
                 terminals=torch.from_numpy((rng.random((M, 1)) < 0.3).astype(np.float32)))
        batches.append(b)
    out = {"S": S, "A": A, "cap": cap, "n_batches": len(batches)}
    for i, b in enumerate(batches):
        buf.add_batch(b)
        for k, v in b.items():
            out[f"b{i}_{k}"] = v.numpy()
        out[f"after{i}_ptr"] = buf.ptr; out[f"after{i}_size"] = buf.size
        for f in ("state", "action", "next_state", "reward", "not_done"):
            out[f"after{i}_{f}"] = getattr(buf, f).numpy().copy()
    ind = rng.integers(0, buf.size, 33)
    with Inject(ind_list=[ind]):
        smp = buf.sample(33)
    out["ind"] = ind.astype(np.int64)
    for f, v in zip(("state", "action", "next_state", "reward", "not_done"), smp):
        out["sample_" + f] = v.numpy()
    return out


def gen_train(S, A, B, seed, n_steps=3):
    """Steady-state MOBODY.train steps (no refresh: total_it starts at 1)."""
    _, _, _, _, ref_utils = _import_reference()
    rng = np.random.default_rng(seed)
    pol, ag, cfg = build_reference_agent(S, A, seed)
    n_src, n_tar, n_fake = 4000, 600, 900

    def fill(buf, n):
        buf.state[:n] = torch.from_numpy(rng.standard_normal((n, S)).astype(np.float32))
        buf.action[:n] = torch.from_numpy(rng.uniform(-1, 1, (n, A)).astype(np.float32))
        buf.next_state[:n] = torch.from_numpy(rng.standard_normal((n, S)).astype(np.float32))
        buf.reward[:n] = torch.from_numpy(rng.standard_normal((n, 1)).astype(np.float32))
        buf.not_done[:n] = torch.from_numpy((rng.random((n, 1)) > 0.1).astype(np.float32))
        buf.size = n; buf.ptr = n
    src = ref_utils.ReplayBuffer(S, A, "cpu", max_size=n_src); fill(src, n_src)
    tar = ref_utils.ReplayBuffer(S, A, "cpu", max_size=n_tar); fill(tar, n_tar)
    pol.fake_replay_buffer = ref_utils.ReplayBuffer(S, A, "cpu", max_size=n_fake); fill(pol.fake_replay_buffer, n_fake)
    pol.total_it = 1
    inds = []
    for _ in range(n_steps):
        inds += [rng.integers(0, n_src, B), rng.integers(0, n_tar, B), rng.integers(0, n_fake, int(0.5 * B))]
    with Inject(ind_list=inds), contextlib.redirect_stdout(None):
        for _ in range(n_steps):
            pol.train(src, tar, B, None, None)
    out = dict(S=S, A=A, B=B, seed=seed, n_steps=n_steps, n_src=n_src, n_tar=n_tar, n_fake=n_fake)
    for nm, b in (("src", src), ("tar", tar), ("fake", pol.fake_replay_buffer)):
        for f in ("state", "action", "next_state", "reward", "not_done"):
            out[f"{nm}_{f}"] = getattr(b, f).numpy()[:b.size].copy()
    for i, ind in enumerate(inds):
        out[f"ind{i}"] = ind.astype(np.int64)
    # post-step parameters: strided subsample + sums (full tensors would be ~1 MB)
    for grp, sd in (("pi", pol.policy.state_dict()), ("q", pol.q_funcs.state_dict()),
                    ("qt", pol.target_q_funcs.state_dict())):
        for k, v in sd.items():
            flat = v.numpy().reshape(-1)
            out[f"post_{grp}_{k}_sub"] = flat[::37].copy()
            out[f"post_{grp}_{k}_sum"] = np.float64(flat.astype(np.float64).sum())
    # optimizer checkpoints (mobody.py:585, 587: torch.save(optimizer.state_dict())): Adam moments per parameter index
    import json
    for grp, opt in (("q", pol.q_optimizer), ("pi", pol.policy_optimizer)):
        sd = opt.state_dict()
        out[f"opt_{grp}_groups"] = json.dumps([{k: v for k, v in g.items()} for g in sd["param_groups"]], sort_keys=True)
        for i, stt in sd["state"].items():
            out[f"opt_{grp}_{i}_step"] = np.float64(float(stt["step"]))
            for mk in ("exp_avg", "exp_avg_sq"):
                flat = stt[mk].numpy().reshape(-1)
                out[f"opt_{grp}_{i}_{mk}_sub"] = flat[::37].copy()
                out[f"opt_{grp}_{i}_{mk}_sum"] = np.float64(flat.astype(np.float64).sum())
    return out


def gen_classifier(S, A, B, seed, n_steps=3):
    """update_classifier steps of the UNMODIFIED reference (mobody.py:146-181) with scripted buffer indices
    (np.random.randint), permutation (torch.randperm) and noise (torch.randn_like); then the reward relabel of
    mobody.py:364-378 computed with the reference Classifier.forward on the whole source buffer."""
    _, _, _, _, ref_utils = _import_reference()
    rng = np.random.default_rng(seed)
    pol, ag, cfg = build_reference_agent(S, A, seed, {"penalty_type": "dara", "penalty_coef": 0.7})
    cl = O.ClassifierState(S, A, seed)
    pol.classifier.load_state_dict(cl.params)
    n_src, n_tar = 1500, 400

    def fill(buf, n, shift):
        buf.state[:n] = torch.from_numpy(rng.standard_normal((n, S)).astype(np.float32))
        buf.action[:n] = torch.from_numpy(rng.uniform(-1, 1, (n, A)).astype(np.float32))
        buf.next_state[:n] = torch.from_numpy((rng.standard_normal((n, S)) + shift).astype(np.float32))
        buf.reward[:n] = torch.from_numpy(rng.standard_normal((n, 1)).astype(np.float32))
        buf.not_done[:n] = 1.0
        buf.size = n; buf.ptr = n
    src = ref_utils.ReplayBuffer(S, A, "cpu", max_size=n_src); fill(src, n_src, 0.0)
    tar = ref_utils.ReplayBuffer(S, A, "cpu", max_size=n_tar); fill(tar, n_tar, 0.5)
    inds, perms, noises = [], [], []
    for _ in range(n_steps):
        inds += [rng.integers(0, n_src, B), rng.integers(0, n_tar, B)]
        perms.append(rng.permutation(2 * B))
        noises += [rng.standard_normal((2 * B, 2 * S + A)).astype(np.float32), rng.standard_normal((2 * B, S + A)).astype(np.float32)]
    calls = {"perm": 0, "randn": 0}
    _rp, _rl = torch.randperm, torch.randn_like

    def randperm(n, **kw):
        r = torch.from_numpy(perms[calls["perm"]].astype(np.int64)); calls["perm"] += 1
        assert r.numel() == n
        return r

    def randn_like(x, **kw):
        r = torch.from_numpy(noises[calls["randn"]]); calls["randn"] += 1
        assert tuple(r.shape) == tuple(x.shape)
        return r
    losses = []
    torch.randperm, torch.randn_like = randperm, randn_like
    try:
        with Inject(ind_list=inds), contextlib.redirect_stdout(None), contextlib.redirect_stderr(None):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for _ in range(n_steps):
                    lsa, lsas = pol.update_classifier(src, tar, B)
                    losses.append([float(lsa), float(lsas)])
                # reward relabel, mobody.py:364-378, through the reference's own forward
                with torch.no_grad():
                    sas_logits, sa_logits = pol.classifier(src.state[:n_src], src.action[:n_src], src.next_state[:n_src], with_noise=False)
                    sas_probs, sa_probs = torch.softmax(sas_logits, -1), torch.softmax(sa_logits, -1)
                    sas_lp, sa_lp = torch.log(sas_probs + 1e-10), torch.log(sa_probs + 1e-10)
                    pen = (sas_lp[:, 1:] - sa_lp[:, 1:] - sas_lp[:, :1] + sa_lp[:, :1]).clamp(-10, 10)
                    new_reward = src.reward[:n_src] + cfg["penalty_coef"] * pen
    finally:
        torch.randperm, torch.randn_like = _rp, _rl
    out = dict(S=S, A=A, B=B, seed=seed, n_steps=n_steps, n_src=n_src, n_tar=n_tar, penalty_coef=np.float32(cfg["penalty_coef"]),
               std=np.float32(cfg["gaussian_noise_std"]), lr=np.float32(cfg["actor_lr"]), losses=np.asarray(losses, np.float64),
               reward_penalty=pen.numpy(), new_reward=new_reward.numpy())
    for nm, b in (("src", src), ("tar", tar)):
        for f in ("state", "action", "next_state", "reward", "not_done"):
            out[f"{nm}_{f}"] = getattr(b, f).numpy()[:b.size].copy()
    for i, ind in enumerate(inds):
        out[f"ind{i}"] = ind.astype(np.int64)
    for i, pm in enumerate(perms):
        out[f"perm{i}"] = pm.astype(np.int64)
    for i, nz in enumerate(noises):
        out[f"noise{i}"] = nz
    for k, v in pol.classifier.state_dict().items():
        flat = v.numpy().reshape(-1)
        out[f"post_{k}_sub"] = flat[::29].copy()
        out[f"post_{k}_sum"] = np.float64(flat.astype(np.float64).sum())
    return out


def recipe_dataset(S, A, n, seed, env):
    """Deterministic synthetic dataset rows (state, action, next_state, reward, not_done) from Philox recipes: a fixture
    only has to store the seed.  States sit around the env's healthy set so rollouts survive a few steps."""
    from oracle.philox import recipe_fill
    hs = healthy(env, S).astype(np.float32)
    return dict(state=hs[None] + recipe_fill((n, S), seed, 0.2), action=recipe_fill((n, A), seed + 1, 0.5).clip(-1, 1),
                next_state=hs[None] + recipe_fill((n, S), seed + 2, 0.2), reward=recipe_fill((n, 1), seed + 3, 1.0),
                not_done=np.ones((n, 1), np.float32))


def refresh_draws(S, seed, T_src, T_tar, n_batch):
    """Scripted draws of one MOBODY.train call at total_it == 1 with penalty_type='par' (mobody.py:428-475), as recipes:
    uniform noise with unit variance stands in for torch.normal's N(0,1) (the reference only multiplies it by std)."""
    from oracle.philox import recipe_fill, buffer_indices
    d = dict(par_eps=recipe_fill((7, n_batch, S), seed + 10, 1.0), par_idx=buffer_indices(seed, 10, n_batch, 5),
             src_eps=recipe_fill((T_src, 7, 50000, S), seed + 11, 1.0), src_idx=np.stack([buffer_indices(seed, 20 + t, 50000, 5) for t in range(T_src)]),
             tar_eps=recipe_fill((T_tar, 7, 2000, S), seed + 12, 1.0), tar_idx=np.stack([buffer_indices(seed, 40 + t, 2000, 5) for t in range(T_tar)]),
             sa_eps=recipe_fill((7, 50000, S), seed + 13, 1.0), sa_idx=buffer_indices(seed, 60, 50000, 5))
    return d


def gen_refresh(env="hopper", S=11, A=3, B=32, seed=51, n_src=3000, n_tar=500, T_src=2, T_tar=3, coef=1.0, h0=0.8, t3_gain=4.0,
                env_filter=0.3615, par_coef=0.1):
    """The synthetic-data refresh of MOBODY.train at total_it == 1 (mobody.py:441-475) run by the UNMODIFIED reference with
    penalty_type='par': fake-buffer contents (every 53rd row + float64 column sums + ptr/size) after the call."""
    _, _, _, _, ref_utils = _import_reference()
    from oracle.philox import buffer_indices
    dyn, _ = build_reference_dynamics(S, A, seed, env, coef, h0, t3_gain)
    pol, ag, cfg = build_reference_agent(S, A, seed, dict(penalty_type="par", penalty_coef=par_coef, env_filter=env_filter,
                                                            src_rollout_length=T_src, trg_rollout_length=T_tar))
    pol.dynamics = dyn
    bufs = {}
    for nm, n, sd in (("src", n_src, seed * 100), ("tar", n_tar, seed * 100 + 50)):
        b = ref_utils.ReplayBuffer(S, A, "cpu", max_size=n)
        for f, v in recipe_dataset(S, A, n, sd, env).items():
            getattr(b, f)[:n] = torch.from_numpy(v)
        b.size = n; b.ptr = 0
        bufs[nm] = b
    dr = refresh_draws(S, seed, T_src, T_tar, B)
    inds = [buffer_indices(seed, 1, B, n_src), buffer_indices(seed, 2, B, n_tar), buffer_indices(seed, 3, 50000, n_src),
            buffer_indices(seed, 4, 2000, n_tar)]
    eps_calls = [dr["par_eps"]] + [dr["src_eps"][t] for t in range(T_src)] + [dr["tar_eps"][t] for t in range(T_tar)] + [dr["sa_eps"]]
    idx_calls = [dr["par_idx"]] + [dr["src_idx"][t] for t in range(T_src)] + [dr["tar_idx"][t] for t in range(T_tar)] + [dr["sa_idx"]]

    class Stop(Exception):
        pass

    def randint_then_stop(c):          # the 5th randint call is fake_replay_buffer.sample: the refresh is over
        raise Stop
    with Inject(lambda c, shp: torch.from_numpy(eps_calls[c][:, :shp[1]].copy()), lambda c, n: idx_calls[c][:n], ind_list=inds) as inj, \
            contextlib.redirect_stdout(None):
        try:
            pol.train(bufs["src"], bufs["tar"], B, None, None)
        except IndexError:             # ind_list exhausted at the fake-buffer sample: everything before it has run
            pass
        assert inj.calls["normal"] == len(eps_calls) and inj.calls["choice"] == len(idx_calls), inj.calls
    fb = pol.fake_replay_buffer
    n = fb.size
    rows = torch.cat([fb.state[:n], fb.action[:n], fb.next_state[:n], fb.reward[:n], fb.not_done[:n]], 1).numpy()
    out = dict(env=env, S=S, A=A, B=B, seed=seed, n_src=n_src, n_tar=n_tar, T_src=T_src, T_tar=T_tar, coef=np.float32(coef),
               h0=np.float64(h0), t3_gain=np.float64(t3_gain), env_filter=np.float32(env_filter), par_coef=np.float32(par_coef),
               fake_size=n, fake_ptr=fb.ptr, col_sums=rows.astype(np.float64).sum(0), not_done_count=int(rows[:, -1].sum()))
    # The same three dynamics calls once more WITHOUT the penalty filter, under the same scripted draws: every transition the
    # block produced, its terminal flag and whether the block kept it.  A checker can then line rows up by (segment, step,
    # start row) even when a reduced-precision run flips a row that sits on a threshold.  Applying the masks must
    # reproduce the fake buffer the reference's train() call built above -- asserted here, so the fixture IS that call.
    src_init = torch.from_numpy(recipe_dataset(S, A, n_src, seed * 100, env)["state"])[inds[2]]
    src_act = torch.from_numpy(recipe_dataset(S, A, n_src, seed * 100, env)["action"])[inds[2]]
    tar_init = torch.from_numpy(recipe_dataset(S, A, n_tar, seed * 100 + 50, env)["state"])[inds[3]]
    pol.config["filter_bad_rollout"] = 0
    rebuilt = []
    for seg, init, T, eps_all, idx_all in (("src", src_init, T_src, dr["src_eps"], dr["src_idx"]), ("tar", tar_init, T_tar, dr["tar_eps"], dr["tar_idx"])):
        with Inject(lambda c, shp: torch.from_numpy(eps_all[c][:, :shp[1]].copy()), lambda c, n_: idx_all[c][:n_]), contextlib.redirect_stdout(None):
            tr, _ = pol.rollout(init, T, True)
        allrows = torch.cat([tr["obss"], tr["actions"], tr["next_obss"], tr["rewards"], 1.0 - tr["terminals"]], 1).numpy()
        keep = (tr["penalty"] <= env_filter).squeeze(1).numpy()                      # mobody.py:648-651
        out.update({seg + "_n_all": len(allrows), seg + "_term": np.packbits(tr["terminals"].numpy().astype(bool).ravel()),
                    seg + "_keep": np.packbits(keep), seg + "_sub_rows": allrows[::41].copy()})
        rebuilt.append(allrows[keep])
    with Inject(lambda c, shp: torch.from_numpy(dr["sa_eps"]), lambda c, n_: dr["sa_idx"]):
        nobs, rew, term, info = dyn.step(src_init, src_act)
    allrows = torch.cat([src_init, src_act, nobs, rew, 1.0 - torch.from_numpy(term.astype(np.float32))], 1).numpy()
    keep = (info["penalty"] < env_filter).squeeze(1).numpy()                         # strict, mobody.py:468
    out.update(sa_n_all=len(allrows), sa_term=np.packbits(term.ravel()), sa_keep=np.packbits(keep), sa_sub_rows=allrows[::41].copy())
    rebuilt.append(allrows[keep])
    assert np.array_equal(np.concatenate(rebuilt, 0), rows), "mask reconstruction differs from the reference's refresh block"
    # the `par` reward penalty of the same call (mobody.py:428-434): the reference keeps it in a local, so it is re-derived
    # here from the reference's own dynamics.step under the draws that call consumed
    ds = recipe_dataset(S, A, n_src, seed * 100, env)
    bs, ba, bns = (torch.from_numpy(ds[k])[inds[0]] for k in ("state", "action", "next_state"))
    with Inject(lambda c, shp: torch.from_numpy(dr["par_eps"]), lambda c, n_: dr["par_idx"]):
        pred, _, _, _ = dyn.step(bs, ba)
    out["par_penalty"] = torch.mean((bns - pred) ** 2, axis=1, keepdims=True).numpy()
    return out


FIT_SAMPLE = 1536      # elements kept per tensor in the dynamics-fitting fixtures (evenly strided)


def fit_sample_index(n):
    return np.unique(np.linspace(0, n - 1, min(n, FIT_SAMPLE)).astype(np.int64))


def gen_dynfit(S, A, B, seed, calls=(True, True, False)):
    """MOBODYEnsembleDynamics.learn (mobody_dynamics.py:594-653) of the UNMODIFIED reference, one mini-batch per call, on
    one model + torch.optim.Adam (train_mobody.py:801-804) across the calls; `calls` = use_trg_data per call.  The
    torch.randn_like draws (six reparameterize calls + reward_loss's randn_like(mean), in that order) are scripted and
    recorded; `.to('cuda')` is a no-op (CPU run).  Kept: the returned loss scalars per call, and evenly strided samples of
    every trained tensor's value / exp_avg / exp_avg_sq after the last call (+ exp_avg after the first: 0.1 * gradient)."""
    dyn, p0 = build_reference_dynamics(S, A, seed, "hopper" if S == 11 else "walker2d", 1.0)
    dyn.config.update(no_vae=0, inverse_sep_reward_loss=0, latent_reward=0, train_together=0)
    dyn.encoder_loss_coef = 1
    dyn.optim = torch.optim.Adam(dyn.model.parameters(), lr=1e-3)
    dyn.total_steps = 0
    rng = np.random.default_rng(seed)
    out = dict(S=S, A=A, B=B, seed=seed, calls=np.asarray(calls, dtype=np.int64))
    real_to, real_randn_like = torch.Tensor.to, torch.randn_like
    names = [n for n in O.dynamics_layer_shapes(S, A)]
    for c, use_trg in enumerate(calls):
        s = (HEALTHY["hopper" if S == 11 else "walker2d"](S)[None, None, :] + 0.3 * rng.standard_normal((7, B, S))).astype(np.float32)
        a = rng.uniform(-1, 1, (7, B, A)).astype(np.float32)
        ns = (s + 0.1 * rng.standard_normal((7, B, S))).astype(np.float32)
        r = rng.standard_normal((7, B, 1)).astype(np.float32)
        eps_l = rng.standard_normal((6, 7, B, 16)).astype(np.float32)
        eps_n = rng.standard_normal((7, B, S)).astype(np.float32)
        draws = [torch.from_numpy(eps_l[k]) for k in range(6)] + [torch.from_numpy(eps_n)]
        count = [0]

        def randn_like(x, **kw):
            d = draws[count[0]]; count[0] += 1
            assert tuple(d.shape) == tuple(x.shape), (count[0], d.shape, x.shape)
            return d.clone()

        def to(self, *args, **kw):
            if args and isinstance(args[0], str) and args[0].startswith("cuda"):
                return self
            return real_to(self, *args, **kw)
        torch.randn_like, torch.Tensor.to = randn_like, to
        try:
            res = dyn.learn(bool(use_trg), *(torch.from_numpy(x) for x in (s, a, ns, r)), B, 0.01)
        finally:
            torch.randn_like, torch.Tensor.to = real_randn_like, real_to
        assert count[0] == 7, count
        out.update({f"c{c}_obs": s, f"c{c}_act": a, f"c{c}_next_obs": ns, f"c{c}_reward": r, f"c{c}_eps_latent": eps_l,
                    f"c{c}_eps_next": eps_n, f"c{c}_losses": np.asarray(res, dtype=np.float64)})   # loss, transition, encoder, recon, kl
        if c == 0:
            for n in names:
                for sfx in ("weight", "bias"):
                    st = dyn.optim.state.get(getattr(getattr(dyn.model, n), sfx))
                    if st:
                        out[f"first_m_{n}.{sfx}"] = st["exp_avg"].numpy().ravel()[fit_sample_index(st["exp_avg"].numel())].copy()
    for n in names:
        for sfx in ("weight", "bias"):
            t = getattr(getattr(dyn.model, n), sfx)
            idx = fit_sample_index(t.numel())
            out[f"p_{n}.{sfx}"] = t.detach().numpy().ravel()[idx].copy()
            st = dyn.optim.state.get(t)
            out[f"has_state_{n}.{sfx}"] = np.int64(bool(st))
            if st:
                out[f"m_{n}.{sfx}"] = st["exp_avg"].numpy().ravel()[idx].copy()
                out[f"v_{n}.{sfx}"] = st["exp_avg_sq"].numpy().ravel()[idx].copy()
                out[f"t_{n}.{sfx}"] = np.int64(int(st["step"]))
    return out


def gen_checkpoint_keys():
    """Key -> shape of every checkpoint file the reference writes for this path: dynamics.pth
    (MOBODYModule.state_dict, mobody_dynamics.py:1158-1160), <name>_actor / <name>_critic (mobody.py:584-588) and the
    classifier (for completeness)."""
    MOBODYModule, _, _, _, _ = _import_reference()
    cfg = {"mopo": 0, "latent_reward": 0, "encoder_loss_coef": 1, "domain_loss_coef": 0, "cycle_loss_coef": 0}
    out = {}
    for S, A in ((17, 6), (11, 3)):
        with contextlib.redirect_stdout(None):
            m = MOBODYModule(S, A, hidden_dims=256, num_ensemble=7, num_elites=5, device="cpu", config=cfg)
        pol, _, _ = build_reference_agent(S, A, 1)
        out[f"S{S}A{A}"] = {
            "dynamics.pth": {k: list(v.shape) for k, v in m.state_dict().items()},
            "_actor": {k: list(v.shape) for k, v in pol.policy.state_dict().items()},
            "_critic": {k: list(v.shape) for k, v in pol.q_funcs.state_dict().items()},
            "classifier": {k: list(v.shape) for k, v in pol.classifier.state_dict().items()},
        }
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0); np.random.seed(0)
    torch.set_num_threads(1)
    steps = [gen_step("walker2d", 17, 6, 64, 11, 5.0), gen_step("halfcheetah", 17, 6, 48, 12, 0.1),
             gen_step("hopper", 11, 3, 40, 13, 5.0), gen_step("ant", 27, 8, 36, 14, 1.0),
             gen_step("ant", 29, 8, 33, 15, 0.0), gen_step("walker2d", 17, 6, 32, 16, 5.0, use_trg=False)]
    for i, d in enumerate(steps):
        np.savez_compressed(os.path.join(OUT, f"step_{i}_{d['env']}_S{d['S']}A{d['A']}.npz"), **d)
    np.savez_compressed(os.path.join(OUT, "termination.npz"), **gen_termination())
    np.savez_compressed(os.path.join(OUT, "rollout_walker2d_T3.npz"), **gen_rollout("walker2d", 17, 6, 96, 3, 21, 5.0, 0.7803, 1.0, 6.0))
    np.savez_compressed(os.path.join(OUT, "rollout_hopper_T5.npz"), **gen_rollout("hopper", 11, 3, 80, 5, 22, 1.0, 0.3233, 1.1, 3.0))
    np.savez_compressed(os.path.join(OUT, "buffer.npz"), **gen_buffer())
    np.savez_compressed(os.path.join(OUT, "train_S17A6_B32.npz"), **gen_train(17, 6, 32, 31))
    np.savez_compressed(os.path.join(OUT, "train_S11A3_B16.npz"), **gen_train(11, 3, 16, 32, n_steps=2))
    np.savez_compressed(os.path.join(OUT, "classifier_S17A6_B32.npz"), **gen_classifier(17, 6, 32, 41))
    np.savez_compressed(os.path.join(OUT, "refresh_hopper.npz"), **gen_refresh())
    np.savez_compressed(os.path.join(OUT, "dynfit_S11A3_B24.npz"), **gen_dynfit(11, 3, 24, 51))
    np.savez_compressed(os.path.join(OUT, "dynfit_S17A6_B40.npz"), **gen_dynfit(17, 6, 40, 52, calls=(True,)))
    import json
    with open(os.path.join(OUT, "checkpoint_keys.json"), "w") as f:
        json.dump(gen_checkpoint_keys(), f, indent=0, sort_keys=True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
