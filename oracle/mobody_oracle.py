"""Torch-CPU restatement of the MOBODY hot path (rollout + Q-weighted BC step).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it follows (paths relative to /root/reference).  All noise and
all random indices are *explicit arguments*, so the same inputs can be fed to
the reference (via monkey-patched RNG, oracle/make_golden.py) and to the CUDA
path.  Floating point is fp32 with the reference's own op order, so on the same
CPU/torch build the oracle reproduces the reference bit for bit.
"""
import copy
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .philox import recipe_fill

E_MEMBERS = 7          # num_ensemble, train_mobody.py:795; also literal 7 in mobody_dynamics.py:218
N_ELITES = 5           # train_mobody.py:796
HIDDEN = 256           # train_mobody.py:794
LATENT = 16            # mobody_module.py:95
ZA_HIDDEN = 32         # mobody_module.py:110-111

# name -> (in, out) builder; mobody_module.py:97-184
def dynamics_layer_shapes(S, A, H=HIDDEN):
    return OrderedDict([
        ("zs1", (S, H)), ("zs2", (H, H)), ("zs3", (H, 2 * LATENT)),
        ("za_src1", (LATENT + A, ZA_HIDDEN)), ("za_src2", (ZA_HIDDEN, 2 * LATENT)),
        ("za_trg1", (LATENT + A, ZA_HIDDEN)), ("za_trg2", (ZA_HIDDEN, 2 * LATENT)),
        ("transition1", (LATENT, H)), ("transition2", (H, H)), ("transition3", (H, S)),
        ("reward_model1", (2 * S + A, H)), ("reward_model2", (H, H)), ("reward_model3", (H, 2)),
    ])


def make_dynamics_params(S, A, seed, H=HIDDEN, E=E_MEMBERS, healthy_state=None, t3_gain=1.0):
    """Recipe weights: W[E,in,out] with std 1/(2 sqrt(in)) (mobody_module.py:386), biases
    randomised (std 0.1) so they matter.  ``healthy_state`` (len S) is added to
    transition3.bias so synthetic multi-step rollouts survive."""
    p = OrderedDict()
    for li, (name, (din, dout)) in enumerate(dynamics_layer_shapes(S, A, H).items()):
        w = recipe_fill((E, din, dout), seed * 1000 + 2 * li, 1.0 / (2.0 * din ** 0.5))
        b = recipe_fill((E, 1, dout), seed * 1000 + 2 * li + 1, 0.1)
        p[name + ".weight"] = torch.from_numpy(w)
        p[name + ".bias"] = torch.from_numpy(b)
    if t3_gain != 1.0:   # more next-state spread / member disagreement for rollout fixtures
        p["transition3.weight"] = p["transition3.weight"] * float(t3_gain)
    if healthy_state is not None:
        p["transition3.bias"] = p["transition3.bias"] * 0.1 + torch.as_tensor(
            healthy_state, dtype=torch.float32).reshape(1, 1, S)
    p["elites"] = torch.arange(N_ELITES)
    return p


def make_mlp_params(din, dout, seed, H=HIDDEN):
    """nn.Linear-layout MLP (weight [out,in]); mobody.py:35-48."""
    p = OrderedDict()
    dims = [(din, H), (H, H), (H, dout)]
    for li, (i, o) in enumerate(dims):
        p[f"network.{2 * li}.weight"] = torch.from_numpy(recipe_fill((o, i), seed * 1000 + 2 * li, 1.0 / i ** 0.5 * 0.6))
        p[f"network.{2 * li}.bias"] = torch.from_numpy(recipe_fill((o,), seed * 1000 + 2 * li + 1, 0.05))
    return p


# --------------------------------------------------------------------------
# ensemble dynamics forward
# --------------------------------------------------------------------------
def swish(x):
    """mobody_module.py:13-15."""
    return x * torch.sigmoid(x)


def ensemble_linear(x, w, b):
    """mobody_module.py:393-404: 2-D input broadcasts over members."""
    if x.dim() == 2:
        y = torch.einsum("ij,bjk->bik", x, w)
    else:
        y = torch.einsum("bij,bjk->bik", x, w)
    return y + b


def _lin(p, name, x):
    return ensemble_linear(x, p[name + ".weight"], p[name + ".bias"])


def encode_state(p, s):
    """mobody_module.py:217-225 with training=False (reparameterize returns mu, :237-243)."""
    h = swish(_lin(p, "zs1", s))
    h = swish(_lin(p, "zs2", h))
    z = _lin(p, "zs3", h)
    mu, logvar = torch.chunk(z, 2, dim=-1)
    return mu, logvar


def encode_action(p, zs, a, use_trg=True):
    """mobody_module.py:258-271 (trg) / :245-256 (src)."""
    E = zs.shape[0]
    if zs.dim() == 3 and a.dim() == 2:
        a = a.unsqueeze(0).repeat(E, 1, 1)
    sa = torch.cat([zs, a], dim=-1)
    pre = "za_trg" if use_trg else "za_src"
    g = swish(_lin(p, pre + "1", sa))
    za = _lin(p, pre + "2", g)
    mu, _ = torch.chunk(za, 2, dim=-1)
    return mu


def encode_transition(p, z):
    """mobody_module.py:287-293."""
    u = swish(_lin(p, "transition1", z))
    u = swish(_lin(p, "transition2", u))
    return _lin(p, "transition3", u)


def encode_reward(p, s, a, ns):
    """mobody_module.py:295-302; logvar soft-clamp (:18-29) kept for completeness."""
    sas = torch.cat([s, a, ns], dim=-1)
    v = swish(_lin(p, "reward_model1", sas))
    v = swish(_lin(p, "reward_model2", v))
    out = _lin(p, "reward_model3", v)
    mu, logvar = torch.chunk(out, 2, dim=-1)
    logvar = 0.5 - F.softplus(0.5 - logvar)
    logvar = -10 + F.softplus(logvar - (-10))
    return mu, logvar


def forward_dynamics(p, s, a, use_trg=True):
    """mobody_module.py:315-330 -> (mean[E,B,S], zs_mu, zs_logvar)."""
    zs, zlv = encode_state(p, s)
    za = encode_action(p, zs, a, use_trg)
    return encode_transition(p, zs + za), zs, zlv


# --------------------------------------------------------------------------
# termination predicates (terminal_funs.py); next_obs is an fp32 numpy array
# --------------------------------------------------------------------------
def term_halfcheetah(x):   # :10-16
    return ~(np.all(x > -100, axis=-1) & np.all(x < 100, axis=-1))


def term_hopper(x):        # :18-30  (no lower bound on dims>=1: abs() is applied to the bool)
    ok = np.isfinite(x).all(-1) & (x[:, 1:] < 100).all(-1) & (x[:, 0] > .7) & (np.abs(x[:, 1]) < .2)
    return ~ok


def term_walker2d(x):      # :63-75
    ok = (np.all(x > -100, -1) & np.all(x < 100, -1) & (x[:, 0] > 0.8) & (x[:, 0] < 2.0)
          & (x[:, 1] > -1.0) & (x[:, 1] < 1.0))
    return ~ok


def term_ant(x):           # :51-61 (antangle :39-49 is identical)
    ok = np.isfinite(x).all(-1) & (x[:, 0] >= 0.2) & (x[:, 0] <= 1.0)
    return ~ok


def term_never(x):         # :32-37, 77-96, 115-121
    return np.zeros(len(x), dtype=bool)


def term_humanoid(x):      # :98-104
    return (x[:, 0] < 1.0) | (x[:, 0] > 2.0)


def term_pen(x):           # :106-113
    return x[:, 26] < 0.075


TERM_KINDS = OrderedDict([  # numeric ids shared with include/mobody_b200.h
    ("never", 0), ("halfcheetah", 1), ("hopper", 2), ("walker2d", 3), ("ant", 4),
    ("humanoid", 5), ("pen", 6),
])
_TERM_FN = {0: term_never, 1: term_halfcheetah, 2: term_hopper, 3: term_walker2d,
            4: term_ant, 5: term_humanoid, 6: term_pen}


def termination_kind(task):
    """Dispatcher order of terminal_funs.py:123-149 (substring match, order matters)."""
    table = [("halfcheetahvel", "never"), ("halfcheetah", "halfcheetah"), ("hopper", "hopper"),
             ("antangle", "ant"), ("ant", "ant"), ("walker2d", "walker2d"),
             ("point2denv", "never"), ("point2dwallenv", "never"), ("pendulum", "never"),
             ("humanoid", "humanoid"), ("pen", "pen"), ("door", "never")]
    for key, kind in table:
        if key in task:
            return TERM_KINDS[kind]
    raise TypeError(f"no termination function for task {task!r}")  # reference: `raise np.zeros` -> TypeError


def termination(kind, next_obs):
    """-> bool[B,1] like the reference's done[:,None]."""
    x = np.asarray(next_obs, dtype=np.float32)
    return _TERM_FN[int(kind)](x)[:, None]


# --------------------------------------------------------------------------
# one imagined step; mobody_dynamics.py:193-265
# --------------------------------------------------------------------------
@torch.no_grad()
def step(p, obs, act, eps, idx, term_kind, penalty_coef, use_penalty=True, use_trg=True):
    """eps: [E,B,S] standard normals (the reference's torch.normal(0,std) == eps*std, :220);
    idx: int64[B] member index per row (np.random.choice(elites, B), :225).
    Returns dict with next_obs, reward, terminal(bool[B,1] numpy), mean, std, raw_reward, penalty."""
    mean, _, _ = forward_dynamics(p, obs, act, use_trg)                       # :212-214
    E = mean.shape[0]
    std = torch.std(mean, dim=0, keepdim=True).repeat(E, 1, 1)                # :218
    samples_all = mean + eps * std                                            # :220
    B = obs.shape[0]
    idx_t = torch.as_tensor(np.asarray(idx), dtype=torch.long)
    next_obs = samples_all[idx_t, torch.arange(B)]                            # :226
    r_mu, _ = encode_reward(p, obs, act, next_obs)                            # :235
    raw_reward = r_mu.mean(0)                                                 # :236 (all 7 members)
    terminal = termination(term_kind, next_obs.numpy())                       # :237
    m = mean[..., :-1]                                                        # :246 (last dim dropped)
    diff = m - torch.mean(m, dim=0)
    penalty = torch.amax(torch.norm(diff, dim=2), dim=0).reshape(B, 1)        # :249-255
    reward = raw_reward
    if penalty_coef and use_penalty:                                          # :261-263
        reward = raw_reward - penalty_coef * penalty
    return dict(next_obs=next_obs, reward=reward, terminal=terminal, mean=mean, std=std[0],
                raw_reward=raw_reward, penalty=penalty)


# --------------------------------------------------------------------------
# MLPs of the agent; mobody.py:35-83
# --------------------------------------------------------------------------
def mlp_forward(p, x, prefix=""):
    h = F.relu(F.linear(x, p[prefix + "network.0.weight"], p[prefix + "network.0.bias"]))
    h = F.relu(F.linear(h, p[prefix + "network.2.weight"], p[prefix + "network.2.bias"]))
    return F.linear(h, p[prefix + "network.4.weight"], p[prefix + "network.4.bias"])


def policy_forward(p, s, max_action):
    """mobody.py:60-72."""
    return torch.tanh(mlp_forward(p, s, "network.")) * max_action


def double_q_forward(p, s, a):
    """mobody.py:74-83."""
    x = torch.cat((s, a), dim=1)
    return mlp_forward(p, x, "network1."), mlp_forward(p, x, "network2.")


# --------------------------------------------------------------------------
# rollout; mobody.py:596-657
# --------------------------------------------------------------------------
@torch.no_grad()
def rollout(dyn_p, pol_p, max_action, init_obs, T, eps_list, idx_list, term_kind, penalty_coef,
            env_filter=10.0, filter_bad=True, use_trg=True):
    """eps_list[t]: [E,B_t,S]; idx_list[t]: [B_t] where B_t is the number of rows alive at step t
    (callables taking B_t are accepted so callers need not know B_t in advance).
    Returns (dict of tensors | None, info | None)."""
    if T == 0:
        return None, None                                                     # :602-603
    out = {k: [] for k in ("obss", "next_obss", "actions", "rewards", "terminals", "penalty")}
    obs = init_obs
    n_trans = 0
    rew_all = []
    A = pol_p["network.network.4.bias"].shape[0]
    for t in range(T):
        act = policy_forward(pol_p, obs, max_action).reshape(-1, A)          # :612, 138-144
        eps = eps_list[t](len(obs)) if callable(eps_list[t]) else eps_list[t]
        idx = idx_list[t](len(obs)) if callable(idx_list[t]) else idx_list[t]
        r = step(dyn_p, obs, act, eps, idx, term_kind, penalty_coef, True, use_trg)
        out["obss"].append(obs); out["next_obss"].append(r["next_obs"]); out["actions"].append(act)
        out["rewards"].append(r["reward"]); out["penalty"].append(r["penalty"])
        out["terminals"].append(torch.from_numpy(r["terminal"].astype(np.float32)))   # :628
        n_trans += len(obs)
        rew_all.append(r["reward"].numpy().flatten())
        alive = ~r["terminal"].flatten()                                      # :635
        if alive.sum() == 0:
            break
        obs = r["next_obs"][torch.from_numpy(alive)]                          # :639
    cat = {k: torch.cat(v, 0) for k, v in out.items()}
    if filter_bad:
        keep = (cat["penalty"] <= env_filter).squeeze(1)                      # :648-651 (<=)
        cat = {k: v[keep] for k, v in cat.items()}
    return cat, {"num_transitions": n_trans, "reward_mean": float(np.concatenate(rew_all).mean())}


# --------------------------------------------------------------------------
# ReplayBuffer; utils.py:13-193
# --------------------------------------------------------------------------
class RingBuffer:
    def __init__(self, S, A, max_size):
        self.max_size, self.ptr, self.size = int(max_size), 0, 0
        self.state = torch.zeros(max_size, S); self.action = torch.zeros(max_size, A)
        self.next_state = torch.zeros(max_size, S); self.reward = torch.zeros(max_size, 1)
        self.not_done = torch.zeros(max_size, 1)

    def add_batch(self, batch):
        """utils.py:43-92: one wrap only; stores 1 - terminals."""
        if batch is None:
            return
        s, ns, a, r, d = batch["obss"], batch["next_obss"], batch["actions"], batch["rewards"], batch["terminals"]
        M = len(s)
        end = min(self.ptr + M, self.max_size)
        used = end - self.ptr
        self.state[self.ptr:end] = s[:used]; self.action[self.ptr:end] = a[:used]
        self.next_state[self.ptr:end] = ns[:used]; self.reward[self.ptr:end] = r[:used]
        self.not_done[self.ptr:end] = 1. - d[:used]
        self.ptr = end % self.max_size
        self.size = min(self.size + used, self.max_size)
        if self.ptr == 0:
            rest = M - used
            self.state[0:rest] = s[used:]; self.action[0:rest] = a[used:]; self.next_state[0:rest] = ns[used:]
            self.reward[0:rest] = r[used:]; self.not_done[0:rest] = 1. - d[used:]
            self.ptr = rest

    def gather(self, ind):
        """utils.py:142-148 with the indices given (np.random.randint(0,size,B), :128)."""
        ind = torch.as_tensor(np.asarray(ind), dtype=torch.long)
        return (self.state[ind], self.action[ind], self.next_state[ind], self.reward[ind], self.not_done[ind])


# --------------------------------------------------------------------------
# steady-state train step; mobody.py:189-208, 246-276, 314-345, 516-578
# --------------------------------------------------------------------------
class AgentState:
    """Parameters + Adam state of policy / twin-Q / target-Q as plain tensors."""
    def __init__(self, S, A, seed, H=HIDDEN):
        self.S, self.A = S, A
        self.policy = OrderedDict(("network." + k, v) for k, v in make_mlp_params(S, A, seed + 1, H).items())
        q = OrderedDict()
        for n, sd in (("network1.", seed + 2), ("network2.", seed + 3)):
            for k, v in make_mlp_params(S + A, 1, sd, H).items():
                q[n + k] = v
        self.q = q
        self.q_target = copy.deepcopy(q)
        self.reset_optim()

    def reset_optim(self):
        self.t_q = 0; self.t_pi = 0
        self.m_q = {k: torch.zeros_like(v) for k, v in self.q.items()}
        self.v_q = {k: torch.zeros_like(v) for k, v in self.q.items()}
        self.m_pi = {k: torch.zeros_like(v) for k, v in self.policy.items()}
        self.v_pi = {k: torch.zeros_like(v) for k, v in self.policy.items()}


def adam_update(params, grads, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (mobody.py:127-131), single-tensor formulation."""
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    for k in params:
        g = grads[k]
        m[k].mul_(b1).add_(g, alpha=1 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v[k].sqrt() / (bc2 ** 0.5)).add_(eps)
        params[k].addcdiv_(m[k], denom, value=-(lr / bc1))


def train_step(ag, batch, n_true, cfg):
    """batch = (s, a, s', r, nd) with rows ordered src,tar,fake (N rows); the first n_true rows
    are the "true" (src+tar) rows used by the BC term.  cfg: gamma,tau,actor_lr,critic_lr,weight,
    bc_coef,max_action.  Mutates ``ag`` in place; returns dict of losses/diagnostics."""
    s, a, s2, r, nd = batch
    ma = cfg["max_action"]
    # ---- critic (mobody.py:189-208, 546-548) ----
    with torch.no_grad():
        a2 = policy_forward(ag.policy, s2, ma)
        t1, t2 = double_q_forward(ag.q_target, s2, a2)
        y = r + nd * cfg["gamma"] * torch.min(t1, t2)
    qp = {k: v.clone().requires_grad_(True) for k, v in ag.q.items()}
    q1, q2 = double_q_forward(qp, s, a)
    q_loss = F.mse_loss(q1, y) + F.mse_loss(q2, y)
    grads = dict(zip(qp.keys(), torch.autograd.grad(q_loss, list(qp.values()))))
    ag.t_q += 1
    with torch.no_grad():
        adam_update(ag.q, grads, ag.m_q, ag.v_q, ag.t_q, cfg["critic_lr"])
        for k in ag.q:                                                        # :183-187 every step
            ag.q_target[k].copy_(cfg["tau"] * ag.q[k] + (1.0 - cfg["tau"]) * ag.q_target[k])
    # ---- actor (mobody.py:314-345, 246-276) with Q frozen ----
    pp = {k: v.clone().requires_grad_(True) for k, v in ag.policy.items()}
    pa = policy_forward(pp, s, ma)
    b1, b2 = double_q_forward(ag.q, s, pa)
    qv = torch.min(b1, b2)
    p_w = cfg["weight"] / qv.abs().mean().detach()
    pi_loss = p_w * (-qv).mean()
    st, at = s[:n_true], a[:n_true]
    pred = policy_forward(pp, st, ma)
    with torch.no_grad():
        h1, h2 = double_q_forward(ag.q, st, at)
        adv = torch.min(h1, h2)
        adv = adv / adv.abs().mean()
    w = torch.exp(3 * adv).clamp(max=100.0)
    bc = torch.mean(w * (pred - at) ** 2)
    loss = pi_loss + cfg["bc_coef"] * bc
    grads = dict(zip(pp.keys(), torch.autograd.grad(loss, list(pp.values()))))
    ag.t_pi += 1
    with torch.no_grad():
        adam_update(ag.policy, grads, ag.m_pi, ag.v_pi, ag.t_pi, cfg["actor_lr"])
    return dict(q_loss=float(q_loss.detach()), pi_loss=float(loss.detach()), bc_loss=float(bc.detach()), q1_mean=float(q1.detach().mean()),
                q_policy=float(qv.detach().mean()), w_mean=float(w.mean()), w_min=float(w.min()), w_max=float(w.max()))


# --------------------------------------------------------------------------
# DARA domain classifier; mobody.py:11-33 (Classifier), :146-181 (update_classifier), :354-381 (reward relabel)
# --------------------------------------------------------------------------
class ClassifierState:
    """Parameters + Adam state of Classifier.sa_classifier / .sas_classifier (MLPNetwork in -> 256 -> 256 -> 2)."""
    def __init__(self, S, A, seed, H=HIDDEN):
        self.S, self.A = S, A
        p = OrderedDict()
        for k, v in make_mlp_params(S + A, 2, seed + 4, H).items():
            p["sa_classifier." + k] = v
        for k, v in make_mlp_params(2 * S + A, 2, seed + 5, H).items():
            p["sas_classifier." + k] = v
        self.params = p
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in p.items()}


def classifier_forward(p, s, a, s2, noise_sas=None, noise_sa=None, std=1.0):
    """Classifier.forward (mobody.py:20-33): returns SOFTMAXED outputs (the reference calls them logits).
    noise_* are the N(0,1) draws of torch.randn_like (with_noise=True) or None (with_noise=False)."""
    sas = torch.cat([s, a, s2], -1)
    if noise_sas is not None:
        sas = sas + noise_sas * std
    sas_logits = torch.softmax(mlp_forward(p, sas, "sas_classifier."), dim=1)
    sa = torch.cat([s, a], -1)
    if noise_sa is not None:
        sa = sa + noise_sa * std
    sa_logits = torch.softmax(mlp_forward(p, sa, "sa_classifier."), dim=1)
    return sas_logits, sa_logits


def classifier_update(cl, s, a, s2, label, noise_sas, noise_sa, std, lr):
    """One update_classifier step (mobody.py:146-181) on an already sampled, concatenated and permuted batch:
    cross_entropy is applied to the softmaxed outputs (double softmax, kept as in the reference); Adam(lr=actor_lr)."""
    pp = {k: v.clone().requires_grad_(True) for k, v in cl.params.items()}
    sas_logits, sa_logits = classifier_forward(pp, s, a, s2, noise_sas, noise_sa, std)
    loss_sas = F.cross_entropy(sas_logits, label)
    loss_sa = F.cross_entropy(sa_logits, label)
    grads = dict(zip(pp.keys(), torch.autograd.grad(loss_sas + loss_sa, list(pp.values()))))
    cl.t += 1
    with torch.no_grad():
        adam_update(cl.params, grads, cl.m, cl.v, cl.t, lr)
    return float(loss_sa.detach()), float(loss_sas.detach())


@torch.no_grad()
def dara_reward_penalty(p, s, a, s2):
    """mobody.py:371-377: log-ratio of the twice-softmaxed classifier outputs, clamped to [-10, 10]."""
    sas_logits, sa_logits = classifier_forward(p, s, a, s2)
    sas_lp = torch.log(torch.softmax(sas_logits, -1) + 1e-10)
    sa_lp = torch.log(torch.softmax(sa_logits, -1) + 1e-10)
    pen = sas_lp[:, 1:] - sa_lp[:, 1:] - sas_lp[:, :1] + sa_lp[:, :1]
    return pen.clamp(-10, 10)


# --------------------------------------------------------------------------
# dynamics fitting step
# --------------------------------------------------------------------------
FIT_LAYERS_SHARED = ("zs1", "zs2", "zs3", "transition1", "transition2", "transition3",
                     "reward_model1", "reward_model2", "reward_model3")


def fit_trained_layers(use_trg):
    """Layers that receive a gradient in MOBODYEnsembleDynamics.learn (mobody_dynamics.py:594-653): the other domain's
    action encoder and the action decoders are never touched, so torch.optim.Adam never creates state for them."""
    return FIT_LAYERS_SHARED + (("za_trg1", "za_trg2") if use_trg else ("za_src1", "za_src2"))


def _reparam(mu, logvar, eps):
    """mobody_module.py:237-243 in training mode, with the randn_like draw injected."""
    return mu + eps * torch.exp(0.5 * logvar)


def _kl(mu, logvar):
    """get_kl_loss, mobody_dynamics.py:330-333."""
    return 0.05 * (-0.5 * (1 + logvar - mu.pow(2) - logvar.exp()).mean(dim=(1, 2))).sum()


def fit_losses(p, s, a, ns, r, eps_latent, eps_next, use_trg, encoder_loss_coef=1.0):
    """The loss of one learn() mini-batch for the default configuration (no_vae = 0, latent_reward = 0,
    inverse_sep_reward_loss = 0): encoder_loss (:300-329) + transition_loss (:336-347) + reward_loss (:349-386),
    combined as in :616-641.  s, ns [E,B,S]; a [E,B,A]; r [E,B,1]; eps_latent [6,E,B,16] = the reparameterize draws in
    call order; eps_next [E,B,S] = randn_like(mean) of reward_loss.  Returns (loss, transition, encoder, recon, kl, reward)."""
    def enc(x, k):
        mu, lv = encode_state(p, x)
        return _reparam(mu, lv, eps_latent[k]), mu, lv
    # encoder_loss
    z1, mu_s, lv_s = enc(s, 0)
    rec_s = encode_transition(p, z1)                                        # encoder_decoder(state)
    z2, mu_n, lv_n = enc(ns, 1)
    rec_n = encode_transition(p, z2)                                        # encoder_decoder(next_state)
    recon = ((rec_s - s) ** 2).mean(dim=(1, 2)).sum() + ((rec_n - ns) ** 2).mean(dim=(1, 2)).sum()
    kl = _kl(mu_s, lv_s) + _kl(mu_n, lv_n)
    z3, _, _ = enc(s, 2)
    zl = z3 + encode_action(p, z3, a, use_trg)
    with torch.no_grad():
        z4, _, _ = enc(ns, 3)
    encoder = 100 * recon + kl + ((zl - z4) ** 2).mean(dim=(1, 2)).sum()
    # transition_loss
    z5, _, _ = enc(s, 4)
    mean5 = encode_transition(p, z5 + encode_action(p, z5, a, use_trg))
    transition = ((mean5 - ns) ** 2).mean(dim=(1, 2)).sum()
    # reward_loss
    z6, _, _ = enc(s, 5)
    mean6 = encode_transition(p, z6 + encode_action(p, z6, a, use_trg))
    fake = mean6 + eps_next * torch.std(mean6, dim=0, keepdim=True)
    r1, _ = encode_reward(p, s, a, fake)
    r2, _ = encode_reward(p, s, a, ns)
    reward = ((r1 - r) ** 2).mean(dim=(1, 2)).sum() + ((r2 - r) ** 2).mean(dim=(1, 2)).sum()
    if not use_trg:
        reward = 0.01 * reward
    loss = transition + (5 if use_trg else 1) * encoder_loss_coef * encoder + reward
    return loss, transition, encoder, recon, kl, reward


def fit_step(p, m, v, t, s, a, ns, r, eps_latent, eps_next, use_trg, lr=1e-3, encoder_loss_coef=1.0):
    """One optimiser step of learn() (:643-647): autograd of fit_losses + torch.optim.Adam on the layers that received a
    gradient.  p, m, v: dicts of tensors updated in place; t: dict layer name -> step count after this step."""
    names = [n + sfx for n in fit_trained_layers(use_trg) for sfx in (".weight", ".bias")]
    leaves = {k: p[k].clone().requires_grad_(True) for k in names}
    q = dict(p); q.update(leaves)
    out = fit_losses(q, s, a, ns, r, eps_latent, eps_next, use_trg, encoder_loss_coef)
    grads = torch.autograd.grad(out[0], [leaves[k] for k in names])
    for k, g in zip(names, grads):
        adam_update({k: p[k]}, {k: g}, m, v, t[k.split(".")[0]], lr)
    return [float(x.detach()) for x in out], dict(zip(names, grads))
