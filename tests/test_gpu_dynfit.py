"""Dynamics fitting step on the GPU, through the C ABI (mobody_dynfit_step), against
 (1) fixtures the UNMODIFIED reference's learn() produced under scripted randn_like draws (oracle/make_golden.py: gen_dynfit),
 (2) the oracle (autograd restatement, oracle/mobody_oracle.py: fit_step) on fresh seeded inputs at the reference's batch size.
Tolerances: loss scalars and gradients 1e-4 relative (helpers.rel_err: relative to the tensor's own scale); Adam second
moments 2e-4 (squares of the gradients); parameters after Adam: see adam_close (test_gpu_train.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import cuda_dynamics, rel_err
from oracle import mobody_oracle as M
from oracle.make_golden import HEALTHY, fit_sample_index
from test_gpu_train import adam_close
from test_oracle_golden import fit_replay

pytestmark = pytest.mark.gpu


def mostly_close(got, want, tol, frac, cap, tag):
    """All but a fraction ``frac`` of the elements within ``tol`` (helpers.rel_err metric per element), every element within ``cap``."""
    got, want = np.asarray(got, np.float64).ravel(), np.asarray(want, np.float64).ravel()
    rel = np.abs(got - want) / (np.abs(want) + np.mean(np.abs(want)) + 1e-30)
    assert np.mean(rel > tol) <= frac and rel.max() <= cap, (tag, float(np.mean(rel > tol)), float(rel.max()))


def _state(dyn, name):
    st = dyn._fit_state()
    lay, sfx = name.rsplit(".", 1)
    k = 0 if sfx == "weight" else 1
    return getattr(getattr(dyn.model, lay), sfx).detach(), st["m"][lay][k], st["v"][lay][k]


@pytest.mark.parametrize("name", ["dynfit_S11A3_B24", "dynfit_S17A6_B40"])
def test_fit_step_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    S, A = int(g["S"]), int(g["A"])
    env = "hopper" if S == 11 else "walker2d"
    dyn, _ = cuda_dynamics(S, A, int(g["seed"]), env, 1.0)
    first_m = {}

    def step(c, use_trg, batch, el, en, t):
        obs, act, nobs, rew = (torch.from_numpy(x).cuda() for x in batch)
        sc = dyn.fit_batch(use_trg, obs, act, nobs, rew, eps_latent=el, eps_next=en).cpu().numpy()
        want = g[f"c{c}_losses"]
        for k in range(5):
            assert abs(sc[k] - want[k]) <= 1e-4 * abs(want[k]), (c, k, sc[k], want[k])
        if c == 0:
            for lay in M.fit_trained_layers(use_trg):
                for sfx in ("weight", "bias"):
                    first_m[f"{lay}.{sfx}"] = _state(dyn, f"{lay}.{sfx}")[1].cpu().numpy().copy()
    counts = fit_replay(g, step)
    n_calls = len(g["calls"])
    for lay in M.dynamics_layer_shapes(S, A):
        for sfx in ("weight", "bias"):
            k = f"{lay}.{sfx}"
            p, m, v = (x.cpu().numpy().ravel() for x in _state(dyn, k))
            idx = fit_sample_index(p.size)
            if not bool(g["has_state_" + k]):          # never trained in the reference: untouched here, bit for bit
                assert np.array_equal(p[idx], g["p_" + k]) and not m.any() and not v.any(), k
                continue
            assert dyn._fit_state()["t"][lay] == counts[lay] == int(g["t_" + k])
            if "first_m_" + k in g.files:              # exp_avg after the first step = 0.1 * gradient: the backward pass itself
                assert rel_err(first_m[k].ravel()[idx], g["first_m_" + k]) <= 1e-4, (k, rel_err(first_m[k].ravel()[idx], g["first_m_" + k]))
            # later steps start from parameters that differ where Adam's sign(g) step flipped on a ~0 gradient element
            if n_calls == 1:
                assert rel_err(m[idx], g["m_" + k]) <= 1e-4, (k, rel_err(m[idx], g["m_" + k]))
                assert rel_err(v[idx], g["v_" + k]) <= 2e-4, (k, rel_err(v[idx], g["v_" + k]))
            else:       # a flipped element moved by 2 * lr: the following gradients differ at the 1e-3 level in a few places
                mostly_close(m[idx], g["m_" + k], 1e-3, 0.02, 0.1, "m " + k)
                mostly_close(v[idx], g["v_" + k], 2e-3, 0.02, 0.2, "v " + k)
            adam_close(p[idx], g["p_" + k], 1e-3, n_calls, k)


@pytest.mark.parametrize("S,A,B,use_trg", [(17, 6, 256, True), (27, 8, 200, False), (11, 3, 37, True)])
def test_fit_step_matches_oracle_at_size(S, A, B, use_trg):
    """The reference's batch size (256 rows per member, mobody_dynamics.py:739) and ragged sizes, one step, vs the oracle."""
    env = "hopper" if S == 11 else "walker2d"
    dyn, p = cuda_dynamics(S, A, 7, env, 1.0)
    rng = np.random.default_rng(S * 1000 + B)
    s = (HEALTHY[env](S)[None, None, :] + 0.3 * rng.standard_normal((7, B, S))).astype(np.float32)
    a = rng.uniform(-1, 1, (7, B, A)).astype(np.float32)
    ns = (s + 0.1 * rng.standard_normal((7, B, S))).astype(np.float32)
    r = rng.standard_normal((7, B, 1)).astype(np.float32)
    el = rng.standard_normal((6, 7, B, 16)).astype(np.float32); en = rng.standard_normal((7, B, S)).astype(np.float32)
    names = [n + sfx for n in M.fit_trained_layers(use_trg) for sfx in (".weight", ".bias")]
    m = {k: torch.zeros_like(p[k]) for k in names}; v = {k: torch.zeros_like(p[k]) for k in names}
    torch.set_num_threads(8)
    want, grads = M.fit_step(p, m, v, {n: 1 for n in M.fit_trained_layers(use_trg)}, *(torch.from_numpy(x) for x in (s, a, ns, r)),
                             torch.from_numpy(el), torch.from_numpy(en), use_trg)
    got = dyn.fit_batch(use_trg, *(torch.from_numpy(x).cuda() for x in (s, a, ns, r)), eps_latent=el, eps_next=en).cpu().numpy()
    for k in range(6):
        assert abs(got[k] - want[k]) <= 1e-4 * abs(want[k]), (k, got[k], want[k])
    for k in names:
        _, gm, gv = _state(dyn, k)
        assert rel_err(gm.cpu().numpy() * 10.0, grads[k].numpy()) <= 1e-4, (k, rel_err(gm.cpu().numpy() * 10.0, grads[k].numpy()))
        assert rel_err(gv.cpu().numpy(), v[k].numpy()) <= 2e-4, k
        adam_close(_state(dyn, k)[0].cpu().numpy(), p[k].numpy(), 1e-3, 1, k)
    other = ("za_src1", "za_src2") if use_trg else ("za_trg1", "za_trg2")
    for lay in other:                                  # the other domain's action encoder is not touched
        assert torch.equal(getattr(dyn.model, lay).weight.detach().cpu(), p[lay + ".weight"])


def test_fit_window_of_epoch_tensor_and_philox_noise():
    """A batch that is a row window of [7, N, .] epoch tensors gives the same step as the copied-out batch; Philox noise is
    deterministic per (seed, draw) and fresh per call."""
    S, A, N, B, lo = 17, 6, 300, 64, 100
    rng = np.random.default_rng(3)
    data = [torch.from_numpy(rng.standard_normal((7, N, w)).astype(np.float32)).cuda() for w in (S, A, S, 1)]
    data[0][..., 0] += 1.25
    el = torch.from_numpy(rng.standard_normal((6, 7, B, 16)).astype(np.float32)); en = torch.from_numpy(rng.standard_normal((7, B, S)).astype(np.float32))
    d1, _ = cuda_dynamics(S, A, 9, "walker2d", 1.0)
    d2, _ = cuda_dynamics(S, A, 9, "walker2d", 1.0)
    s1 = d1.fit_batch(True, *data, rows=B, lo=lo, eps_latent=el, eps_next=en)
    s2 = d2.fit_batch(True, *(x[:, lo:lo + B].contiguous() for x in data), eps_latent=el, eps_next=en)
    assert torch.equal(s1, s2)
    for lay in M.fit_trained_layers(True):
        assert torch.equal(getattr(d1.model, lay).weight, getattr(d2.model, lay).weight), lay
    d3, _ = cuda_dynamics(S, A, 9, "walker2d", 1.0)
    d4, _ = cuda_dynamics(S, A, 9, "walker2d", 1.0)
    a0, b0 = d3.fit_batch(True, *data, rows=B, lo=lo).clone(), d4.fit_batch(True, *data, rows=B, lo=lo).clone()
    assert torch.equal(a0, b0) and torch.isfinite(a0[:6]).all()                    # same (seed, draw): same noise
    a1 = d3.fit_batch(True, *data, rows=B, lo=lo)
    assert not torch.equal(a0, a1)                                                   # next draw (and moved parameters)
    assert not torch.equal(a0, s1)                                                   # Philox noise != the injected tensors


def test_learn_validate_train_api():
    """learn() / validate() / select_elites() / train() of the reference's dynamics object (mobody_dynamics.py:594-653,
    731-978, 1114-1156) on a small learnable synthetic problem: the transition loss falls, elites are picked, load_save ran."""
    import mobody_b200 as mb
    S, A = 11, 3
    torch.manual_seed(0)
    m = mb.MOBODYModule(S, A, 256, 7, 5, device="cuda", config={"mopo": 0, "latent_reward": 0})
    cfg = {"encoder_loss_coef": 1, "no_vae": 0, "latent_reward": 0, "inverse_sep_reward_loss": 0, "train_together": 0, "train_with_src_threshold": 1}
    dyn = mb.MOBODYEnsembleDynamics(cfg, m, torch.optim.Adam(m.parameters(), lr=1e-3), None, mb.get_termination_fn("hopper-medium-v2"))
    rng = np.random.default_rng(1)

    def data(n, shift):
        s = rng.standard_normal((n, S)).astype(np.float32)
        a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
        ns = (0.9 * s + shift * np.pad(a, ((0, 0), (0, S - A)))).astype(np.float32)
        r = (s[:, :1] + a[:, :1]).astype(np.float32)
        return tuple(torch.from_numpy(x) for x in (s, a, ns, r))
    src, trg = data(1200, 0.5), data(700, 0.8)
    idx = torch.randint(700, (7, 700))
    first = dyn.learn(True, *(x[idx] for x in trg), 256, 0.01)          # 3 mini-batches, the last one ragged (188 rows)
    assert len(first) == 5 and all(np.isfinite(first)) and dyn.total_steps == 3
    for _ in range(6):
        last = dyn.learn(True, *(x[idx] for x in trg), 256, 0.01)
    assert last[1] < 0.9 * first[1], (first, last)                       # transition loss falls
    tl, rl = dyn.validate(True, *(x[:200] for x in trg))
    assert len(tl) == 7 and len(rl) == 7 and dyn.model.training
    assert dyn.select_elites([3.0, 1.0, 2.0, 9.0, 0.5, 7.0, 8.0]) == [4, 1, 2, 0, 5]
    dyn.train(src, trg, max_epochs=2, batch_size=256)
    assert len(dyn.model.elites) == 5 and not dyn.model.training
    assert torch.equal(dyn.model.zs1.weight, dyn.model.zs1.saved_weight)  # load_save (:975)
    nobs, rew, term, info = dyn.step(src[0][:64], src[1][:64])           # the fitted model rolls (packed image follows by checksum)
    assert torch.isfinite(nobs).all() and nobs.shape == (64, S)


_FIT_DIGEST = r"""
import hashlib, sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from helpers import cuda_dynamics
S, A, N, B = 17, 6, 700, 256
torch.manual_seed(0)          # the layers the recipe does not cover (action decoders, saved_* copies) are torch-initialised
rng = np.random.default_rng(4)
data = [torch.from_numpy(rng.standard_normal((7, N, w)).astype(np.float32)).cuda() for w in (S, A, S, 1)]
dyn, _ = cuda_dynamics(S, A, 5, "walker2d", 1.0)
for ep in range(2):
    for trg in (True, False):
        out = dyn.learn(trg, *data, B, 0.01)
torch.cuda.synchronize()
h = hashlib.sha256()
for v in dyn.model.state_dict().values():
    h.update(v.detach().cpu().numpy().tobytes())
h.update(np.asarray(out, np.float64).tobytes())
print("DIGEST", h.hexdigest())
"""


def test_side_stream_and_launch_mode_do_not_change_a_bit():
    """The weight gradients run on a side stream and the chain is launched with programmatic dependent launch: neither may
    change the result (12 optimiser steps over ragged mini-batches, both domains) -- a missing dependency would."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def digest(**env):
        r = subprocess.run([sys.executable, "-c", _FIT_DIGEST], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300, cwd=root)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return [ln.split()[1] for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][0]
    base = digest(MOBODY_DYNFIT_SIDE="1", MOBODY_PDL="1")
    assert base == digest(MOBODY_DYNFIT_SIDE="0", MOBODY_PDL="1") == digest(MOBODY_DYNFIT_SIDE="1", MOBODY_PDL="0")
