"""MOBODY agent: fused model rollout + Q-weighted behaviour-cloning update.

Mirror of algo/offline_offline/mobody.py (``MLPNetwork``, ``Policy``, ``DoubleQFunc``,
``ValueFunc``, ``MOBODY.select_action / rollout / train / save / load``).  The nn.Modules keep
the reference's names so ``_actor`` / ``_critic`` checkpoints (mobody.py:584-594) load unchanged;
they only own parameters — the hot path hands their device pointers to the CUDA kernels.
"""
import copy
import ctypes as C
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn

from . import _ffi
from .buffer import ReplayBuffer
from .dynamics import StepWorkspace


class MLPNetwork(nn.Module):                                   # mobody.py:35-48
    def __init__(self, input_dim, output_dim, hidden_size=256):
        super().__init__()
        self.network = nn.Sequential(nn.Linear(input_dim, hidden_size), nn.ReLU(),
                                     nn.Linear(hidden_size, hidden_size), nn.ReLU(),
                                     nn.Linear(hidden_size, output_dim))

    def forward(self, x):
        return self.network(x)


class ValueFunc(nn.Module):                                    # mobody.py:50-57
    def __init__(self, state_dim, action_dim, hidden_size=256):
        super().__init__()
        self.network = MLPNetwork(state_dim, 1, hidden_size)

    def forward(self, state):
        return self.network(state)


class Policy(nn.Module):                                       # mobody.py:60-72
    def __init__(self, state_dim, action_dim, max_action, hidden_size=256):
        super().__init__()
        self.action_dim, self.max_action = action_dim, max_action
        self.network = MLPNetwork(state_dim, action_dim, hidden_size)

    def forward(self, x):
        """tanh(MLP(x)) * max_action through the CUDA policy kernel (no autograd; training uses the
        fused actor update)."""
        x = x.contiguous().float()
        out = torch.empty(x.shape[0], self.action_dim, dtype=torch.float32, device=x.device)
        mp, _keep = _ffi.mlp_params(self.network)
        _ffi.check(_ffi.lib().mobody_policy_forward(_ffi.ptr(x), x.shape[0], x.shape[1], self.action_dim, C.byref(mp),
                                                    float(self.max_action), _ffi.ptr(out), _ffi.stream_ptr(x.device)))
        return out


class DoubleQFunc(nn.Module):                                  # mobody.py:74-83
    def __init__(self, state_dim, action_dim, hidden_size=256):
        super().__init__()
        self.network1 = MLPNetwork(state_dim + action_dim, 1, hidden_size)
        self.network2 = MLPNetwork(state_dim + action_dim, 1, hidden_size)


class MOBODY(object):
    def __init__(self, config, device, target_entropy=None):   # mobody.py:91-135
        self.config, self.device = config, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mobody_b200.MOBODY needs a CUDA device (there is no CPU path)")
        if config.get("hidden_sizes", 256) != 256:
            raise NotImplementedError("mobody_b200 kernels are built for hidden_sizes=256")
        self.discount, self.tau = config["gamma"], config["tau"]
        self.update_interval = config.get("update_interval", 2)   # read but unused by the reference too
        S, A = config["state_dim"], config["action_dim"]
        self.fake_replay_buffer = ReplayBuffer(S, A, self.device)
        self.penalty_type = config.get("penalty_type", "par")
        self.total_it = 0
        self.q_funcs = DoubleQFunc(S, A).to(self.device)
        self.target_q_funcs = copy.deepcopy(self.q_funcs)
        for p in self.target_q_funcs.parameters():
            p.requires_grad = False
        self.v_func = ValueFunc(S, A).to(self.device)
        self.policy = Policy(S, A, config["max_action"]).to(self.device)
        self.dynamics = None                                     # injected by the caller (train_mobody.py:888)

    def select_action(self, state, policy, cuda=False):          # mobody.py:138-144
        with torch.no_grad():
            s = _ffi.f32(state, self.device).view(-1, self.config["state_dim"])
            action = policy(s)
            return action.squeeze() if cuda else action.squeeze().cpu().numpy()

    # ------------------------------------------------------------------ rollout
    @torch.no_grad()
    def rollout_device(self, init_obss, rollout_length, use_trg=True, *, eps=None, idx=None, row0=0):
        """T-step imagined rollout entirely on the device (mobody.py:596-657 without its per-step
        D2H copies and host masks).  One host read at the end (the transition counts).

        eps: optional [T,7,B,S] / idx: optional [T,B] injected draws (step t uses the first B_t rows,
        exactly what the reference consumes when fed the same arrays).
        Returns (dict of CUDA tensors, info) or (None, None) when rollout_length == 0."""
        if rollout_length == 0:
            return None, None                                    # mobody.py:602-603
        dyn, dev, T = self.dynamics, self.device, int(rollout_length)
        init_obss = _ffi.f32(init_obss, dev)
        B, S = init_obss.shape
        A = self.config["action_dim"]
        lib, st = _ffi.lib(), _ffi.stream_ptr(dev)
        f = dict(dtype=torch.float32, device=dev)
        obss = torch.empty(T, B, S, **f); obss[0].copy_(init_obss)
        nexts, acts = torch.empty(T, B, S, **f), torch.empty(T, B, A, **f)
        rews, raws = torch.empty(T, B, 1, **f), torch.empty(B, 1, **f)
        pens = torch.full((T, B, 1), float("inf"), **f)          # rows never written fail every `<=` filter
        terms = torch.full((T, B), 0xFF, dtype=torch.uint8, device=dev)   # 0xFF = row not produced
        mean = torch.empty(7, B, S, **f)
        row_ids = torch.empty(T, B, dtype=torch.int64, device=dev)
        row_ids[0] = torch.arange(row0, row0 + B, device=dev)
        counts = torch.zeros(T + 2, dtype=torch.int32, device=dev); counts[0] = B   # [B_0..B_T, M]
        pos = torch.empty(max(T * B, 1), dtype=torch.int32, device=dev)
        scratch = torch.empty(int(lib.mobody_compact_scratch_ints(T * B)), dtype=torch.int32, device=dev)
        if eps is not None:
            eps = _ffi.f32(eps, dev); assert tuple(eps.shape) == (T, 7, B, S)
        if idx is not None:
            idx = torch.as_tensor(np.asarray(idx) if not torch.is_tensor(idx) else idx).to(device=dev, dtype=torch.int64).contiguous()
        ws = StepWorkspace.__new__(StepWorkspace)
        for t in range(T):
            ws.next_obs, ws.reward, ws.raw_reward, ws.penalty = nexts[t], rews[t], raws, pens[t]
            ws.terminal, ws.mean, ws.act = terms[t], mean, acts[t]
            dyn.launch_step(obss[t], None, ws, policy=self.policy.network, max_action=self.policy.max_action,
                            use_trg=use_trg, eps=None if eps is None else eps[t], idx=None if idx is None else idx[t],
                            n_rows_dev=counts[t:t + 1], row_ids=row_ids[t], step=t)
            if t + 1 < T:   # nonterm_mask compaction (mobody.py:635-639), stable order, no host round trip
                _ffi.check(lib.mobody_compact(_ffi.KEEP_U8_ZERO, _ffi.ptr(terms[t]), None, 0.0, B, _ffi.ptr(counts[t:t + 1]),
                                              _ffi.ptr(scratch), _ffi.ptr(pos), _ffi.ptr(counts[t + 1:t + 2]), st))
                _ffi.check(lib.mobody_gather_pos(_ffi.ptr(nexts[t]), S, S, _ffi.ptr(pos), _ffi.ptr(counts[t + 1:t + 2]), B,
                                                 _ffi.ptr(obss[t + 1]), S, st))
                _ffi.check(lib.mobody_gather_pos_i64(_ffi.ptr(row_ids[t]), _ffi.ptr(pos), _ffi.ptr(counts[t + 1:t + 2]), B,
                                                     _ffi.ptr(row_ids[t + 1]), st))
        # concat over steps + penalty filter (mobody.py:641-653): one stable compaction over T*B slots
        if self.config.get("filter_bad_rollout", 1):
            kind, flags, vals, thr = _ffi.KEEP_F32_LE, None, pens, float(self.config["env_filter"])
        else:
            kind, flags, vals, thr = _ffi.KEEP_U8_VALID, terms, None, 0.0
        _ffi.check(lib.mobody_compact(kind, _ffi.ptr(flags), _ffi.ptr(vals), thr, T * B, None, _ffi.ptr(scratch),
                                      _ffi.ptr(pos), _ffi.ptr(counts[T + 1:T + 2]), st))
        valid = terms != 0xFF
        rew_sum = torch.where(valid, rews.view(T, B), torch.zeros((), **f)).double().sum()
        host = torch.cat([counts.double(), rew_sum.view(1)]).cpu()       # the single host read
        n_per_step = [int(v) for v in host[:T]]
        M, num_transitions = int(host[T + 1]), int(sum(n_per_step))
        termf = terms.float()
        out = {}
        mdev = counts[T + 1:T + 2]
        for name, src, w in (("obss", obss, S), ("next_obss", nexts, S), ("actions", acts, A), ("rewards", rews, 1),
                             ("terminals", termf, 1), ("penalty", pens, 1)):
            dst = torch.empty(M, w, **f)
            if M:
                _ffi.check(lib.mobody_gather_pos(_ffi.ptr(src), w, w, _ffi.ptr(pos), _ffi.ptr(mdev), M, _ffi.ptr(dst), w, st))
            out[name] = dst
        info = {"num_transitions": num_transitions, "reward_mean": float(host[T + 2]) / max(num_transitions, 1),
                "rows_per_step": n_per_step, "kept": M}
        return out, info

    def rollout(self, init_obss, rollout_length, use_trg=True, **kw):
        """Reference signature and return convention (mobody.py:596-657): dict of CPU tensors + info."""
        out, info = self.rollout_device(init_obss, rollout_length, use_trg, **kw)
        if out is None:
            return None, None
        if self.config.get("filter_bad_rollout", 1):
            print("filtered rollout", info["kept"], info["num_transitions"])     # mobody.py:653
        return {k: v.cpu() for k, v in out.items()}, {"num_transitions": info["num_transitions"],
                                                      "reward_mean": info["reward_mean"]}

    # ------------------------------------------------------------------ checkpoints
    def save(self, filename):                                    # mobody.py:584-588 (optimizer files: see train step)
        torch.save(self.q_funcs.state_dict(), filename + "_critic")
        torch.save(self.policy.state_dict(), filename + "_actor")

    def load(self, filename):                                    # mobody.py:590-594
        self.q_funcs.load_state_dict(torch.load(filename + "_critic"))
        self.policy.load_state_dict(torch.load(filename + "_actor"))
