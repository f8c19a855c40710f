"""MOBODY agent: fused model rollout + Q-weighted behaviour-cloning update.

Mirror of algo/offline_offline/mobody.py (``MLPNetwork``, ``Policy``, ``DoubleQFunc``,
``ValueFunc``, ``MOBODY.select_action / rollout / train / save / load``).  The nn.Modules keep
the reference's names so ``_actor`` / ``_critic`` checkpoints (mobody.py:584-594) load unchanged;
they only own parameters — the hot path hands their device pointers to the CUDA kernels.
"""
import copy
import ctypes as C

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


class Classifier(nn.Module):                                   # mobody.py:11-33 (domain classifier of DARA / DARC)
    def __init__(self, state_dim, action_dim, hidden_size=256, gaussian_noise_std=1.0):
        super().__init__()
        self.action_dim, self.gaussian_noise_std = action_dim, gaussian_noise_std
        self.sa_classifier = MLPNetwork(state_dim + action_dim, 2, hidden_size)
        self.sas_classifier = MLPNetwork(2 * state_dim + action_dim, 2, hidden_size)



_SM_COUNT = {}


def _wgrad_splits(n_rows, tiles_per_launch, device, sm_count=None):
    """Row splits of the weight-gradient launches.  A launch runs ``tiles * nsplit`` CTAs of 128 x 64 outputs, two per
    SM, each walking its rows in chunks of 32: pick the split count (<= 64) that minimises rounds x chunks per CTA
    summed over the launches of an update (tiles: critic 2 x (8 + 2), actor 8 + 2), i.e. avoid a nearly empty last
    round.  Any value gives the same result up to fp32 summation order; the order is fixed for a given value."""
    if sm_count is None:
        key = str(device)
        if key not in _SM_COUNT:
            _SM_COUNT[key] = torch.cuda.get_device_properties(device).multi_processor_count
        sm_count = _SM_COUNT[key]
    slots, chunks = 2 * sm_count, (n_rows + 31) // 32
    best, best_cost = 1, None
    for n in range(1, min(64, chunks) + 1):
        per = (chunks + n - 1) // n
        cost = sum(((t * n + slots - 1) // slots) for t in tiles_per_launch) * per
        if best_cost is None or cost < best_cost:
            best, best_cost = n, cost
    return best

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
        self.classifier = Classifier(S, A, 256, config.get("gaussian_noise_std", 1.0)).to(self.device)   # :134
        self.dynamics = None                                     # injected by the caller (train_mobody.py:888)
        # Adam moments of the fused train step (torch.optim.Adam equivalents of mobody.py:127-131)
        z = lambda mlp: [torch.zeros_like(t) for t in _ffi.mlp_tensors(mlp)]      # noqa: E731
        self._adam = {"pi": (z(self.policy.network), z(self.policy.network)),
                      "q1": (z(self.q_funcs.network1), z(self.q_funcs.network1)),
                      "q2": (z(self.q_funcs.network2), z(self.q_funcs.network2))}
        self._adam["sas"] = (z(self.classifier.sas_classifier), z(self.classifier.sas_classifier))
        self._adam["sa"] = (z(self.classifier.sa_classifier), z(self.classifier.sa_classifier))
        self._t_q = self._t_pi = self._t_cls = 0
        self._cls_ws, self._cls_scalars = None, torch.zeros(2, dtype=torch.float32, device=self.device)
        self._train_ws = None
        self._scalars = torch.zeros(16, dtype=torch.float32, device=self.device)
        self._roll_ws = {}                                       # (T, B, S, A) -> rollout scratch
        self._host_slabs, self._host_turn = [None, None], 0      # pinned staging of rollout() results
        self._pipe_streams, self._pipe_hdr = None, None          # side streams / pinned header of the pipelined rollout()
        self._pipe_plan_cache = None                             # (key, per-chunk launch plan)

    def select_action(self, state, policy, cuda=False):          # mobody.py:138-144
        with torch.no_grad():
            s = _ffi.f32(state, self.device).view(-1, self.config["state_dim"])
            action = policy(s)
            return action.squeeze() if cuda else action.squeeze().cpu().numpy()

    # ------------------------------------------------------------------ rollout
    def _rollout_workspace(self, T, B, S, A, slot=0):
        """Device scratch of mobody_rollout for (T, B): allocated once and reused (the result slab is not part of it).
        ``slot`` separates workspaces of rollouts that are in flight at the same time on different streams."""
        key = (T, B, S, A, slot)
        ws = self._roll_ws.get(key)
        if ws is None:
            dev, f = self.device, dict(dtype=torch.float32, device=self.device)
            ws = dict(obss=torch.empty(T, B, S, **f), acts=torch.empty(T, B, A, **f), nexts=torch.empty(T, B, S, **f),
                      rews=torch.empty(T, B, 1, **f), pens=torch.empty(T, B, 1, **f), raw=torch.empty(B, 1, **f),
                      terms=torch.empty(T, B, dtype=torch.uint8, device=dev), mean=torch.empty(7, B, S, **f),
                      row_ids=torch.empty(T, B, dtype=torch.int64, device=dev),
                      counts=torch.zeros(T + 2, dtype=torch.int32, device=dev),
                      pos=torch.empty(max(T * B, 1), dtype=torch.int32, device=dev),
                      scratch=torch.empty(int(_ffi.lib().mobody_compact_scratch_ints(T * B)), dtype=torch.int32, device=dev),
                      stats=torch.zeros(2 + 2 * 148, dtype=torch.float64, device=dev),
                      ticket=torch.zeros(1, dtype=torch.int32, device=dev))
            if len(self._roll_ws) >= 8:                       # a handful of shapes per run (50 000 / 2 000 starts)
                self._roll_ws.pop(next(iter(self._roll_ws)))
            self._roll_ws[key] = ws
        return ws

    @torch.no_grad()
    def rollout_device(self, init_obss, rollout_length, use_trg=True, *, eps=None, idx=None, row0=0, out_packed=None,
                       sync=True, ws_slot=0):
        """T-step imagined rollout entirely on the device (mobody.py:596-657 without its per-step D2H copies and
        host masks): ONE C-ABI call (mobody_rollout) enqueues the T fused steps, the compactions between them, the
        concatenation + penalty filter and the packing of the kept transitions.

        eps: optional [T,7,B,S] / idx: optional [T,B] injected draws (step t uses the first B_t rows, exactly what
        the reference consumes when fed the same arrays).
        out_packed: optional preallocated [>= T*B, 2S+A+3] CUDA tensor receiving the kept transitions as rows
        [obs | act | next_obs | reward | terminal | penalty] (the slab the multi-GPU all-gather ships).
        sync=True: one host read at the end (row counts + reward sum); the returned dict holds [M, .] column views of
        the slab.  sync=False: nothing is read back; the dict holds capacity-sized views (rows >= kept are
        unspecified) and info carries device tensors only (``kept_dev``, ``counts_dev``, ``stats_dev`` — views of the
        cached workspace: consume them on the stream before the next rollout of the same shape overwrites them).
        Returns (dict of CUDA tensors, info) or (None, None) when rollout_length == 0."""
        if rollout_length == 0:
            return None, None                                    # mobody.py:602-603
        dyn, dev, T = self.dynamics, self.device, int(rollout_length)
        init_obss = _ffi.f32(init_obss, dev)
        B, S = init_obss.shape
        A = self.config["action_dim"]
        W = 2 * S + A + 3
        ws = self._rollout_workspace(T, B, S, A, ws_slot)
        packed = out_packed if out_packed is not None else torch.empty(max(T * B, 1), W, dtype=torch.float32, device=dev)
        assert packed.shape[1] == W and packed.shape[0] >= T * B and packed.is_contiguous()
        if eps is not None:
            eps = _ffi.f32(eps, dev); assert tuple(eps.shape) == (T, 7, B, S)
        if idx is not None:
            idx = torch.as_tensor(np.asarray(idx) if not torch.is_tensor(idx) else idx).to(device=dev, dtype=torch.int64).contiguous()
        counts = ws["counts"]
        if B > 0:
            d = _ffi.RolloutDesc()
            keep = dyn.fill_step_desc(d.step, B, S, dev, policy=self.policy.network, max_action=self.policy.max_action,  # noqa: F841
                                      use_trg=use_trg)
            d.step.obs, d.step.mean, d.step.raw_reward = _ffi.ptr(init_obss), _ffi.ptr(ws["mean"]), _ffi.ptr(ws["raw"])
            d.step.step, d.step.row0 = 0, int(row0)
            d.T, d.filter_bad_rollout = T, int(bool(self.config.get("filter_bad_rollout", 1)))
            d.env_filter = float(self.config.get("env_filter", 0.0))
            d.eps_all, d.idx_all = _ffi.ptr(eps), _ffi.ptr(idx)
            for k in ("obss", "acts", "nexts", "rews", "pens", "terms", "row_ids", "counts", "pos", "scratch", "stats", "ticket"):
                setattr(d, k, _ffi.ptr(ws[k]))
            d.packed = _ffi.ptr(packed)
            _ffi.check(_ffi.lib().mobody_rollout(C.byref(d), _ffi.stream_ptr(dev)))
        else:
            counts.zero_(); ws["stats"][:2].zero_()
        mdev = counts[T + 1:T + 2]
        if sync:
            host = torch.cat([counts.double(), ws["stats"][:2]]).cpu()       # the single host read
            n_per_step = [int(v) for v in host[:T]]
            M, num_transitions, rew_sum = int(host[T + 1]), int(host[T + 3]), float(host[T + 2])
            info = {"num_transitions": num_transitions, "reward_mean": rew_sum / max(num_transitions, 1),
                    "rows_per_step": n_per_step, "kept": M, "kept_dev": mdev}
        else:
            M = T * B
            info = {"kept_dev": mdev, "counts_dev": counts, "stats_dev": ws["stats"][:2], "capacity": T * B}
        out, c0 = {}, 0
        for name, w in (("obss", S), ("actions", A), ("next_obss", S), ("rewards", 1), ("terminals", 1), ("penalty", 1)):
            out[name] = packed[:M, c0:c0 + w]
            c0 += w
        info["packed"] = packed
        return out, info

    def _host_slab(self, rows, W):
        """Pinned staging slab for rollout() results, double-buffered: the CPU tensors a rollout() call returns stay
        valid until the second-next rollout() call (the reference's only consumer, add_batch, copies immediately)."""
        self._host_turn ^= 1
        slab = self._host_slabs[self._host_turn]
        if slab is None or slab.shape[0] < rows or slab.shape[1] != W:
            slab = torch.empty(max(rows, 1), W, dtype=torch.float32, pin_memory=True)
            self._host_slabs[self._host_turn] = slab
        return slab

    PIPE_ROWS = 2 * 148 * 128     # start states per pipelined chunk: two full waves of 128-row tiles on 148 SMs

    def rollout(self, init_obss, rollout_length, use_trg=True, **kw):
        """Reference signature and return convention (mobody.py:596-657): dict of CPU tensors + info.
        The kept rows are copied D2H once, into a pinned staging slab; the dict values are column views of it.
        One-step rollouts of many start states are pipelined in chunks on two side streams (H2D of chunk c+1 and D2H
        of chunk c-1 overlap the kernels of chunk c); rows are independent and Philox is keyed on the global row id,
        so the result is identical to the unchunked call (row order included)."""
        T = int(rollout_length)
        if T == 0:
            return None, None
        S, A = self.config["state_dim"], self.config["action_dim"]
        W = 2 * S + A + 3
        B = int(init_obss.shape[0])
        if T == 1 and B >= 2 * self.PIPE_ROWS and not kw:
            res_host, n_tr, rsum, M = self._rollout_pipelined(init_obss, use_trg, S, A, W)
        else:
            out, info = self.rollout_device(init_obss, T, use_trg, **kw)
            M, packed = info["kept"], info["packed"]
            res_host = self._host_slab(M, W)[:M]
            res_host.copy_(packed[:M], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            n_tr, rsum = info["num_transitions"], info["reward_mean"] * info["num_transitions"]
        if self.config.get("filter_bad_rollout", 1):
            print("filtered rollout", M, n_tr)                                   # mobody.py:653
        res, c0 = {}, 0
        for name, w in (("obss", S), ("actions", A), ("next_obss", S), ("rewards", 1), ("terminals", 1), ("penalty", 1)):
            res[name] = res_host[:, c0:c0 + w]
            c0 += w
        return res, {"num_transitions": n_tr, "reward_mean": rsum / max(n_tr, 1)}

    PIPE_FIRST = 148 * 128        # first chunk: one wave, so the first kernel starts after a short H2D

    def _pipe_bounds(self, B):
        """Chunk boundaries of a pipelined one-step rollout: a one-wave first chunk (little H2D exposed before the first
        kernel), two-wave chunks after it, and the sub-wave remainder as its own last chunk (little D2H exposed after the
        last kernel).  Whole waves per chunk: chunking adds no partially filled round to the step kernel."""
        bounds, lo = [0], min(self.PIPE_FIRST, B)
        while lo < B:
            bounds.append(lo)
            lo += self.PIPE_ROWS
        bounds.append(B)
        wave = 148 * 128
        tail = (B - bounds[-2]) % wave
        if 0 < tail < B - bounds[-2] and tail <= wave // 2:
            bounds.insert(-1, B - tail)                           # split the sub-wave remainder off the last chunk
        return bounds

    def _pipe_plan(self, B, S, A, W, use_trg, bounds):
        """Per-chunk launch plan of the pipelined rollout (device input / output buffers, workspace, a filled
        mobody_rollout_desc, an event), cached while nothing it points to changes: a chunk then costs one H2D
        enqueue, ONE C-ABI call and two 8-byte D2H enqueues of host time."""
        dev, dyn = self.device, self.dynamics
        probe = _ffi.StepDesc()
        # also (re)builds the packed weight images on the caller's stream, before the side streams fork from it
        keep0 = dyn.fill_step_desc(probe, 1, S, dev, policy=self.policy.network, max_action=self.policy.max_action, use_trg=use_trg)
        filt, env_filter = int(bool(self.config.get("filter_bad_rollout", 1))), float(self.config.get("env_filter", 0.0))
        key = (tuple(bounds), S, A, bool(use_trg), probe.precision, probe.dyn_pack, probe.policy_pack,
               C.addressof(probe.dyn.contents), C.addressof(probe.policy.contents), probe.elites, probe.n_elites,
               probe.penalty_coef, probe.term_kind, probe.seed, probe.max_action, filt, env_filter)
        if self._pipe_plan_cache is not None and self._pipe_plan_cache[0] == key:
            return self._pipe_plan_cache[1]
        plan = []
        for c in range(len(bounds) - 1):
            lo, hi = bounds[c], bounds[c + 1]
            n = hi - lo
            ws = self._rollout_workspace(1, n, S, A, 1 + (c & 1))
            x = torch.empty(n, S, dtype=torch.float32, device=dev)
            packed = torch.empty(n, W, dtype=torch.float32, device=dev)
            d = _ffi.RolloutDesc()
            keep = dyn.fill_step_desc(d.step, n, S, dev, policy=self.policy.network, max_action=self.policy.max_action, use_trg=use_trg)
            d.step.obs, d.step.mean, d.step.raw_reward = _ffi.ptr(x), _ffi.ptr(ws["mean"]), _ffi.ptr(ws["raw"])
            d.step.step, d.step.row0 = 0, int(lo)
            d.T, d.filter_bad_rollout, d.env_filter = 1, filt, env_filter
            d.eps_all, d.idx_all = None, None
            for k in ("obss", "acts", "nexts", "rews", "pens", "terms", "row_ids", "counts", "pos", "scratch", "stats", "ticket"):
                setattr(d, k, _ffi.ptr(ws[k]))
            d.packed = _ffi.ptr(packed)
            plan.append(dict(lo=lo, hi=hi, x=x, packed=packed, desc=d, ref=C.byref(d), keep=(keep, keep0, ws),
                             kept_dev=ws["counts"][2:3], stats_dev=ws["stats"][:2], ev=torch.cuda.Event()))
        self._pipe_plan_cache = (key, plan)
        return plan

    def _rollout_pipelined(self, init_obss, use_trg, S, A, W):
        dev, lib = self.device, _ffi.lib()
        if not torch.is_tensor(init_obss):
            init_obss = torch.as_tensor(np.asarray(init_obss, dtype=np.float32))
        B = init_obss.shape[0]
        bounds = self._pipe_bounds(B)
        n = len(bounds) - 1
        if self._pipe_streams is None:
            self._pipe_streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)]   # 2 compute + 1 D2H
        if self._pipe_hdr is None or self._pipe_hdr[0].shape[0] < n:
            self._pipe_hdr = (torch.zeros(max(n, 8), 1, dtype=torch.int32).pin_memory(),      # kept rows per chunk
                              torch.zeros(max(n, 8), 2, dtype=torch.float64).pin_memory())    # [reward sum, produced]
        hdr_kept, hdr_stats = self._pipe_hdr
        cur = torch.cuda.current_stream(dev)
        plan = self._pipe_plan(B, S, A, W, use_trg, bounds)
        ready = torch.cuda.Event(); ready.record(cur)
        host = self._host_slab(B, W)
        for c, ch in enumerate(plan):
            st = self._pipe_streams[c & 1]
            with torch.cuda.stream(st):
                st.wait_event(ready)
                ch["x"].copy_(init_obss[ch["lo"]:ch["hi"]], non_blocking=True)
                _ffi.check(lib.mobody_rollout(ch["ref"], C.c_void_p(st.cuda_stream)))
                # [kept | reward sum, produced] of this chunk -> pinned header rows (stream-ordered 4- and 16-byte D2H)
                hdr_kept[c].copy_(ch["kept_dev"], non_blocking=True)
                hdr_stats[c].copy_(ch["stats_dev"], non_blocking=True)
                ch["ev"].record(st)
        off, n_tr, rsum = 0, 0, 0.0
        copy_stream = self._pipe_streams[2]                       # dedicated copy stream: never queued behind a later chunk's kernels
        for c, ch in enumerate(plan):
            ch["ev"].synchronize()                                # chunk c is done; later chunks keep the GPU busy
            m = int(hdr_kept[c, 0]); rsum += float(hdr_stats[c, 0]); n_tr += int(hdr_stats[c, 1])
            with torch.cuda.stream(copy_stream):
                host[off:off + m].copy_(ch["packed"][:m], non_blocking=True)
            off += m
        copy_stream.synchronize()
        return host[:off], n_tr, rsum, off

    # ------------------------------------------------------------------ DARA domain classifier
    def classifier_step_on_rows(self, rows, label, *, noise_sas=None, noise_sa=None):
        """One fused classifier update (forward on noise-perturbed inputs, cross-entropy on the softmaxed outputs,
        backward, Adam) on packed batch rows [N, RW] with int32 domain labels [N].  Asynchronous; the two losses land
        in ``self._cls_scalars`` = [loss_sa, loss_sas] (device).  mobody.py:11-33, 146-181."""
        cfg = self.config
        S, A = cfg["state_dim"], cfg["action_dim"]
        N = rows.shape[0]
        nsplit = _wgrad_splits(N, (16 + 2 * ((2 * S + A + 63) // 64) + 2 * ((S + A + 63) // 64),), self.device)
        lib = _ffi.lib()
        need = int(lib.mobody_classifier_workspace_bytes(N, S, A, nsplit))
        if self._cls_ws is None or self._cls_ws.numel() < need:
            self._cls_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._t_cls += 1
        d = _ffi.ClassifierDesc()
        d.rows, d.N, d.S, d.A, d.row_width = _ffi.ptr(rows), N, S, A, rows.shape[1]
        d.label = _ffi.ptr(label)
        # injected draws: the device copies must stay referenced until the launch is enqueued (a dropped temporary's block
        # is handed to the very next allocation by the caching allocator)
        noise_sas = None if noise_sas is None else _ffi.f32(noise_sas, self.device)
        noise_sa = None if noise_sa is None else _ffi.f32(noise_sa, self.device)
        d.noise_sas, d.noise_sa = _ffi.ptr(noise_sas), _ffi.ptr(noise_sa)
        d.noise_std, d.seed, d.draw = float(self.classifier.gaussian_noise_std), int(cfg.get("seed", 0)), self._t_cls
        cl = self.classifier
        d.sas, d.sa = _ffi.mlp_state(_ffi.mlp_tensors(cl.sas_classifier)), _ffi.mlp_state(_ffi.mlp_tensors(cl.sa_classifier))
        d.sas_m, d.sas_v = _ffi.mlp_state(self._adam["sas"][0]), _ffi.mlp_state(self._adam["sas"][1])
        d.sa_m, d.sa_v = _ffi.mlp_state(self._adam["sa"][0]), _ffi.mlp_state(self._adam["sa"][1])
        d.t, d.lr, d.nsplit = self._t_cls, float(cfg["actor_lr"]), nsplit           # Adam(lr=actor_lr), mobody.py:135
        d.workspace, d.workspace_bytes, d.scalars_out = _ffi.ptr(self._cls_ws), self._cls_ws.numel(), _ffi.ptr(self._cls_scalars)
        _ffi.check(lib.mobody_classifier_step(C.byref(d), _ffi.stream_ptr(self.device)))
        return self._cls_scalars

    def update_classifier(self, src_replay_buffer, tar_replay_buffer, batch_size, writer=None, *, _inject=None):
        """Reference signature (mobody.py:146-181) -> (loss_sa, loss_sas) as 0-d device tensors.
        ``_inject`` optionally scripts the draws of np.random.randint / torch.randperm / torch.randn_like:
        {src, tar, perm, noise_sas, noise_sa}."""
        if self.config.get("penalize_fake", 0):
            raise NotImplementedError("penalize_fake=1 mixes batch sizes the reference's own labels do not cover (mobody.py:149-162)")
        inj = _inject or {}
        B = int(batch_size)
        RW = src_replay_buffer.RW
        rows = torch.empty(2 * B, RW, dtype=torch.float32, device=self.device)
        src_replay_buffer.sample_rows(B, inj.get("src"), out=rows[:B])
        tar_replay_buffer.sample_rows(B, inj.get("tar"), out=rows[B:])
        perm = inj.get("perm")
        perm = torch.randperm(2 * B, device=self.device) if perm is None else torch.as_tensor(np.asarray(perm)).to(self.device)
        rows = rows[perm]                                                            # :164-166
        label = (perm >= B).to(torch.int32).contiguous()                             # zeros for src, ones for tar (:162)
        sc = self.classifier_step_on_rows(rows, label, noise_sas=inj.get("noise_sas"), noise_sa=inj.get("noise_sa"))
        if writer is not None and self.total_it % 5000 == 0:                         # :176-179
            v = sc.cpu().tolist()
            writer.add_scalar("train/sas classifier loss", v[1], global_step=self.total_it)
            writer.add_scalar("train/sa classifier loss", v[0], global_step=self.total_it)
        return sc[0], sc[1]

    @torch.no_grad()
    def dara_relabel(self, src_replay_buffer, penalty_out=None):
        """One-off relabel of the source rewards with the trained classifier (mobody.py:364-378), in place on the
        device-resident buffer rows: reward += penalty_coef * clamp(log-ratio, -10, 10)."""
        cl = self.classifier
        sas, k0 = _ffi.mlp_params(cl.sas_classifier)
        sa, k1 = _ffi.mlp_params(cl.sa_classifier)
        buf = src_replay_buffer
        _ffi.check(_ffi.lib().mobody_dara_relabel(_ffi.ptr(buf._rows), buf.size, buf.S, buf.A, buf.RW, C.byref(sas), C.byref(sa),
                                                  float(self.config["penalty_coef"]), _ffi.ptr(penalty_out),
                                                  _ffi.stream_ptr(self.device)))
        del k0, k1

    # ------------------------------------------------------------------ train step
    def train_on_rows(self, rows, n_true, *, nsplit=None):
        """One fused critic + Polyak + actor update on packed batch rows [N, RW] (device, rows ordered
        src, tar, fake; the first ``n_true`` rows are the src+tar rows of the BC term).  Asynchronous:
        losses land in ``self._scalars`` (device).  mobody.py:541-573.
        ``nsplit`` overrides the row-split count of the weight-gradient GEMMs (1..64; tests)."""
        cfg = self.config
        if cfg.get("advantage", 0) or not cfg.get("scale_Q", 1) or not cfg.get("q_weighted", 1):
            raise NotImplementedError("mobody_b200 implements the default advantage=0, scale_Q=1, q_weighted=1 update")
        S, A = cfg["state_dim"], cfg["action_dim"]
        N = rows.shape[0]
        if nsplit is None:
            # row splits of the weight-gradient GEMMs (partials summed in Adam, fixed order); 128 x 64 tiles per split:
            # critic 2 x (256 x 256 -> 8, 256 x (S+A) -> 2 per 64 columns), actor 8 + 2 per 64 columns of S
            nsplit = _wgrad_splits(N, (2 * (8 + 2 * ((S + A + 63) // 64)), 8 + 2 * ((S + 63) // 64)), self.device)
        lib = _ffi.lib()
        need = int(lib.mobody_train_workspace_bytes(N, S, A, nsplit))
        if self._train_ws is None or self._train_ws.numel() < need:
            self._train_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._t_q += 1; self._t_pi += 1
        d = _ffi.TrainDesc()
        d.rows, d.N, d.n_true, d.S, d.A, d.row_width = _ffi.ptr(rows), N, int(n_true), S, A, rows.shape[1]
        qt = self.target_q_funcs
        d.policy = _ffi.mlp_state(_ffi.mlp_tensors(self.policy.network))
        d.q1, d.q2 = _ffi.mlp_state(_ffi.mlp_tensors(self.q_funcs.network1)), _ffi.mlp_state(_ffi.mlp_tensors(self.q_funcs.network2))
        d.q1_target, d.q2_target = _ffi.mlp_state(_ffi.mlp_tensors(qt.network1)), _ffi.mlp_state(_ffi.mlp_tensors(qt.network2))
        d.policy_m, d.policy_v = _ffi.mlp_state(self._adam["pi"][0]), _ffi.mlp_state(self._adam["pi"][1])
        d.q1_m, d.q1_v = _ffi.mlp_state(self._adam["q1"][0]), _ffi.mlp_state(self._adam["q1"][1])
        d.q2_m, d.q2_v = _ffi.mlp_state(self._adam["q2"][0]), _ffi.mlp_state(self._adam["q2"][1])
        d.t_q, d.t_pi = self._t_q, self._t_pi
        d.gamma, d.tau = float(self.discount), float(self.tau)
        d.critic_lr, d.actor_lr = float(cfg["critic_lr"]), float(cfg["actor_lr"])
        d.weight, d.bc_coef, d.max_action = float(cfg["weight"]), float(cfg.get("bc_coef", 1.0)), float(cfg["max_action"])
        d.nsplit, d.workspace, d.workspace_bytes = nsplit, _ffi.ptr(self._train_ws), self._train_ws.numel()
        d.scalars_out = _ffi.ptr(self._scalars)
        _ffi.check(lib.mobody_train_step(C.byref(d), _ffi.stream_ptr(self.device)))
        # parameters were updated in place behind autograd's version counters: bump the epoch that the
        # tensor-core weight images (dynamics._packed_policy) are keyed on
        self.policy.network._b200_epoch = getattr(self.policy.network, "_b200_epoch", 0) + 1
        return self._scalars

    def loss_scalars(self):
        """Host copy of the last step's diagnostics (one D2H sync; call sparingly)."""
        v = self._scalars.cpu().tolist()
        return dict(q_loss=v[0], q1_mean=v[1], pi_loss=v[2], bc_loss=v[3], q_policy=v[4], q_abs_mean=v[5],
                    w_mean=v[6], w_min=v[7], w_max=v[8], p_w=v[9])

    def train(self, src_replay_buffer, tar_replay_buffer, batch_size=128, writer=None, wandbrun=None, *, _inject=None):
        """Reference signature (mobody.py:347-578).  ``_inject`` optionally supplies the buffer indices the
        reference would draw with np.random.randint, as a dict {src, tar, fake, src_init, tar_init}."""
        cfg = self.config
        self.total_it += 1
        self.src_replay_buffer, self.tar_replay_buffer = src_replay_buffer, tar_replay_buffer
        if self.penalty_type == "dara" and self.total_it == 1:                                        # :354-378
            for it in range(10 * 500):
                loss_sa, loss_sas = self.update_classifier(src_replay_buffer, tar_replay_buffer, batch_size, writer)
                if it % 2000 == 0:
                    print(loss_sa, loss_sas)
            self.dara_relabel(src_replay_buffer)
        inj = _inject or {}
        S, A = cfg["state_dim"], cfg["action_dim"]
        n_src, n_tar = int(cfg["src_ratio"] * batch_size), int(cfg["trg_ratio"] * batch_size)
        n_fake = int(cfg["fake_batch_scale"] * batch_size) if cfg["fake_batch_scale"] != 0 else 0
        RW = src_replay_buffer.RW
        rows = torch.empty(n_src + n_tar + n_fake, RW, dtype=torch.float32, device=self.device)
        refresh = (self.total_it - 1) % 5000 == 0
        fused = (not inj and not refresh and self.penalty_type != "par" and n_fake and
                 all(b.index_source == "philox" and b.size > 0 for b in (src_replay_buffer, tar_replay_buffer, self.fake_replay_buffer)))
        if fused:   # the three buffer samples of this step (:399, 400, 524) as ONE launch: Philox draw + 128-bit row gather
            jobs = (_ffi.SampleJob * 3)()
            for jb, (b, n, lo) in zip(jobs, ((src_replay_buffer, n_src, 0), (tar_replay_buffer, n_tar, n_src),
                                             (self.fake_replay_buffer, n_fake, n_src + n_tar))):
                jb.rows, jb.n, jb.size, jb.draw, jb.seed = _ffi.ptr(b._rows), n, b.size, b._draw, b.seed
                jb.out = rows.data_ptr() + 4 * RW * lo
                b._draw += 1
            _ffi.check(_ffi.lib().mobody_sample_rows(jobs, 3, RW, _ffi.stream_ptr(self.device)))
            self.train_on_rows(rows, n_src + n_tar)
            self._train_side_effects(writer, wandbrun)
            return
        src_replay_buffer.sample_rows(n_src, inj.get("src"), out=rows[:n_src])                        # :399
        tar_replay_buffer.sample_rows(n_tar, inj.get("tar"), out=rows[n_src:n_src + n_tar])           # :400
        if self.penalty_type == "par":                                                                # :428-434
            s, a_, ns = rows[:n_src, :S].contiguous(), rows[:n_src, S:S + A].contiguous(), rows[:n_src, S + A:2 * S + A]
            pred, _, _, _ = self.dynamics.step(s, a_)
            rows[:n_src, 2 * S + A] -= cfg["penalty_coef"] * ((ns - pred) ** 2).mean(1)
        if (self.total_it - 1) % 5000 == 0:                                                           # :441-475 refresh
            self.refresh_fake_buffer(src_replay_buffer, tar_replay_buffer, inj, batch_size)
        if n_fake:
            self.fake_replay_buffer.sample_rows(n_fake, inj.get("fake"), out=rows[n_src + n_tar:])    # :524
        self.train_on_rows(rows, n_src + n_tar)
        self._train_side_effects(writer, wandbrun)

    def _train_side_effects(self, writer, wandbrun):
        cfg = self.config
        if self.total_it % 1000 == 0 and cfg.get("q_weighted", 1):                                    # :269-270
            v = self.loss_scalars()
            print(v["w_mean"], v["w_min"], v["w_max"])
        if writer is not None and self.total_it % 5000 == 0:                                          # :203-205, 272-274, 332-344
            v = self.loss_scalars()
            writer.add_scalar("train/q1", v["q1_mean"], self.total_it)
            writer.add_scalar("train/exp_adv", v["w_mean"], self.total_it)
            writer.add_scalar("train/bc_loss", v["bc_loss"], self.total_it)
            writer.add_scalar("train/q_policy", v["q_policy"], self.total_it)
            writer.add_scalar("train/policy_loss", v["pi_loss"], self.total_it)
            if wandbrun is not None:
                wandbrun.log({"train/q1": v["q1_mean"], "train/q_policy": v["q_policy"], "train/policy_loss": v["pi_loss"]},
                             step=self.total_it)

    def refresh_fake_buffer(self, src_buf, tar_buf, inj=None, batch_size=128):
        """Synthetic-data refresh (mobody.py:441-475): two policy rollouts through the learned target dynamics and one
        dynamics step on dataset (s, a) pairs, all inserted into the fake buffer without leaving the device."""
        cfg, inj = self.config, inj or {}
        S, A = cfg["state_dim"], cfg["action_dim"]
        src_rows = src_buf.sample_rows(50000, inj.get("src_init"))                                    # :442 (sizes hard-coded there)
        tar_rows = tar_buf.sample_rows(2000, inj.get("tar_init"))                                     # :443
        for init, T in ((src_rows[:, :S].contiguous(), cfg["src_rollout_length"]), (tar_rows[:, :S].contiguous(), cfg["trg_rollout_length"])):
            out, info = self.rollout_device(init, T)                                                  # :444, 453
            if cfg.get("filter_bad_rollout", 1) and out is not None:
                print("filtered rollout", info["kept"], info["num_transitions"])                     # :653
            self.fake_replay_buffer.add_batch(out)
        if cfg.get("use_src_sa_to_get_target_next_state", 1):                                         # :460-475
            s, a_ = src_rows[:, :S].contiguous(), src_rows[:, S:S + A].contiguous()
            ws = StepWorkspace(s.shape[0], S, A, self.device)
            self.dynamics.launch_step(s, a_, ws)
            keep = (ws.penalty < cfg["env_filter"]).squeeze(1)                                        # strict < here (quirk 5)
            self.fake_replay_buffer.add_batch({"obss": s[keep], "next_obss": ws.next_obs[keep], "actions": a_[keep],
                                               "rewards": ws.reward[keep], "terminals": ws.terminal[keep].float()[:, None]})
        if cfg.get("rollout_from_src", 0):                                                            # :477-510
            if self.penalty_type != "dara":
                self.update_classifier(src_buf, tar_buf, batch_size)
            starts = torch.cat([src_buf.sample_rows(50000, inj.get("src_init2"))[:, :S],
                                tar_buf.sample_rows(100, inj.get("tar_init2"))[:, :S]], 0).contiguous()
            out, info = self.rollout_device(starts, cfg["rollout_from_src_length"], use_trg=False)   # source-dynamics heads
            if out is not None:
                if cfg.get("filter_bad_rollout", 1):
                    print("filtered rollout", info["kept"], info["num_transitions"])
                rows = self.fake_replay_buffer._pack(out["obss"], out["actions"], out["next_obss"], out["rewards"],
                                                     out["terminals"], True)
                cl = self.classifier                                                                  # :498-505 reward penalty
                sas, k0 = _ffi.mlp_params(cl.sas_classifier)
                sa, k1 = _ffi.mlp_params(cl.sa_classifier)
                _ffi.check(_ffi.lib().mobody_dara_relabel(_ffi.ptr(rows), rows.shape[0], S, A, rows.shape[1], C.byref(sas),
                                                          C.byref(sa), float(cfg["penalty_coef"]), None, _ffi.stream_ptr(self.device)))
                self.fake_replay_buffer.add_packed(rows, rows.shape[0])

    # ------------------------------------------------------------------ checkpoints
    def save(self, filename):                                    # mobody.py:584-588 (optimizer files: see train step)
        torch.save(self.q_funcs.state_dict(), filename + "_critic")
        torch.save(self.policy.state_dict(), filename + "_actor")

    def load(self, filename):                                    # mobody.py:590-594
        self.q_funcs.load_state_dict(torch.load(filename + "_critic"))
        self.policy.load_state_dict(torch.load(filename + "_actor"))
