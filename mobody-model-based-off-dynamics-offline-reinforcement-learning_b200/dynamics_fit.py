"""Model fitting half of MOBODYEnsembleDynamics.

Mirror of algo/dynamics/mobody_dynamics.py: ``learn`` (:594-653), ``validate`` (:1114-1150), ``select_elites``
(:1152-1156) and the default path of ``train`` (:731-978; train_together = 0, train_with_src_threshold = 1,
inverse_sep_reward_loss = 0, no_vae = 0, latent_reward = 0).  One mini-batch of ``learn`` -- the three losses, the backward
pass and the Adam step -- is ONE C-ABI call (mobody_dynfit_step, csrc/dynfit.cu); an epoch keeps its data, its loss
scalars and its optimiser state on the device and reads the scalars back once.
"""
import ctypes as C

import numpy as np
import torch

from . import _ffi

FIT_SHARED = ("zs1", "zs2", "zs3", "transition1", "transition2", "transition3", "reward_model1", "reward_model2", "reward_model3")


class DynamicsFitting(object):
    """Mixed into MOBODYEnsembleDynamics."""

    FIT_NSPLIT = 8            # K splits of the weight-gradient GEMMs (measured at 7 x 256 rows: 1 -> 1 215, 2 -> 1 541, 4 -> 1 754, 8 -> 1 830, 16 -> 1 780 steps/s)

    # ------------------------------------------------------------------ optimiser state
    def _fit_state(self):
        """Adam moments of every EnsembleLinear (created lazily, zero) + per-layer step counts.  torch.optim.Adam creates
        state only for parameters that received a gradient; the counts reproduce that (the two action encoders step
        separately, the action decoders never)."""
        if getattr(self, "_fit", None) is None:
            m, v = {}, {}
            for n in _ffi.DYN_LAYER_NAMES:
                lay = getattr(self.model, n)
                m[n] = (torch.zeros_like(lay.weight), torch.zeros_like(lay.bias))
                v[n] = (torch.zeros_like(lay.weight), torch.zeros_like(lay.bias))
            self._fit = {"m": m, "v": v, "t": {n: 0 for n in _ffi.DYN_LAYER_NAMES}, "ws": None, "draw": 0}
        return self._fit

    def _fit_lr(self):
        if self.optim is not None and getattr(self.optim, "param_groups", None):
            return float(self.optim.param_groups[0]["lr"])                      # train_mobody.py:801-804
        return float((self.config or {}).get("dynamics_lr", 1e-3))

    def _dyn_state(self, pairs):
        st, keep = _ffi.DynState(), []
        for i, n in enumerate(_ffi.DYN_LAYER_NAMES):
            w, b = pairs[n]
            w, b = w.detach(), b.detach()
            if w.dtype != torch.float32 or not w.is_cuda or not w.is_contiguous() or not b.is_contiguous():
                raise RuntimeError(f"mobody_b200: {n} tensors must be contiguous fp32 CUDA tensors")
            st.w[i], st.b[i] = w.data_ptr(), b.data_ptr()
            keep += [w, b]
        return st, keep

    # ------------------------------------------------------------------ one mini-batch on the device
    def fit_batch(self, use_trg, obs, act, next_obs, reward, *, rows=None, lo=0, eps_latent=None, eps_next=None, scalars_out=None):
        """One optimiser step of learn() on the batch ``[:, lo:lo+rows]`` of the device tensors obs / next_obs [7,N,S],
        act [7,N,A], reward [7,N,1] (contiguous fp32).  Enqueues on the current stream, no host synchronisation; the loss
        scalars [loss, transition, encoder, recon, kl, reward] land in ``scalars_out`` (device float[8])."""
        dev = self.model.elites.device
        S, A = self.model.obs_dim, self.model.action_dim
        E, N = obs.shape[0], obs.shape[1]
        B = N - lo if rows is None else int(rows)
        for t, w in ((obs, S), (act, A), (next_obs, S), (reward, 1)):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != (7, N, w):
                raise RuntimeError("mobody_b200: fit_batch wants contiguous fp32 CUDA tensors [7, N, .]")
        if E != 7 or B < 1 or lo < 0 or lo + B > N:
            raise ValueError("fit_batch: 7 members, at least 1 row per member, window inside the data")
        st = self._fit_state()
        trained = FIT_SHARED + (("za_trg1", "za_trg2") if use_trg else ("za_src1", "za_src2"))
        for n in trained:
            st["t"][n] += 1
        ns = self.FIT_NSPLIT
        need = int(_ffi.lib().mobody_dynfit_workspace_bytes(B, S, A, ns))
        if st["ws"] is None or st["ws"].numel() < need:
            st["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
        if scalars_out is None:
            scalars_out = torch.zeros(8, dtype=torch.float32, device=dev)
        d = _ffi.DynFitDesc()
        d.S, d.A, d.B, d.use_trg = S, A, B, int(bool(use_trg))
        d.obs, d.act = obs.data_ptr() + 4 * lo * S, act.data_ptr() + 4 * lo * A
        d.next_obs, d.reward = next_obs.data_ptr() + 4 * lo * S, reward.data_ptr() + 4 * lo
        d.member_stride = N
        if (eps_latent is None) != (eps_next is None):
            raise ValueError("fit_batch: inject both noise tensors or neither")
        if eps_latent is not None:
            eps_latent, eps_next = _ffi.f32(eps_latent, dev), _ffi.f32(eps_next, dev)
            assert tuple(eps_latent.shape) == (6, 7, B, 16) and tuple(eps_next.shape) == (7, B, S)
            d.eps_latent, d.eps_next = eps_latent.data_ptr(), eps_next.data_ptr()
        d.seed, d.draw = self.seed, st["draw"]
        st["draw"] += 1
        coef = float(self.encoder_loss_coef)
        d.encoder_coef = (5.0 if use_trg else 1.0) * coef                         # :623-626
        d.reward_coef = 1.0 if use_trg else 0.01                                  # :381-384
        params = {n: (getattr(self.model, n).weight, getattr(self.model, n).bias) for n in _ffi.DYN_LAYER_NAMES}
        d.params, k0 = self._dyn_state(params)
        d.adam_m, k1 = self._dyn_state(st["m"])
        d.adam_v, k2 = self._dyn_state(st["v"])
        d.t_shared, d.t_action = st["t"]["zs1"], st["t"]["za_trg1" if use_trg else "za_src1"]
        d.lr, d.nsplit = self._fit_lr(), ns
        d.workspace, d.workspace_bytes = st["ws"].data_ptr(), st["ws"].numel()
        d.scalars_out = scalars_out.data_ptr()
        _ffi.check(_ffi.lib().mobody_dynfit_step(C.byref(d), _ffi.stream_ptr(dev)))
        del k0, k1, k2
        return scalars_out

    # ------------------------------------------------------------------ reference API
    def learn(self, use_trg_data, train_obss, train_actions, train_next_obss, train_rewards, batch_size, logvar_loss_coef,
              trg_transition=None):
        """mobody_dynamics.py:594-653 for the default configuration: mini-batches over the second axis of the
        bootstrapped [7, N, .] tensors -> (mean loss, mean transition loss, mean encoder loss, mean recon loss, mean kl
        loss) as Python floats.  The data goes to the device once per call, the scalars come back once per call."""
        cfg = self.config or {}
        if cfg.get("no_vae") or cfg.get("latent_reward") or cfg.get("inverse_sep_reward_loss") or cfg.get("mopo"):
            raise NotImplementedError("mobody_b200 fits the default configuration (no_vae = latent_reward = inverse_sep_reward_loss = mopo = 0)")
        self.model.train()
        dev = self.model.elites.device
        obs, act, nobs, rew = (_ffi.f32(x, dev) for x in (train_obss, train_actions, train_next_obss, train_rewards))
        N = obs.shape[1]
        rew = rew.reshape(7, N, 1)
        n_batch = int(np.ceil(N / batch_size))
        out = torch.zeros(n_batch, 8, dtype=torch.float32, device=dev)
        for b in range(n_batch):
            self.total_steps = getattr(self, "total_steps", 0) + 1
            lo = b * batch_size
            self.fit_batch(bool(use_trg_data), obs, act, nobs, rew, rows=min(batch_size, N - lo), lo=lo, scalars_out=out[b])
        m = out.mean(0).cpu()
        return float(m[0]), float(m[1]), float(m[2]), float(m[3]), float(m[4])

    @torch.no_grad()
    def validate(self, use_trg_data, holdout_obss, holdout_actions, holdout_next_obss, holdout_rewards):
        """mobody_dynamics.py:1114-1150 -> (per-member transition losses, per-member reward losses) as lists of 7 floats.
        Runs once per epoch on <= 1000 held-out rows: plain torch ops on the device through the module's eval-mode forward."""
        self.model.eval(); self.model.inference()
        dev = self.model.elites.device
        obs, act, nobs, rew = (_ffi.f32(x, dev) for x in (holdout_obss, holdout_actions, holdout_next_obss, holdout_rewards))
        fwd = self.model.forward_trg if use_trg_data else self.model.forward_src
        mean, _, _ = fwd(obs, act)
        mean = self.obs_scaler.inverse_transform(mean)
        transition_loss = ((mean - nobs) ** 2).mean(dim=(1, 2))
        rep = lambda x: x.unsqueeze(0).repeat(7, 1, 1) if x.dim() == 2 else x          # noqa: E731
        pred_reward, _ = self.model.encode_reward(rep(obs), rep(act), mean)
        encode_loss = ((pred_reward - rew) ** 2).mean(dim=(1, 2))
        state_hat = self.model.encoder_decoder(obs)[0].mean(0)
        vae_recon_eval_loss = torch.sqrt(((state_hat - obs) ** 2).sum(dim=-1)).mean(dim=0)
        both = torch.cat([transition_loss, encode_loss, vae_recon_eval_loss.reshape(1)]).cpu().numpy()
        print("vae_recon_eval_loss", float(both[14]))
        self.model.uninference()
        return list(both[:7]), list(both[7:14])

    def select_elites(self, metrics):
        """mobody_dynamics.py:1152-1156."""
        pairs = sorted(zip(metrics, range(len(metrics))), key=lambda x: x[0])
        return [pairs[i][1] for i in range(self.model.num_elites)]

    def shuffle_rows(self, arr):
        """mobody_dynamics.py:656-658 (on whatever device ``arr`` lives)."""
        idxes = torch.argsort(torch.rand(arr.shape, device=arr.device), dim=-1)
        return torch.gather(arr, 1, idxes)

    def train(self, src_data, trg_data, max_epochs=None, max_epochs_since_update=5, batch_size=256, holdout_ratio=0.2,
              logvar_loss_coef=0.01, writer=None, buffer=None):
        """Default path of mobody_dynamics.py:731-978: per epoch one pass of learn() over the bootstrapped source data,
        three over the target data, validation on the held-out rows, ``update_save`` of the members that improved by more
        than 1 %, early stop after ``max_epochs_since_update`` epochs without improvement; then elites + ``load_save``.
        src_data / trg_data = (obs, action, next_obs, reward, ...) as returned by ReplayBuffer.sample_all."""
        cfg = self.config or {}
        if cfg.get("train_together") or cfg.get("train_with_src_threshold", 1) != 1:
            raise NotImplementedError("mobody_b200 mirrors the default train() path (train_together = 0, train_with_src_threshold = 1)")
        dev = self.model.elites.device
        self.total_steps = 0
        E = self.model.num_ensemble
        src = [_ffi.f32(x, dev) for x in src_data[:4]]
        trg = [_ffi.f32(x, dev) for x in trg_data[:4]]
        n_src, n_trg = src[0].shape[0], trg[0].shape[0]
        src_hold, trg_hold = min(int(n_src * holdout_ratio), 1000), min(int(n_trg * holdout_ratio), 500)

        def split(data, n, hold):
            perm = torch.randperm(n, device=dev)                                 # torch.utils.data.random_split (:773-776)
            tr, ho = perm[:n - hold], perm[n - hold:]
            return [x[tr] for x in data], [x[ho] for x in data]
        src_tr, src_ho = split(src, n_src, src_hold)
        trg_tr, trg_ho = split(trg, n_trg, trg_hold)
        n_src_tr, n_trg_tr = n_src - src_hold, n_trg - trg_hold
        trg_holdout_losses = [1e10] * E
        src_idx = torch.randint(n_src_tr, (E, n_src_tr), device=dev)             # bootstrap (:829-830)
        trg_idx = torch.randint(n_trg_tr, (E, n_trg_tr), device=dev)
        epoch, cnt = 0, 0
        print("Training dynamics:")
        while True:
            epoch += 1
            self.epoch = epoch
            sb = [x[src_idx] for x in src_tr]
            _, src_transition_loss, src_encoder_loss, recon_loss, kl_loss = self.learn(False, *sb, batch_size, logvar_loss_coef)
            src_new, _ = self.validate(False, *src_ho)
            src_holdout_loss = float(np.sort(src_new)[:self.model.num_elites].mean())
            print(epoch)
            for k, val in (("src_loss/dynamics_train_loss", src_transition_loss), ("src_loss/dynamics_encoder_loss", src_encoder_loss),
                           ("src_loss/dynamics_recon_loss", recon_loss), ("src_loss/dynamics_kl_loss", kl_loss),
                           ("src_loss/dynamics_holdout_loss", src_holdout_loss)):
                print(k, val)
            if writer is not None:
                writer.add_scalar("src_loss/dynamics_train_loss", src_transition_loss, global_step=epoch)
                writer.add_scalar("src_loss/dynamics_encoder_loss", src_encoder_loss, global_step=epoch)
                writer.add_scalar("src_loss/dynamics_domain_loss", recon_loss, global_step=epoch)
                writer.add_scalar("src_loss/dynamics_holdout_loss", src_holdout_loss, global_step=epoch)
            tb = [x[trg_idx] for x in trg_tr]
            for _ in range(3):                                                    # :914-925
                _, trg_transition_loss, trg_encoder_loss, recon_loss, kl_loss = self.learn(True, *tb, batch_size, logvar_loss_coef)
            trg_new, trg_new_rew = self.validate(True, *trg_ho)
            trg_holdout_loss = float(np.sort(trg_new)[:self.model.num_elites].mean())
            trg_holdout_reward_loss = float(np.sort(trg_new_rew)[:self.model.num_elites].mean())
            for k, val in (("trg_loss/dynamics_train_loss", trg_transition_loss), ("trg_loss/dynamics_encoder_loss", trg_encoder_loss),
                           ("trg_loss/dynamics_recon_loss", recon_loss), ("trg_loss/dynamics_kl_loss", kl_loss),
                           ("trg_loss/dynamics_holdout_loss", trg_holdout_loss), ("trg_loss/dynamics_holdout_reward_loss", trg_holdout_reward_loss)):
                print(k, val)
            print(" ")
            if writer is not None:
                writer.add_scalar("trg_loss/dynamics_train_loss", trg_transition_loss, global_step=epoch)
                writer.add_scalar("trg_loss/dynamics_encoder_loss", trg_encoder_loss, global_step=epoch)
                writer.add_scalar("trg_loss/dynamics_holdout_loss", trg_holdout_loss, global_step=epoch)
            src_idx, trg_idx = self.shuffle_rows(src_idx), self.shuffle_rows(trg_idx)
            indexes = []
            for i, (new_loss, old_loss) in enumerate(zip(trg_new, trg_holdout_losses)):
                if (old_loss - new_loss) / old_loss > 0.01:                       # :952-956
                    indexes.append(i)
                    trg_holdout_losses[i] = float(new_loss)
            if indexes:
                self.model.update_save(indexes)
                cnt = 0
            else:
                cnt += 1
            if cnt >= max_epochs_since_update or (max_epochs and epoch >= max_epochs):
                print("src_loss/dynamics_holdout_loss", round(src_holdout_loss, 5))
                print("trg_loss/dynamics_holdout_loss", round(trg_holdout_loss, 5))
                break
        indexes = self.select_elites(trg_holdout_losses)
        self.model.set_elites(indexes)
        self.model.load_save()
        self.model.eval()
        print("elites:{} , holdout loss: {}".format(indexes, np.sort(trg_holdout_losses)[:self.model.num_elites].mean()))
