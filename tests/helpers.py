"""Shared builders for the GPU parity tests (test infrastructure; may import oracle/)."""
import numpy as np
import torch

from oracle import mobody_oracle as M
from oracle.make_golden import HEALTHY, healthy  # noqa: F401  (pure helpers; nothing here touches /root/reference)


def cuda_module(S, A, seed, env=None, h0=None, t3_gain=1.0, device="cuda"):
    """mobody_b200.MOBODYModule on the GPU carrying the oracle's recipe weights."""
    import mobody_b200 as mb
    hs = healthy(env, S, h0) if env is not None else None
    p = M.make_dynamics_params(S, A, seed, healthy_state=hs, t3_gain=t3_gain)
    m = mb.MOBODYModule(S, A, 256, 7, 5, device=device, config={"mopo": 0, "latent_reward": 0})
    with torch.no_grad():
        for name in M.dynamics_layer_shapes(S, A):
            getattr(m, name).weight.copy_(p[name + ".weight"])
            getattr(m, name).bias.copy_(p[name + ".bias"])
    return m, p


def cuda_dynamics(S, A, seed, env, coef, precision="fp32", **kw):
    import mobody_b200 as mb
    m, p = cuda_module(S, A, seed, env, **kw)
    task = {"walker2d": "walker2d-medium-v2", "hopper": "hopper-medium-v2", "halfcheetah": "halfcheetah-medium-v2",
            "ant": "ant-medium-v2"}[env]
    dyn = mb.MOBODYEnsembleDynamics({"encoder_loss_coef": 1, "domain_loss_coef": 0, "cycle_loss_coef": 0}, m, None, None,
                                    mb.get_termination_fn(task), penalty_coef=coef, precision=precision)
    return dyn, p


AGENT_CFG = dict(max_action=1.0, hidden_sizes=256, gamma=0.99, tau=0.005, update_interval=2, actor_lr=3e-4,
                 critic_lr=3e-4, gaussian_noise_std=1.0, weight=2.5, penalty_type="none", penalty_coef=0.1, mopo=0,
                 latent_reward=0, advantage=0, q_weighted=1, scale_Q=1, bc_coef=1.0, fake_batch_scale=0.5, src_ratio=1,
                 trg_ratio=1, filter_bad_rollout=1, env_filter=10.0, src_rollout_length=1, trg_rollout_length=1,
                 use_src_sa_to_get_target_next_state=1, rollout_from_src=0)


def cuda_agent(S, A, seed, **overrides):
    import mobody_b200 as mb
    cfg = dict(AGENT_CFG, state_dim=S, action_dim=A); cfg.update(overrides)
    ag = mb.MOBODY(cfg, torch.device("cuda"))
    st = M.AgentState(S, A, seed)
    ag.policy.load_state_dict(st.policy)
    ag.q_funcs.load_state_dict(st.q)
    ag.target_q_funcs.load_state_dict(st.q_target)
    return ag, st


def rel_err(a, b):
    """max |a-b| / (|b| + scale) with scale = mean|b| of the tensor: the "relative" error of north_star,
    with an absolute floor tied to the tensor's own magnitude so that elements that happen to be ~0
    (e.g. joint angles around zero next to a torso height of 1.25) are judged at the tensor's scale."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if b.size == 0:
        return 0.0
    scale = float(np.mean(np.abs(b))) + 1e-12
    return float(np.max(np.abs(a - b) / (np.abs(b) + scale)))


def rel_err_strict(a, b):
    """max |a-b| / max(|b|, mean|b|): the TRUE relative error for every element at or above the tensor's mean magnitude, the
    tensor's scale as the floor below it (rel_err's |b| + mean|b| denominator is up to 2x more lenient).  The fp32-parity
    modes (fp32, bf16x2) are held to north_star's 1e-4 under this metric as well."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if b.size == 0:
        return 0.0
    scale = float(np.mean(np.abs(b))) + 1e-12
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), scale)))
