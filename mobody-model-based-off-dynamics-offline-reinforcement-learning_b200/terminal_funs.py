"""Termination predicates, device resident.

Mirror of algo/mb_utils/terminal_funs.py: ``get_termination_fn(task)`` keeps the reference's
substring dispatch order (:123-149) and returns a callable ``fn(obs, act, next_obs) -> done[B,1]``.
The predicate itself runs in CUDA (csrc/term.cuh) — inside the fused step kernel on the hot path,
or through ``mobody_termination`` when called stand-alone.  NumPy inputs are accepted (copied to
the device and back) so reference call sites keep working; there is no CPU implementation.
"""
import numpy as np
import torch

from . import _ffi

TERM_KINDS = {"never": 0, "halfcheetah": 1, "hopper": 2, "walker2d": 3, "ant": 4, "humanoid": 5, "pen": 6}
_DISPATCH = [("halfcheetahvel", "never"), ("halfcheetah", "halfcheetah"), ("hopper", "hopper"),
             ("antangle", "ant"), ("ant", "ant"), ("walker2d", "walker2d"), ("point2denv", "never"),
             ("point2dwallenv", "never"), ("pendulum", "never"), ("humanoid", "humanoid"), ("pen", "pen"),
             ("door", "never")]


class TerminationFn:
    def __init__(self, name, kind):
        self.__name__ = "termination_fn_" + name
        self.kind = kind

    def device_mask(self, next_obs):
        """next_obs: CUDA fp32 [B,S] -> CUDA uint8 [B] (no host sync)."""
        assert next_obs.dim() == 2
        x = next_obs.contiguous()
        out = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
        _ffi.check(_ffi.lib().mobody_termination(_ffi.ptr(x), x.shape[0], x.shape[1], self.kind, _ffi.ptr(out),
                                                 _ffi.stream_ptr(x.device)))
        return out

    def __call__(self, obs, act, next_obs):
        assert len(obs.shape) == len(next_obs.shape) == len(act.shape) == 2      # terminal_funs.py:11
        if torch.is_tensor(next_obs) and next_obs.is_cuda:
            return self.device_mask(next_obs.float()).bool()[:, None]
        x = _ffi.f32(np.ascontiguousarray(next_obs), torch.device("cuda"))
        return self.device_mask(x).bool()[:, None].cpu().numpy()


def get_termination_fn(task):
    for key, name in _DISPATCH:
        if key in task:
            return TerminationFn(name, TERM_KINDS[name])
    raise TypeError(f"no termination function for task {task!r}")   # reference: `raise np.zeros` is a TypeError
