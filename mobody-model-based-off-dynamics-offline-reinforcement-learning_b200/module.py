"""Host-side mirror of the ensemble latent dynamics network.

Same class / parameter names and shapes as the reference
(algo/dynamics/mobody_module.py:51-214, 371-416) so ``dynamics.pth`` state_dicts load unchanged
and ``train_mobody.py:791-799`` can construct it.  The nn.Module only *owns* the parameters;
the inference hot path (MOBODYEnsembleDynamics.step) hands their device pointers to the fused
CUDA kernels; the fitting step (dynamics_fit.py) updates them in place.  ``forward_trg`` /
``forward_src`` / ``encode_reward`` stay as plain torch code for callers that want autograd
through the model and for the once-per-epoch ``validate``; they are not on the rollout path.
"""
from typing import List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F

LATENT_DIM = 16   # mobody_module.py:95


class Swish(nn.Module):
    def forward(self, x):                                   # mobody_module.py:13-15
        return x * torch.sigmoid(x)


def soft_clamp(x, _min=None, _max=None):                    # mobody_module.py:18-29
    if _max is not None:
        x = _max - F.softplus(_max - x)
    if _min is not None:
        x = _min + F.softplus(x - _min)
    return x


class EnsembleLinear(nn.Module):
    """weight [E,in,out], bias [E,1,out]; 2-D input broadcasts over members (mobody_module.py:371-404)."""

    def __init__(self, input_dim, output_dim, num_ensemble, weight_decay=0.0):
        super().__init__()
        self.num_ensemble = num_ensemble
        self.weight = nn.Parameter(torch.zeros(num_ensemble, input_dim, output_dim))
        self.bias = nn.Parameter(torch.zeros(num_ensemble, 1, output_dim))
        nn.init.trunc_normal_(self.weight, std=1 / (2 * input_dim ** 0.5))
        self.saved_weight = nn.Parameter(self.weight.detach().clone())
        self.saved_bias = nn.Parameter(self.bias.detach().clone())
        self.weight_decay = weight_decay

    def forward(self, x):
        if x.dim() == 2:
            x = torch.einsum("ij,bjk->bik", x, self.weight)
        else:
            x = torch.einsum("bij,bjk->bik", x, self.weight)
        return x + self.bias

    def load_save(self):
        self.weight.data.copy_(self.saved_weight.data)
        self.bias.data.copy_(self.saved_bias.data)

    def update_save(self, indexes):
        self.saved_weight.data[indexes] = self.weight.data[indexes]
        self.saved_bias.data[indexes] = self.bias.data[indexes]

    def get_decay_loss(self):
        return self.weight_decay * (0.5 * ((self.weight ** 2).sum()))


class MOBODYModule(nn.Module):
    def __init__(self, obs_dim: int, action_dim: int, hidden_dims: Union[int, List[int], Tuple[int]] = 256,
                 num_ensemble: int = 7, num_elites: int = 5, activation=Swish,
                 weight_decays: Optional[Union[List[float], Tuple[float]]] = None, with_reward: bool = True,
                 device: str = "cuda", reward_relu: bool = False, config=None) -> None:
        super().__init__()
        config = config if config is not None else {"mopo": 0, "latent_reward": 0}
        if config.get("mopo") or config.get("latent_reward"):
            # the reference's own mopo / latent_reward branches call functions with the wrong arity
            # (mobody_dynamics.py:397) and are not reachable from train_mobody.py defaults
            raise NotImplementedError("mobody_b200 supports config['mopo']=0 and config['latent_reward']=0 only")
        if isinstance(hidden_dims, (list, tuple)):
            hidden_dims = hidden_dims[0]
        if hidden_dims != 256 or num_ensemble != 7:
            raise NotImplementedError("mobody_b200 kernels are built for hidden 256 and 7 members "
                                      "(train_mobody.py:791-799; literal 7 in mobody_dynamics.py:218)")
        self.training = True
        self.config = config
        self.num_ensemble, self.num_elites = num_ensemble, num_elites
        self._with_reward = 0
        self.device = torch.device(device)
        self.activation = activation()
        self.reward_relu = reward_relu
        self.encode_trg_diff = 0
        self.obs_dim, self.action_dim = obs_dim, action_dim
        wd, E, H, L = 5e-5, num_ensemble, hidden_dims, LATENT_DIM
        mk = lambda i, o: EnsembleLinear(i, o, E, wd)       # noqa: E731
        self.zs1, self.zs2, self.zs3 = mk(obs_dim, H), mk(H, H), mk(H, 2 * L)
        self.za_src1, self.za_src2 = mk(L + action_dim, 32), mk(32, 2 * L)
        self.za_de_src1, self.za_de_src2 = mk(L, 8), mk(8, action_dim)
        self.za_trg1, self.za_trg2 = mk(L + action_dim, 32), mk(32, 2 * L)
        self.za_de_trg1, self.za_de_trg2 = mk(L, 8), mk(8, action_dim)
        self.transition1, self.transition2, self.transition3 = mk(L, H), mk(H, H), mk(H, obs_dim)
        self.reward_model1, self.reward_model2, self.reward_model3 = mk(2 * obs_dim + action_dim, H), mk(H, H), mk(H, 2)
        self.module_list = [self.zs1, self.zs2, self.zs3, self.za_src1, self.za_src2, self.za_de_src1,
                            self.za_de_src2, self.za_trg1, self.za_trg2, self.za_de_trg1, self.za_de_trg2,
                            self.transition1, self.transition2, self.transition3,
                            self.reward_model1, self.reward_model2, self.reward_model3]
        self.max_logvar = nn.Parameter(torch.ones(obs_dim) * 0.5, requires_grad=False)
        self.min_logvar = nn.Parameter(torch.ones(obs_dim) * -10, requires_grad=False)
        self.max_logvar_latent = nn.Parameter(torch.ones(L) * 20, requires_grad=False)
        self.min_logvar_latent = nn.Parameter(torch.ones(L) * -20, requires_grad=False)
        self.elites = nn.Parameter(torch.tensor(list(range(0, num_elites))), requires_grad=False)
        self.to(self.device)

    # ---- differentiable torch forward (model fitting; not the rollout path) ----
    def reparameterize(self, mu, logvar):                    # mobody_module.py:237-243
        if self.training:
            std = torch.exp(0.5 * logvar)
            return mu + torch.randn_like(std) * std
        return mu

    def encode_state(self, state):                           # :217-225
        zs = self.activation(self.zs1(state))
        zs = self.activation(self.zs2(zs))
        mu, logvar = torch.chunk(self.zs3(zs), 2, dim=-1)
        return self.reparameterize(mu, logvar), mu, logvar

    def _encode_action(self, l1, l2, s, a):
        if s.dim() == 3 and a.dim() == 2:
            a = a.unsqueeze(0).repeat(self.num_ensemble, 1, 1)
        za = l2(self.activation(l1(torch.cat([s, a], dim=-1))))
        mu, _ = torch.chunk(za, 2, dim=-1)
        return mu

    def encode_src_action(self, s, a, reparam=True):         # :245-256
        return self._encode_action(self.za_src1, self.za_src2, s, a)

    def encode_trg_action(self, s, a, reparam=True):         # :258-271
        return self._encode_action(self.za_trg1, self.za_trg2, s, a)

    def decode_src_action(self, z):                          # :273-279
        return self.za_de_src2(self.activation(self.za_de_src1(z)))

    def decode_trg_action(self, z):                          # :280-285 (uses the *src* decoder, as the reference does)
        return self.za_de_src2(self.activation(self.za_de_src1(z)))

    def encode_transition(self, z):                          # :287-293
        z = self.activation(self.transition1(z))
        z = self.activation(self.transition2(z))
        return self.transition3(z)

    def encode_reward(self, s, a, next_s):                   # :295-302
        v = self.activation(self.reward_model1(torch.cat([s, a, next_s], dim=-1)))
        v = self.activation(self.reward_model2(v))
        mu, logvar = torch.chunk(self.reward_model3(v), 2, dim=-1)
        return mu, soft_clamp(logvar, -10, 0.5)

    def forward_src(self, state, action):                    # :315-321
        zs, zs_mu, zs_logvar = self.encode_state(state)
        return self.encode_transition(zs + self.encode_src_action(zs, action, False)), zs_mu, zs_logvar

    def forward_trg(self, state, action):                    # :323-330
        zs, zs_mu, zs_logvar = self.encode_state(state)
        return self.encode_transition(zs + self.encode_trg_action(zs, action, False)), zs_mu, zs_logvar

    def encoder_decoder(self, state):                        # :332-335
        zs, zs_mu, zs_logvar = self.encode_state(state)
        return self.encode_transition(zs), zs_mu, zs_logvar

    # ---- bookkeeping ----
    def load_save(self):
        for layer in self.module_list:
            layer.load_save()

    def update_save(self, indexes):
        for layer in self.module_list:
            layer.update_save(indexes)

    def get_decay_loss(self):
        return sum(layer.get_decay_loss() for layer in self.module_list)

    def set_elites(self, indexes):                           # :351-353
        assert len(indexes) <= self.num_ensemble and max(indexes) < self.num_ensemble
        self.register_parameter("elites", nn.Parameter(torch.tensor(indexes, device=self.elites.device), requires_grad=False))

    def random_elite_idxs(self, batch_size: int) -> np.ndarray:   # :355-357 (host NumPy RNG, kept for API parity)
        return np.random.choice(self.elites.data.cpu().numpy(), size=batch_size)

    def inference(self):
        self.training = False

    def uninference(self):
        self.training = True
