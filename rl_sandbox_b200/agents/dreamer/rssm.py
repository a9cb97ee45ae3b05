"""Flat RSSM: state container and module (reference: rl_sandbox/agents/dreamer/rssm.py:12-209).

The module owns the fp32 parameters under the reference's names (checkpoint compatible).  Its
torch ``forward`` is used by the world-model observe loop; the imagination step
(``predict_next`` repeated H times with the actor in the loop) runs in librlsb (K1).
"""
import typing as t
from dataclasses import dataclass

import torch
from torch import nn

from rl_sandbox_b200.agents.dreamer.common import Dist, GRUCell, View
from rl_sandbox_b200.utils.schedulers import LinearScheduler


@dataclass
class State:
    """determ (seq, batch, D); stoch_logits (seq, batch, 32, 32) — softmax over the LAST axis;
    stoch_ (seq, batch, 1024) is sampled lazily (straight-through) on first use (rssm.py:34-37)."""
    determ: torch.Tensor
    stoch_logits: torch.Tensor
    stoch_: t.Optional[torch.Tensor] = None

    def _map(self, fn):
        return State(fn(self.determ), fn(self.stoch_logits), None if self.stoch_ is None else fn(self.stoch_))

    def flatten(self):
        return self._map(lambda x: x.flatten(0, 1).unsqueeze(0))

    def detach(self):
        return self._map(lambda x: x.detach())

    @property
    def stoch(self):
        if self.stoch_ is None:
            sample = Dist(self.stoch_logits).rsample()
            self.stoch_ = sample.reshape(self.stoch_logits.shape[:2] + (-1,))
        return self.stoch_

    @property
    def combined(self):
        return torch.cat([self.determ, self.stoch], dim=-1)

    @property
    def stoch_dist(self):
        return Dist(self.stoch_logits)

    @classmethod
    def stack(cls, states: list['State'], dim=0):
        stochs = torch.cat([s.stoch for s in states], dim=dim) if states[0].stoch_ is not None else None
        return State(torch.cat([s.determ for s in states], dim=dim),
                     torch.cat([s.stoch_logits for s in states], dim=dim), stochs)


class Quantize(nn.Module):
    """Parameter/buffer holder of the reference's (dead) codebook (rssm.py:54-105): every shipped
    config has discrete_rssm=false, so only the state-dict entries are kept for checkpoints."""

    def __init__(self, dim, n_embed):
        super().__init__()
        self.dim, self.n_embed = dim, n_embed
        embed = torch.randn(dim, n_embed)
        self.inp_in = nn.Linear(1024, n_embed * dim)
        self.inp_out = nn.Linear(n_embed * dim, 1024)
        self.register_buffer("embed", embed)
        self.register_buffer("cluster_size", torch.zeros(n_embed))
        self.register_buffer("embed_avg", embed.clone())

    def forward(self, inp):
        raise NotImplementedError("discrete_rssm is not part of the B200 hot path (unused by every shipped config)")


class RSSM(nn.Module):
    def __init__(self, latent_dim, hidden_size, actions_num, latent_classes, discrete_rssm,
                 norm_layer: t.Type[nn.Module]):
        super().__init__()
        self.latent_dim, self.latent_classes = latent_dim, latent_classes
        self.hidden_size = hidden_size
        self.ensemble_num = 1
        self.discrete_rssm = discrete_rssm
        stoch = latent_dim * latent_classes
        img_sz = 4 * 384  # embedding width of the conv encoder (reference hard-codes it, rssm.py:156)

        def two_layer(n_in):
            return nn.Sequential(nn.Linear(n_in, hidden_size), norm_layer(hidden_size), nn.ELU(inplace=True),
                                 nn.Linear(hidden_size, stoch), View((1, -1, latent_dim, latent_classes)))

        self.pre_determ_recurrent = nn.Sequential(nn.Linear(stoch + actions_num, hidden_size),
                                                  norm_layer(hidden_size), nn.ELU(inplace=True))
        self.determ_recurrent = GRUCell(input_size=hidden_size, hidden_size=hidden_size, norm=True)
        self.ensemble_prior_estimator = two_layer(hidden_size)
        self.stoch_net = two_layer(hidden_size + img_sz)
        self.determ_discretizer = Quantize(16, 16)
        self.discretizer_scheduler = LinearScheduler(1.0, 0.0, 1_000_000)
        self.determ_layer_norm = nn.LayerNorm(hidden_size)

    def estimate_stochastic_latent(self, prev_determ):
        return self.ensemble_prior_estimator(prev_determ)

    def on_train_step(self):
        pass

    def predict_next(self, prev_state: State, action):
        if self.discrete_rssm:
            raise NotImplementedError("discrete_rssm")
        x = self.pre_determ_recurrent(torch.cat([prev_state.stoch, action], dim=-1))
        x, determ = self.determ_recurrent(x, prev_state.determ)
        return State(determ, self.estimate_stochastic_latent(x)), 0

    def update_current(self, prior: State, embed) -> State:
        return State(prior.determ, self.stoch_net(torch.cat([prior.determ, embed], dim=-1)))

    def forward(self, h_prev: State, embed, action):
        prior, diff = self.predict_next(h_prev, action)
        return prior, self.update_current(prior, embed), diff
