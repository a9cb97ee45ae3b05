"""Slotted RSSM: state container and module (reference: rl_sandbox/agents/dreamer/rssm_slots_attention.py:13-243).

Slots are folded into the batch for the flat cell (pre_determ_recurrent, GRU, prior MLP); after the GRU
``attention_block_num`` mixer blocks let the slots exchange information before the prior logits are
formed; the State keeps the UN-mixed ``determ`` and carries the mixed one as ``determ_updated``
(:207-208).  The module owns the fp32 parameters under the reference's names.  Its torch ``forward``
serves the world-model observe loop; the imagination step runs in librlsb (K1 with ``slots > 1``:
rlsb_imagine.cu + rlsb_mixer.cu).
"""
import typing as t
from dataclasses import dataclass

import torch
from torch import nn

from rl_sandbox_b200.agents.dreamer.common import Dist, GRUCell, View
from rl_sandbox_b200.utils.schedulers import LinearScheduler


@dataclass
class State:
    """determ (seq, batch, slots, D); stoch_logits (seq, batch, slots, 32, 32); stoch_ (seq, batch, slots, 1024)
    sampled lazily; pos_enc (1, 1, slots, D + 1024) is ADDED to cat[determ, stoch] in ``combined_slots``."""
    determ: torch.Tensor
    stoch_logits: torch.Tensor
    stoch_: t.Optional[torch.Tensor] = None
    pos_enc: t.Optional[torch.Tensor] = None
    determ_updated: t.Optional[torch.Tensor] = None

    def flatten(self):
        f = lambda x: x.flatten(0, 1).unsqueeze(0)
        return State(f(self.determ), f(self.stoch_logits), None if self.stoch_ is None else f(self.stoch_), self.pos_enc)

    def detach(self):
        return State(self.determ.detach(), self.stoch_logits.detach(),
                     None if self.stoch_ is None else self.stoch_.detach(),
                     None if self.pos_enc is None else self.pos_enc.detach())

    @property
    def combined(self):
        return self.combined_slots.flatten(2, 3)

    @property
    def combined_slots(self):
        state = torch.cat([self.determ, self.stoch], dim=-1)
        return state + self.pos_enc if self.pos_enc is not None else state

    @property
    def stoch(self):
        if self.stoch_ is None:
            self.stoch_ = Dist(self.stoch_logits).rsample().reshape(self.stoch_logits.shape[:3] + (-1,))
        return self.stoch_

    @property
    def stoch_dist(self):
        return Dist(self.stoch_logits)

    @classmethod
    def stack(cls, states: list['State'], dim=0):
        stochs = torch.cat([s.stoch for s in states], dim=dim) if states[0].stoch_ is not None else None
        return State(torch.cat([s.determ for s in states], dim=dim),
                     torch.cat([s.stoch_logits for s in states], dim=dim), stochs, states[0].pos_enc)


class RSSM(nn.Module):
    def __init__(self, latent_dim, hidden_size, actions_num, latent_classes, discrete_rssm,
                 norm_layer: t.Type[nn.Module], full_qk_from: int = 1, symmetric_qk: bool = False,
                 attention_block_num: int = 3, embed_size=2 * 2 * 384):
        super().__init__()
        self.latent_dim, self.latent_classes = latent_dim, latent_classes
        self.ensemble_num = 1
        self.hidden_size = hidden_size
        self.discrete_rssm = discrete_rssm
        self.symmetric_qk = symmetric_qk
        stoch = latent_dim * latent_classes

        def two_layer(n_in):
            return nn.Sequential(nn.Linear(n_in, hidden_size), norm_layer(hidden_size), nn.ELU(inplace=True),
                                 nn.Linear(hidden_size, stoch), View((1, -1, latent_dim, latent_classes)))

        self.pre_determ_recurrent = nn.Sequential(nn.Linear(stoch + actions_num, hidden_size),
                                                  norm_layer(hidden_size), nn.ELU(inplace=True))
        self.determ_recurrent = GRUCell(input_size=hidden_size, hidden_size=hidden_size, norm=True)
        self.ensemble_prior_estimator = two_layer(hidden_size)
        self.stoch_net = two_layer(hidden_size + embed_size)
        self.hidden_attention_proj = nn.Linear(hidden_size, 3 * hidden_size, bias=False)
        self.pre_norm = nn.LayerNorm(hidden_size)
        self.fc = nn.Linear(hidden_size, hidden_size)
        self.fc_norm = nn.LayerNorm(hidden_size)
        self.attention_scheduler = LinearScheduler(0.0, 1.0, full_qk_from)
        self.attention_block_num = attention_block_num
        self.att_scale = hidden_size ** (-0.5)
        self.eps = 1e-8
        self.last_attention = None

    def on_train_step(self):
        self.attention_scheduler.step()

    def estimate_stochastic_latent(self, prev_determ):
        return self.ensemble_prior_estimator(prev_determ)

    def mix(self, determ_post: torch.Tensor) -> torch.Tensor:
        """the slot-mixing blocks (:186-203) on (seq, batch, slots, D)"""
        attn = None
        for _ in range(self.attention_block_num):
            q, k, v = self.hidden_attention_proj(self.pre_norm(determ_post)).chunk(3, dim=-1)
            if self.symmetric_qk:
                k = q
            qk = torch.einsum('lbih,lbjh->lbij', q, k).float()
            attn = torch.softmax(self.att_scale * qk, dim=-1) + self.eps
            attn = attn / attn.sum(dim=-1, keepdim=True)
            coeff = self.attention_scheduler.val
            attn = coeff * attn + (1 - coeff) * torch.eye(q.shape[-2], device=q.device)
            updates = torch.einsum('lbjd,lbij->lbid', v, attn)
            determ_post = determ_post + self.fc(self.fc_norm(updates))
        if attn is not None:
            self.last_attention = attn.mean(dim=1).squeeze()
        return determ_post

    def predict_next(self, prev_state: State, action):
        if self.discrete_rssm:
            raise NotImplementedError("discrete rssm was not adopted for slot attention")
        slots = prev_state.determ.shape[2]
        x = self.pre_determ_recurrent(torch.cat([prev_state.stoch, action.unsqueeze(2).repeat((1, 1, slots, 1))], dim=-1))
        x, determ_prior = self.determ_recurrent(x.flatten(1, 2), prev_state.determ.flatten(1, 2))
        determ_post = self.mix(determ_prior.reshape(prev_state.determ.shape))
        logits = self.estimate_stochastic_latent(determ_post.reshape(determ_prior.shape)).reshape(
            prev_state.stoch_logits.shape)
        return State(determ_prior.reshape(prev_state.determ.shape), logits, pos_enc=prev_state.pos_enc,
                     determ_updated=determ_post), 0

    def update_current(self, prior: State, embed) -> State:
        return State(prior.determ,
                     self.stoch_net(torch.cat([prior.determ_updated, embed], dim=-1)).flatten(1, 2).reshape(
                         prior.stoch_logits.shape), pos_enc=prior.pos_enc)

    def forward(self, h_prev: State, embed, action):
        prior, diff = self.predict_next(h_prev, action)
        return prior, self.update_current(prior, embed), diff
