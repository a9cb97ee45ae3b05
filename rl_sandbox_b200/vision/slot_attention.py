"""Slot attention (reference: rl_sandbox/vision/slot_attention.py:13-101).

``SlotAttention`` owns the fp32 parameters under the reference's names.  Its forward runs in
librlsb (K3, rlsb_slot_attention_fwd) whenever no gradient is required (acting, metrics, parity);
K3 has no backward yet, so a call that needs gradients (the world-model loss) is evaluated with
torch ops on the same parameters — recorded in DESIGN.md as an open item.
"""
import typing as t

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class SlotAttention(nn.Module):
    def __init__(self, num_slots: int, n_dim: int, n_iter: int, use_prev_slots: bool):
        super().__init__()
        self.n_slots, self.n_iter, self.n_dim = num_slots, n_iter, n_dim
        self.scale = n_dim ** (-1 / 2)
        self.epsilon = 1e-8
        self.use_prev_slots = use_prev_slots
        shared = 1 if use_prev_slots else num_slots      # one shared Gaussian vs one per slot
        self.slots_mu = nn.Parameter(torch.randn(1, shared, n_dim))
        self.slots_logsigma = nn.Parameter(torch.zeros(1, shared, n_dim))
        nn.init.xavier_uniform_(self.slots_logsigma)
        self.slots_proj = nn.Linear(n_dim, n_dim, bias=False)
        self.slots_proj_2 = nn.Sequential(nn.Linear(n_dim, n_dim * 4), nn.ReLU(inplace=True),
                                          nn.Linear(n_dim * 4, n_dim))
        self.slots_norm = nn.LayerNorm(n_dim)
        self.slots_norm_2 = nn.LayerNorm(n_dim)
        self.slots_reccur = nn.GRUCell(input_size=n_dim, hidden_size=n_dim)
        self.inputs_proj = nn.Linear(n_dim, n_dim * 2, bias=False)
        self.inputs_norm = nn.LayerNorm(n_dim)
        self.prev_slots = None
        self.last_attention = None
        self._engine = None
        self._engines = None
        self._engine_key = None

    def generate_initial(self, batch: int):
        mu = self.slots_mu.expand(batch, self.n_slots, -1)
        sigma = self.slots_logsigma.exp().expand(batch, self.n_slots, -1)
        return mu + sigma * torch.randn(mu.shape, device=mu.device)

    def mark_weights_changed(self):
        self._engine_key = None

    def _kernel_forward_prepare(self, X):
        from rl_sandbox_b200 import ops
        key = (X.shape[0], X.shape[1], tuple(p._version for p in self.parameters()))
        # one engine per (frames, tokens), kept alive: CUDA graphs captured around a call replay its raw workspace pointers
        shape = (X.shape[0], X.shape[1])
        if self._engines is None:
            self._engines = {}
        if shape not in self._engines:
            self._engines[shape] = ops.SlotAttentionEngine(self.n_slots, self.n_dim, X.shape[1], self.n_iter, device=X.device)
        if self._engine is not self._engines[shape]:
            self._engine = self._engines[shape]
            self._engine_key = None
        if self._engine_key != key:
            self._engine.pack(self.state_dict())
            self._engine_key = key

    def _kernel_forward(self, X, slots):
        self._kernel_forward_prepare(X)
        return self._engine.forward(X.float(), slots.float())

    def forward(self, X: torch.Tensor, prev_slots: t.Optional[torch.Tensor]) -> torch.Tensor:
        batch = X.shape[0]
        if prev_slots is None:
            slots = self.generate_initial(batch)
            self.prev_slots = slots.clone()
        else:
            slots = prev_slots
        needs_grad = torch.is_grad_enabled() and (X.requires_grad or slots.requires_grad or
                                                  any(p.requires_grad for p in self.parameters()))
        if X.is_cuda and not needs_grad:
            out, attn = self._kernel_forward(X, slots)
            self.last_attention = attn
            return out
        if not X.is_cuda and not needs_grad:
            raise RuntimeError("SlotAttention.forward runs on the B200 kernels: tensors must be on CUDA "
                               "(rl_sandbox_b200 has no CPU fallback)")
        if X.is_cuda and self.kernel_backward:
            return self._kernel_autograd_forward(X, slots)
        return self._autograd_forward(X, slots)

    kernel_backward = True   # False: differentiate through the torch-op restatement below

    def _kernel_autograd_forward(self, X, slots):
        """Training: K3 forward with an activation tape, K3 backward (rlsb_slot_attention_bwd) under torch autograd."""
        from rl_sandbox_b200 import ops
        self._kernel_forward_prepare(X)
        names = list(ops.SlotAttentionEngine.KEYS.values())
        sd = dict(self.named_parameters())
        out, attn = ops.SlotAttentionFn.apply(self._engine, names, X.float(), slots.float(), *[sd[n] for n in names])
        self.last_attention = attn
        return out

    def _autograd_forward(self, X, slots):
        """Differentiable evaluation with torch ops (the reference's op sequence; CPU checks and kernel_backward=False)."""
        k, v = self.inputs_proj(self.inputs_norm(X)).chunk(2, dim=-1)
        self.last_attention = None
        for _ in range(self.n_iter):
            prev = slots
            q = self.slots_proj(self.slots_norm(slots))
            attn = F.softmax(self.scale * torch.einsum('bik,bjk->bij', q, k).float(), dim=1) + self.epsilon
            attn = attn / attn.sum(dim=-1, keepdim=True)
            self.last_attention = attn
            updates = torch.einsum('bjd,bij->bid', v, attn)
            slots = self.slots_reccur(updates.reshape(-1, self.n_dim), prev.reshape(-1, self.n_dim))
            slots = slots.reshape(X.shape[0], self.n_slots, self.n_dim)
            slots = slots + self.slots_proj_2(self.slots_norm_2(slots))
        return slots


def build_grid(resolution):
    """(1, H, W, 4) grid of (y, x, 1-y, 1-x) in [0, 1] (slot_attention.py:79-86)."""
    axes = [np.linspace(0.0, 1.0, num=r) for r in resolution]
    grid = np.stack(np.meshgrid(*axes, sparse=False, indexing="ij"), axis=-1)
    grid = grid.reshape(resolution[0], resolution[1], -1)[None].astype(np.float32)
    return np.concatenate([grid, 1.0 - grid], axis=-1)


class PositionalEmbedding(nn.Module):
    def __init__(self, n_dim: int, res: t.Tuple[int, int], channel_last=False):
        super().__init__()
        self.n_dim = n_dim
        self.proj = nn.Linear(4, n_dim)
        self.channel_last = channel_last
        self.register_buffer('grid', torch.from_numpy(build_grid(res)))

    def forward(self, X) -> torch.Tensor:
        emb = self.proj(self.grid)
        return X + (emb if self.channel_last else emb.permute(0, 3, 1, 2))
