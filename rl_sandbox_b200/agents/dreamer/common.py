"""Building blocks shared by the world models (reference: rl_sandbox/agents/dreamer/common.py:8-81)."""
import numpy as np
import torch
import torch.distributions as td
from torch import nn

from rl_sandbox_b200.utils.dists import DistLayer


def get_position_encoding(seq_len, d, n=10000):
    """Sinusoidal table (seq_len, d): even columns sin, odd columns cos (common.py:8-15)."""
    k = np.arange(seq_len, dtype=np.float64)[:, None]
    i = np.arange(d // 2, dtype=np.float64)[None, :]
    angle = k / np.power(float(n), 2 * i / d)
    table = np.zeros((seq_len, d))
    table[:, 0:2 * (d // 2):2] = np.sin(angle)
    table[:, 1:2 * (d // 2):2] = np.cos(angle)
    return table


class View(nn.Module):
    def __init__(self, shape):
        super().__init__()
        self.shape = shape

    def forward(self, x):
        return x.view(*self.shape)


def Dist(val):
    """32 independent one-hot categoricals with straight-through gradients (common.py:27-28)."""
    return td.Independent(DistLayer('onehot')(val), 1)


class Normalizer(nn.Module):
    """Running magnitude normaliser; with momentum 1.0 (world_model.py:111) it is the identity."""

    def __init__(self, momentum=0.99, scale=1.0, eps=1e-8):
        super().__init__()
        self.momentum, self.scale, self.eps = momentum, scale, eps
        self.register_buffer('mag', torch.ones(1, dtype=torch.float32))

    def update(self, x):
        # in place: the buffer keeps its address (a CUDA-graph replay of the caller reads and writes this tensor)
        with torch.no_grad():
            self.mag.mul_(self.momentum).add_((1 - self.momentum) * x.abs().mean())

    def forward(self, x):
        self.update(x)
        return (x / (self.mag + self.eps)) * self.scale


class GRUCell(nn.Module):
    """LayerNorm GRU of DreamerV2 (common.py:50-81): one Linear over cat[x, h] producing
    reset | candidate | update, LayerNorm over all 3D jointly in fp32, update bias -1.
    The imagination path evaluates this cell inside librlsb (rlsb_imagine.cu, gru_gate_kernel);
    this torch forward serves the observe loop of the world-model loss."""

    def __init__(self, input_size, hidden_size, norm=False, update_bias=-1, **kwargs):
        super().__init__()
        self._size = hidden_size
        self._update_bias = update_bias
        self._layer = nn.Linear(input_size + hidden_size, 3 * hidden_size, bias=norm is not None, **kwargs)
        self._norm = nn.LayerNorm(3 * hidden_size) if norm else None

    @property
    def state_size(self):
        return self._size

    def forward(self, x, h):
        gates = self._layer(torch.cat([x, h], -1))
        if self._norm is not None:
            gates = self._norm(gates.float()).to(gates.dtype)
        reset, cand, update = gates.chunk(3, dim=-1)
        cand = torch.tanh(torch.sigmoid(reset) * cand)
        update = torch.sigmoid(update + self._update_bias)
        out = update * cand + (1 - update) * h
        return out, out
