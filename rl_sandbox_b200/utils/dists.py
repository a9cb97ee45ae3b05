"""Output distributions of the heads (reference: rl_sandbox/utils/dists.py:108-129, 168-203).

Only the variants the three shipped configs reach are provided: 'mse', 'onehot', 'normal_trunc',
'binary'.  Semantics kept from the reference, including its quirks:
  * TruncatedNormal overrides ``sample`` (clamped) but inherits the UNCLAMPED ``rsample`` of Normal;
  * 'onehot' is OneHotCategoricalStraightThrough on fp32 logits and is NOT wrapped in Independent;
  * every other head is Independent(..., 1);
  * argument validation is off, as the reference's train.py:38 switches it off globally (the
    discount head is trained on gamma*(1-done) targets, which are not booleans).
"""
import torch
import torch.distributions as td
from torch import nn
from torch.distributions.utils import _standard_normal


class TruncatedNormal(td.Normal):
    def __init__(self, loc, scale, low=-1.0, high=1.0, eps=1e-6):
        super().__init__(loc, scale, validate_args=False)
        self.low, self.high, self.eps = low, high, eps

    def _clamp(self, x):
        hard = torch.clamp(x, self.low + self.eps, self.high - self.eps)
        return x - x.detach() + hard.detach()

    def sample(self, sample_shape=torch.Size(), clip=None):
        noise = _standard_normal(self._extended_shape(sample_shape), dtype=self.loc.dtype, device=self.loc.device)
        noise = noise * self.scale
        if clip is not None:
            noise = torch.clamp(noise, -clip, clip)
        return self._clamp(self.loc + noise)


def trunc_normal_params(raw: torch.Tensor, min_std: float = 0.1):
    """(loc, scale) of the continuous actor head: tanh(mean), 2*sigmoid(std/2)+min_std (dists.py:187-190)."""
    mean, std = raw.chunk(2, dim=-1)
    return torch.tanh(mean).float(), (2 * torch.sigmoid(std / 2) + min_std).float()


class DistLayer(nn.Module):
    KINDS = ('mse', 'onehot', 'normal_trunc', 'binary')

    def __init__(self, type: str):
        super().__init__()
        if type not in self.KINDS:
            raise RuntimeError("Invalid dist layer")
        self._dist = type

    def forward(self, x):
        if self._dist == 'onehot':
            return td.OneHotCategoricalStraightThrough(logits=x.float(), validate_args=False)
        if self._dist == 'mse':
            base = td.Normal(x.float(), torch.ones((), device=x.device), validate_args=False)
        elif self._dist == 'binary':
            base = td.Bernoulli(logits=x.float(), validate_args=False)  # targets are gamma*(1-done), not {0,1}
        else:
            base = TruncatedNormal(*trunc_normal_params(x))
        return td.Independent(base, 1, validate_args=False)
