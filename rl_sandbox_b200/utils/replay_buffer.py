"""Boundary types of DreamerV2.train / preprocess (reference: rl_sandbox/utils/replay_buffer.py:21-56).

Only the dataclasses are provided: they are the input type of the hot path.  The replay buffer
itself is host-side data plumbing and out of scope."""
from dataclasses import dataclass, field, fields

import torch

Observation = torch.Tensor
Action = torch.Tensor
Observations = torch.Tensor
States = torch.Tensor
State = torch.Tensor
Actions = torch.Tensor
Rewards = torch.Tensor
TerminationFlags = torch.Tensor
IsFirstFlags = torch.Tensor


def unpack(obj):
    """What the reference gets from the third-party `unpackable.unpack` (dreamer_v2.py:161)."""
    return tuple(getattr(obj, f.name) for f in fields(obj))


@dataclass
class EnvStep:
    obs: Observation
    action: Action
    reward: float
    is_finished: bool
    is_first: bool
    additional_data: dict = field(default_factory=dict)


@dataclass
class Rollout:
    obs: Observations
    actions: Actions
    rewards: Rewards
    is_finished: TerminationFlags
    is_first: IsFirstFlags
    additional_data: dict = field(default_factory=dict)

    def __len__(self):
        return len(self.obs)

    def to(self, device, non_blocking: bool = False):
        for f in ("obs", "actions", "rewards", "is_finished", "is_first"):
            setattr(self, f, getattr(self, f).to(device, non_blocking=True))
        self.additional_data = {k: v.to(device, non_blocking=True) for k, v in self.additional_data.items()}
        if not non_blocking and torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        return self


@dataclass
class RolloutChunks(Rollout):
    pass
