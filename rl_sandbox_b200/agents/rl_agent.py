"""Agent interface (reference: rl_sandbox/agents/rl_agent.py:9-30)."""
from abc import ABCMeta, abstractmethod
from pathlib import Path
from typing import Any


class RlAgent(metaclass=ABCMeta):
    @abstractmethod
    def get_action(self, obs) -> Any:
        ...

    @abstractmethod
    def train(self, rollout_chunks) -> dict[str, Any]:
        """Returns a dict of losses / metrics for logging."""

    def reset(self):
        """Agents with internal state (the RSSM posterior) clear it between rollouts."""

    @abstractmethod
    def save_ckpt(self, epoch_num: int, losses: dict[str, float]):
        ...

    @abstractmethod
    def load_ckpt(self, ckpt_path: Path):
        ...
