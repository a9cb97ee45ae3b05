"""Alias of the reference dotted path `rl_sandbox.utils.optimizer` (drop-in boundary)."""
from rl_sandbox_b200.utils.optimizer import *  # noqa: F401,F403
from rl_sandbox_b200.utils.optimizer import Optimizer, WarmupScheduler, DecayScheduler  # noqa: F401
