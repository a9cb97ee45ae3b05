"""Alias of the reference dotted path `rl_sandbox.agents.dreamer.vision` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer.vision import *  # noqa: F401,F403
