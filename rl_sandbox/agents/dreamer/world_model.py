"""Alias of the reference dotted path `rl_sandbox.agents.dreamer.world_model` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer.world_model import *  # noqa: F401,F403
from rl_sandbox_b200.agents.dreamer.world_model import WorldModel, State  # noqa: F401
