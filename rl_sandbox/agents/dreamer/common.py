"""Alias of the reference dotted path `rl_sandbox.agents.dreamer.common` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer.common import *  # noqa: F401,F403
