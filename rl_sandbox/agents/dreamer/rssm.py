"""Alias of the reference dotted path `rl_sandbox.agents.dreamer.rssm` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer.rssm import *  # noqa: F401,F403
from rl_sandbox_b200.agents.dreamer.rssm import RSSM, State  # noqa: F401
