"""Alias of the reference dotted path `rl_sandbox.agents.rl_agent` (drop-in boundary)."""
from rl_sandbox_b200.agents.rl_agent import *  # noqa: F401,F403
from rl_sandbox_b200.agents.rl_agent import RlAgent  # noqa: F401
