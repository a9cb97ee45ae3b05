"""Alias of the reference dotted path `rl_sandbox.agents.dreamer_v2` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer_v2 import *  # noqa: F401,F403
from rl_sandbox_b200.agents.dreamer_v2 import DreamerV2, ImaginativeActor, ImaginativeCritic, State  # noqa: F401
