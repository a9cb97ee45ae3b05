"""Alias of the reference dotted path `rl_sandbox.agents.dreamer.ac` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer.ac import *  # noqa: F401,F403
from rl_sandbox_b200.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic  # noqa: F401
