"""Alias of the reference dotted path `rl_sandbox.utils.fc_nn` (drop-in boundary)."""
from rl_sandbox_b200.utils.fc_nn import *  # noqa: F401,F403
