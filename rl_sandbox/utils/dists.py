"""Alias of the reference dotted path `rl_sandbox.utils.dists` (drop-in boundary)."""
from rl_sandbox_b200.utils.dists import *  # noqa: F401,F403
