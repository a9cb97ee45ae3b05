"""Alias of the reference dotted path `rl_sandbox.utils.schedulers` (drop-in boundary)."""
from rl_sandbox_b200.utils.schedulers import *  # noqa: F401,F403
