"""Alias of the reference dotted path `rl_sandbox.utils.replay_buffer` (drop-in boundary)."""
from rl_sandbox_b200.utils.replay_buffer import *  # noqa: F401,F403
