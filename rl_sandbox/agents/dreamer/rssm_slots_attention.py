"""Alias of the reference dotted path `rl_sandbox.agents.dreamer.rssm_slots_attention` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer.rssm_slots_attention import RSSM, State  # noqa: F401
