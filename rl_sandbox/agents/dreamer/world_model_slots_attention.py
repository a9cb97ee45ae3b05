"""Alias of the reference dotted path `rl_sandbox.agents.dreamer.world_model_slots_attention` (drop-in boundary)."""
from rl_sandbox_b200.agents.dreamer.world_model_slots_attention import State, WorldModel  # noqa: F401
