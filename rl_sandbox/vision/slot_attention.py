"""Alias of the reference dotted path `rl_sandbox.vision.slot_attention` (drop-in boundary)."""
from rl_sandbox_b200.vision.slot_attention import PositionalEmbedding, SlotAttention, build_grid  # noqa: F401
