"""Alias of the reference dotted path `rl_sandbox.vision.dino` (drop-in boundary)."""
from rl_sandbox_b200.vision.dino import ViTFeat, VisionTransformer, vit_base, vit_small  # noqa: F401
