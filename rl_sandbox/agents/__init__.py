from rl_sandbox_b200.agents.dreamer_v2 import DreamerV2  # noqa: F401
