from rl_sandbox_b200.agents.dreamer.common import *  # noqa: F401,F403
