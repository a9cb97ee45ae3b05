"""Drop-in alias package: resolves the reference Hydra `_target_` strings onto rl_sandbox_b200."""
