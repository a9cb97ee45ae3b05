"""rl_sandbox_b200 — B200-native (sm_100a) implementation of the DreamerV2 imagination +
lambda-return + actor-critic hot path of Midren/rl_sandbox, behind the reference's agent API.

Layout
  csrc/ + librlsb.so   hand-written CUDA kernels and the C ABI (include/rlsb.h)
  _lib.py, ops.py      ctypes binding and torch-facing wrappers (torch = device memory + streams)
  agents/, utils/      host-side mirror of the reference interface for this path
The package `rl_sandbox` at the repository root aliases the reference's dotted paths
(`rl_sandbox.agents.DreamerV2`, ...) onto these modules so the Hydra configs resolve unchanged.
"""
__version__ = "0.1.0"
