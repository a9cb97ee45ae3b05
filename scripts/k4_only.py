"""K4 alone at the benched shape (32 768 start states x H = 15, config-1 dims) on synthetic state images: timing of the
update's stages with CUDA events and a target for `ncu -k regex:gemm_kernel<3` (the dX GEMM with the fused ELU' /
LayerNorm-backward epilogue).  usage: python scripts/k4_only.py [rows] [updates]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sandbox_b200 import ops  # noqa: E402
from rl_sandbox_b200.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, D, A = 15, 1024, 17
dev = "cuda"
torch.manual_seed(0)
cfg = ops.ImagineConfig(D=D, A=A, discrete=True, layer_norm=True, predict_discount=True, H=H)
actor = ImaginativeActor(latent_dim=D + 1024, actions_num=A, is_discrete=True, layer_norm=True, reinforce_fraction=None,
                         entropy_scale=3e-3).to(dev)
critic = ImaginativeCritic(discount_factor=0.999, update_interval=100, soft_update_fraction=1, value_target_lambda=0.95,
                           latent_dim=D + 1024, layer_norm=True).to(dev)
ac = ops.ACUpdateEngine(cfg, rho=1.0, eta=3e-3, metrics_samples=128)
ac.pack(actor.state_dict(), critic.state_dict())
rows = ops.round_up(N, 128)
g = torch.Generator(device=dev).manual_seed(1)
k1 = {"determ": torch.empty((H + 1, N, 1), device=dev),
      "determ_packed": (0.5 * torch.randn((H + 1, rows, D), device=dev, generator=g)).bfloat16(),
      "stoch_packed": (torch.rand((H + 1, rows, 1024), device=dev, generator=g) < 1 / 32).bfloat16(),
      "values": torch.randn((H + 1, N), device=dev, generator=g),
      "actions": torch.nn.functional.one_hot(torch.randint(0, A, (H + 1, N), device=dev, generator=g), A).float()}
vs = torch.randn((H, N), device=dev, generator=g)
w = torch.rand((H + 1, N), device=dev, generator=g)
for _ in range(2):
    ac.update(k1, vs, w, actor.actor, critic.critic, seed=1, horizon=H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    ac.update(k1, vs, w, actor.actor, critic.critic, seed=2 + i, horizon=H)
e1.record()
torch.cuda.synchronize()
print(f"K4 update, {N} start states x H={H}: {e0.elapsed_time(e1) / reps:.3f} ms per update")
