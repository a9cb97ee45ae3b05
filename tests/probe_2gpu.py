"""Bring-up probe for the N>1 path (torchrun): reports where a non-finite value first appears."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rl_sandbox_b200.agents.dreamer.rssm import State

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
torch.backends.cuda.matmul.allow_tf32 = True
dims = bench.DIMS["sweep"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
agent = bench.build_agent(dims, 15, dev, 128)
g = torch.Generator(device=dev).manual_seed(1 + rank)
h0 = 0.5 * torch.randn(N, 1024, device=dev, generator=g)
z0 = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device=dev, generator=g), 32).float().view(N, 1024)
state = State(h0.unsqueeze(0), torch.zeros(1, N, 32, 32, device=dev), z0.unsqueeze(0))
for it in range(8):
    losses, metrics = agent.behaviour_update(state, noise={"seed": 1000 + it, "row_offset": rank * N})
    out = agent.last_rollout
    bad = {k: int((~torch.isfinite(v)).sum()) for k, v in out.items() if v is not None and v.dtype.is_floating_point}
    pbad = sum(int((~torch.isfinite(p)).sum()) for p in list(agent.actor.parameters()) + list(agent.critic.parameters()))
    print(f"[rank {rank}] it {it} loss_a {losses['loss_actor'].item():.4f} loss_c {losses['loss_critic'].item():.4f} "
          f"nonfinite rollout {bad} params {pbad}", flush=True)
dist.barrier()
dist.destroy_process_group()
