"""Phase breakdown of the full DreamerV2.train() (bench workloads crafter / slotted): CUDA-event time per phase and the
GPU-busy time (sum of kernel durations from torch.profiler) — launch-bound phases show busy << elapsed."""
import statistics
import sys
import torch
sys.path.insert(0, ".")
import bench
from rl_sandbox_b200.utils.replay_buffer import RolloutChunks

wl = sys.argv[1] if len(sys.argv) > 1 else "crafter"
dims = bench.DIMS[wl]
dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = True
H, B, T = 15, 16, 50
N = B * T
agent = bench.build_agent(dims, H, dev, 128)
g = torch.Generator().manual_seed(1)
obs = agent.preprocess_obs(torch.randint(0, 255, (N, 64, 64, 3), dtype=torch.uint8, generator=g).to(dev))
A = dims["A"]
act = (torch.randint(0, A, (N, 1), generator=g) if dims["discrete"] else torch.randn(N, A, generator=g)).to(dev)
rew = torch.tanh(torch.randn(N, generator=g)).to(dev)
first = torch.zeros(N); first[::T] = 1
add = {"d_features": torch.randn(N, 384, 196, generator=g).to(dev)} if dims.get("slots") else {}
chunks = RolloutChunks(obs=obs, actions=act, rewards=rew, is_finished=torch.zeros(N, device=dev), is_first=first.to(dev), additional_data=add)
ev = lambda: torch.cuda.Event(enable_timing=True)

def phases():
    import torch.nn.functional as F
    e = [ev() for _ in range(4)]
    o, a, r, fin, fi, ad = chunks.obs, chunks.actions, chunks.rewards, chunks.is_finished, chunks.is_first, chunks.additional_data
    if agent.is_discrete:
        a = F.one_hot(a.to(torch.int64), num_classes=agent.actions_num).squeeze()
    disc = agent.critic.gamma * (1 - fin).float()
    e[0].record()
    losses_wm, st, m = agent.world_model.calculate_loss(o, a, r, disc, fi.float(), ad)
    e[1].record()
    agent.world_model_optimizer.step(losses_wm['loss_wm'])
    agent.mark_weights_changed()
    e[2].record()
    agent.behaviour_update(st.flatten().detach())
    e[3].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(3)]

for _ in range(3): phases()
ts = [phases() for _ in range(5)]
med = [statistics.median(t[i] for t in ts) for i in range(3)]
print(f"[{wl}] wm calculate_loss fwd {med[0]:.2f} ms | wm backward+clip+AdamW {med[1]:.2f} ms | behaviour_update {med[2]:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    phases()
    torch.cuda.synchronize()
ka = prof.key_averages()
busy = sum(k.device_time_total for k in ka if k.device_type == torch.autograd.DeviceType.CUDA) / 1e3
nk = sum(k.count for k in ka if k.device_type == torch.autograd.DeviceType.CUDA)
print(f"[{wl}] GPU busy (sum of kernel time) {busy:.2f} ms over {nk} kernels")
top = sorted((k for k in ka if k.device_type == torch.autograd.DeviceType.CUDA), key=lambda k: -k.device_time_total)[:14]
for k in top:
    print(f"   {k.device_time_total/1e3:8.2f} ms  x{k.count:5d}  {k.key[:90]}")
