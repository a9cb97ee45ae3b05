"""K1 rollout of N start states as one launch sequence vs two half-batches on two streams (do the HBM-bound
elementwise kernels of one half overlap the tensor-bound GEMMs of the other?)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rl_sandbox_b200 import ops
from rl_sandbox_b200.agents.dreamer.rssm import State

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dims = bench.DIMS["sweep"]
dev = "cuda:0"
agent = bench.build_agent(dims, 15, dev, 128)
g = torch.Generator(device=dev).manual_seed(1)
h0 = 0.5 * torch.randn(N, dims["D"], device=dev, generator=g)
z0 = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device=dev, generator=g), 32).float().view(N, 1024)
logits0 = torch.zeros(N, 1024, device=dev)
with torch.no_grad():
    agent.imagine_trajectory(State(h0[None, :256], logits0[None, :256].view(1, 256, 32, 32), z0[None, :256]), noise={"seed": 1})
eng = agent._engine
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timed(fn, reps=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps

out1 = {}
def one():
    global out1
    out1 = eng.rollout(h0, z0, logits0, seed=5, keep_packed=True, want_stoch=False, out=out1 or None)
t1 = timed(one)
print(f"single stream N={N}: {t1:.2f} ms")
del out1
torch.cuda.empty_cache()

engs = [eng]
for _ in range(parts - 1):
    e = ops.ImaginationEngine(eng.cfg)
    e.packed = eng.packed
    engs.append(e)
streams = [torch.cuda.Stream() for _ in range(parts)]
per = N // parts
outs = [None] * parts
def split():
    cur = torch.cuda.current_stream()
    for i, (e, s) in enumerate(zip(engs, streams)):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            a = i * per
            outs[i] = e.rollout(h0[a:a + per], z0[a:a + per], logits0[a:a + per], seed=5, row_offset=a, keep_packed=True,
                                want_stoch=False, out=outs[i])
    for s in streams:
        cur.wait_stream(s)
t2 = timed(split)
print(f"{parts} streams x {per}: {t2:.2f} ms  ({t1 / t2:.3f}x)")
