"""Per-phase timing of the persistent backward rollout kernel (rlsb_rollout_bwd) from its %globaltimer stamps (cluster 0).

usage: python scripts/rollout_bwd_trace.py [rows]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from oracle import oracle_port as orc   # noqa: E402  (synthetic parameters only)
from rl_sandbox_b200 import _lib, ops   # noqa: E402

PHASES = ["head grads", "head4^T", "head3^T", "head2^T", "head1^T", "head0^T", "softmax bwd", "prior2^T", "prior1^T",
          "gate bwd", "gru x|h ^T", "img_in^T"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
dims = dict(D=200, A=12, discrete=False, layer_norm=False, predict_discount=False)
H = 15
wm, actor, critic = orc.make_params(1, **dims)
eng = ops.ImaginationEngine(ops.ImagineConfig(H=H, with_backward=True, **dims))
to = lambda sd: {k: v.cuda() for k, v in sd.items()}
eng.pack(to(wm), to(actor), to(critic))
h0, z0 = orc.make_start(3, n, dims["D"])
out = eng.rollout(h0.cuda(), z0.cuda(), None, None, None, seed=5, persistent=True, tape=True, keep_packed=True)
g = torch.Generator(device="cuda").manual_seed(1)
g_r, g_v = torch.randn(H + 1, n, device="cuda", generator=g), torch.randn(H + 1, n, device="cuda", generator=g)
lib = _lib.load()
for _ in range(3):
    eng.backward(out, g_r, g_v, persistent=True)
torch.cuda.synchronize()
trace = torch.zeros((H + 1) * 12 * 8, dtype=torch.int64, device="cuda")
lib.rlsb_rollout_set_trace(C.c_void_p(trace.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.backward(out, g_r, g_v, persistent=True)
e1.record()
torch.cuda.synchronize()
lib.rlsb_rollout_set_trace(None)
e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e2.record()
eng.backward(out, g_r, g_v, persistent=False)
e3.record()
torch.cuda.synchronize()
t = trace.view(H + 1, 12, 8).cpu().double()
print(f"{n} start states, cluster of {eng.rollout_cluster_for(n)}: persistent backward {e0.elapsed_time(e1) * 1e3:.0f} us "
      f"(chained, eager launches: {e2.elapsed_time(e3) * 1e3:.0f} us)")
print("mean over steps H-1..2, us: start = previous phase finished\n| phase | -> first stage | -> MMAs issued | -> acc ready | -> gradients | -> exchanged | -> stored | -> finished | total |\n|---|---|---|---|---|---|---|---|---|")
tot = 0.0
st = slice(2, H)
for i, name in enumerate(PHASES):
    prev = t[st, i - 1, 7] if i > 0 else t[3:H + 1, 11, 7]
    cells, last = [], prev
    for slot in (1, 2, 3, 4, 5, 6, 7):
        cur = t[st, i, slot]
        if bool((cur > 0).all()):
            cells.append(f"{(cur - last).mean().item() / 1e3:.2f}")
            last = cur
        else:
            cells.append("")
    total = (t[st, i, 7] - prev).mean().item() / 1e3
    tot += total
    print(f"| {name} | " + " | ".join(cells) + f" | {total:.2f} |")
print(f"| step | | | | | | | | {tot:.2f} |")
