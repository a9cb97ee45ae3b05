"""Per-phase timing of the persistent rollout kernel (rlsb_rollout_fwd) from its %globaltimer stamps (cluster 0).

usage: python scripts/rollout_trace.py [config1|dino] [rows]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from oracle import oracle_port as orc   # noqa: E402  (synthetic parameters only)
from rl_sandbox_b200 import _lib, ops   # noqa: E402

PHASES = ["head0", "head1", "head2", "head3", "head4", "readout+action", "img_in", "gru", "prior1", "prior2", "latent draw"]
which = sys.argv[1] if len(sys.argv) > 1 else "config1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 800
dims = dict(config1=dict(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True),
            dino=dict(D=200, A=12, discrete=False, layer_norm=False, predict_discount=False))[which]
H = 15
wm, actor, critic = orc.make_params(1, **dims)
eng = ops.ImaginationEngine(ops.ImagineConfig(H=H, with_backward=not dims["discrete"], **dims))
to = lambda sd: {k: v.cuda() for k, v in sd.items()}
eng.pack(to(wm), to(actor), to(critic))
h0, z0 = orc.make_start(3, n, dims["D"])
h0, z0 = h0.cuda(), z0.cuda()
lib = _lib.load()
trace = torch.zeros((H + 1) * 11 * 8, dtype=torch.int64, device="cuda")
kw = dict(persistent=True, tape=not dims["discrete"], keep_packed=True)
for it in range(3):
    out = eng.rollout(h0, z0, None, None, None, seed=5 + it, **kw)
torch.cuda.synchronize()
lib.rlsb_rollout_set_trace(C.c_void_p(trace.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = eng.rollout(h0, z0, None, None, None, seed=9, **kw)
e1.record()
torch.cuda.synchronize()
lib.rlsb_rollout_set_trace(None)
t = trace.view(H + 1, 11, 8).cpu().double()
print(f"{which}: {n} start states, cluster of {eng.rollout_cluster_for(n)}: call {e0.elapsed_time(e1) * 1e3:.0f} us, "
      f"kernel (first to last stamp) {(t[H, 5, 7] - t[0, 0, 0]) / 1e3:.0f} us")
print("mean over steps 1..H-1, us: start = previous phase finished")
print("| phase | -> first stage | -> MMAs issued | -> acc ready | -> stats | -> exchanged | -> stored | -> finished | total |")
print("|---|---|---|---|---|---|---|---|---|")
tot = 0.0
st = slice(1, H)
for i, name in enumerate(PHASES):
    prev = t[st, i - 1, 7] if i > 0 else t[0:H - 1, 10, 7]
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
