"""Per-phase timing of the fused GRU cell kernel (EPI_GRU) from the %globaltimer stamps of every CTA's first tiles.

usage: python scripts/gru_cell_trace.py [rows] [D]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from rl_sandbox_b200 import _lib, ops   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = "cuda"
x = torch.randn(n, 2 * D, device=dev)
w = torch.randn(3 * D, 2 * D, device=dev) / (2 * D) ** 0.5
cell = ops.GRUCellOp(D, D).pack(w, None, None, None)
xp, hp = ops.pack_rows(x[:, :D].contiguous()), ops.pack_rows(x[:, D:].contiguous())
h_prev = x[:, D:].contiguous()
bufs = dict(h_next=torch.empty((n, D), device=dev), h_next_packed=torch.empty(ops.round_up(n, 128) * D, device=dev, dtype=torch.bfloat16))
for _ in range(3):
    cell.forward_packed(xp, hp, h_prev, n, **bufs)
torch.cuda.synchronize()
lib = _lib.load()
ctas = 148
trace = torch.zeros(ctas * 64 * 8, dtype=torch.int64, device=dev)
lib.rlsb_gemm_set_trace(C.c_void_p(trace.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
cell.forward_packed(xp, hp, h_prev, n, **bufs)
e1.record()
torch.cuda.synchronize()
lib.rlsb_gemm_set_trace(None)
t = trace.view(ctas, 64, 8).cpu().double()
used = (t[:, :, 7] > 0)
tiles = used.sum(1)
print(f"{n} rows, D={D}: {e0.elapsed_time(e1) * 1e3:.0f} us; tiles per CTA min {int(tiles.min())} max {int(tiles.max())}")
t0 = t[used][:, 0].min()
names = ["wait acc (0->1)", "copy out + stats (1->2)", "partials written (2->3)", "wait partners (3->4)", "totals (4->5)", "gates (5->6)",
         "store (6->7)"]
print("| phase | mean us (all CTAs, tiles 3..) | p90 | max |\n|---|---|---|---|")
sel = used.clone()
sel[:, :3] = False
for i, nm in enumerate(names):
    d = (t[:, :, i + 1] - t[:, :, i])[sel] / 1e3
    print(f"| {nm} | {d.mean():.2f} | {d.quantile(0.9):.2f} | {d.max():.2f} |")
per = (t[:, 1:, 0] - t[:, :-1, 0])[sel[:, 1:] & sel[:, :-1]] / 1e3
print(f"| tile period (0 -> next 0) | {per.mean():.2f} | {per.quantile(0.9):.2f} | {per.max():.2f} |")
ep = (t[:, :, 7] - t[:, :, 1])[sel] / 1e3
print(f"| epilogue busy (1->7) | {ep.mean():.2f} | {ep.quantile(0.9):.2f} | {ep.max():.2f} |")
# per-CTA view of one cluster pair (CTAs 0, 1) and one far CTA
for c in (0, 1, 73, 147):
    row = ((t[c, 3:12] - t0) / 1e3)
    print(f"CTA {c}: " + " | ".join(" ".join(f"{v:.1f}" for v in r[[0, 1, 2, 4, 7]]) for r in row))
