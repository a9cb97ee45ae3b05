"""hidden head layer (K=448, N=400, 131072 rows) under epilogue variants: which part of the LayerNorm + ELU epilogue costs what"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_sandbox_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
m, K, N = 4 * 32768, 448, 400
x = torch.randn(m, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
rb, nb = ops.plan_blocks(N); kp = ops.round_up(K, 64)
xp = ops.pack_rows(x); wp = ops.pack_rows(w, row_block=rb, rows_pad=rb * nb, k_pad=kp)
m_pad = ops.round_up(m, 128)
pad = lambda t, f: torch.cat([t, torch.full((rb - N,), f, device=dev)]).contiguous()
bp, gp, ep = pad(torch.zeros(N, device=dev), 0.), pad(torch.ones(N, device=dev), 1.), pad(torch.zeros(N, device=dev), 0.)
outp = torch.empty(m_pad * ops.round_up(N, 64), device=dev, dtype=torch.bfloat16)
outf = torch.empty((m_pad, N), device=dev)
stats = torch.empty((nb, m_pad, 2), device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s = lambda: torch.cuda.current_stream().cuda_stream
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps * 1e3
for name, gam, act in (("LN + ELU", gp, 1), ("LN + ReLU", gp, 2), ("LN only", gp, 0), ("ELU only (no LN)", None, 1), ("bias only -> bf16", None, 0)):
    g_ptr = gam.data_ptr() if gam is not None else None
    e_ptr = ep.data_ptr() if gam is not None else None
    fn = lambda: _lib.check(lib.rlsb_gemm_ln_act(xp.data_ptr(), kp, wp.data_ptr(), rb, bp.data_ptr(), m, N, g_ptr, e_ptr, 1e-5, act,
                                                  outp.data_ptr(), ops.round_up(N, 64), s()))
    print(f"{name:24s} {t(fn):7.1f} us", flush=True)
bufs = dict(out=outf, stats=stats, bias_p=bp)
print(f"{'plain fp32 out (RB=416)':24s} {t(lambda: ops.gemm_bias(xp, kp, wp, rb, nb, None, m, N, want_stats=False, **bufs)):7.1f} us")
