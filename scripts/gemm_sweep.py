"""Per-layer timing of the tcgen05 GEMM at the sweep shape for cluster sizes 1/2/4 (CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_sandbox_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def time_it(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps

layers = [("gru   K2048 N3072 stats", M, 2048, 3072, "stats"), ("img_in K1088 N1024 stats", M, 1088, 1024, "stats"),
          ("prior2 K1024 N1024 plain", M, 1024, 1024, "plain"), ("headL0 K2048 N400 ln (x4 groups)", 4 * M, 2048, 400, "ln"),
          ("hidden K448 N400 ln (x4 groups)", 4 * M, 448, 400, "ln"), ("headL4 K448 N17 plain (x4)", 4 * M, 448, 17, "plain")]
res = {}
for name, m, K, N, mode in layers:
    x = torch.randn(m, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
    rb, nb = ops.plan_blocks(N); kp = ops.round_up(K, 64)
    xp = ops.pack_rows(x); wp = ops.pack_rows(w, row_block=rb, rows_pad=rb * nb, k_pad=kp)
    del x
    m_pad = ops.round_up(m, 128)
    gam, bet, b = torch.ones(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, device=dev)
    if mode == "ln":
        outp = torch.empty(m_pad * ops.round_up(N, 64), device=dev, dtype=torch.bfloat16)
        import ctypes as C
        pad = lambda t, f: torch.cat([t, torch.full((rb - N,), f, device=dev)]).contiguous()
        bp, gp, ep = pad(b, 0.), pad(gam, 1.), pad(bet, 0.)
        fn = lambda: _lib.check(lib.rlsb_gemm_ln_act(xp.data_ptr(), kp, wp.data_ptr(), rb, bp.data_ptr(), m, N, gp.data_ptr(),
                                                      ep.data_ptr(), 1e-5, 1, outp.data_ptr(), ops.round_up(N, 64),
                                                      torch.cuda.current_stream().cuda_stream))
    else:
        bufs = dict(out=torch.empty((m_pad, N), device=dev), stats=torch.empty((nb, m_pad, 2), device=dev),
                    bias_p=torch.zeros(rb * nb, device=dev))
        fn = lambda: ops.gemm_bias(xp, kp, wp, rb, nb, None, m, N, want_stats=(mode == "stats"), **bufs)
    row = {}
    for cs in (1, 2, 4):
        lib.rlsb_set_cluster_size(cs)
        t = time_it(fn)
        row[cs] = t
    fl = 2.0 * m * K * N / 1e12
    print(f"{name:36s} " + "  ".join(f"cs{cs}: {row[cs]*1e3:7.1f} us ({fl/row[cs]*1e3:6.0f} TF/s)" for cs in row), flush=True)
    res[name] = row
json.dump(res, open("gpurun_out/gemm_sweep.json", "w"))
