"""Achieved HBM bandwidth of the HBM-bound kernels outside the rollout: K2 (lambda-return + weights + advantage,
rlsb_lambda_return_fwd / _bwd) at sweep sizes and K3's per-frame attention (slot_attn_kernel inside rlsb_slot_attention_fwd).
Algorithmic bytes per unit as in SURVEY section 8(d) / DESIGN section 3.  Peak = MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_sandbox_b200 import ops, _lib
import bench

pk = bench.peaks()
dev = "cuda"
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

def timed(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # inputs leave the 126 MB L2
        torch.cuda.synchronize(); ev0.record(); fn(); ev1.record(); torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3

out = {}
H = 15
for N in (262144, 1048576, 4194304):
    g = torch.Generator(device=dev).manual_seed(N)
    r = torch.randn(H + 1, N, device=dev, generator=g)
    v = torch.randn(H + 1, N, device=dev, generator=g)
    d = (torch.rand(H + 1, N, device=dev, generator=g) > 0.05).float()
    t = timed(lambda: ops.lambda_return(r, v, d, 0.95))
    by = 372 * N                            # read r, v, d (16 each); write vs (15), w (16), adv (14) floats
    print(f"K2 fwd time-major  N={N:8d}: {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s = {by/t/1e9/pk['hbm']*100:5.1f} % of {pk['hbm']:.1f}")
    out[f"k2_fwd_{N}"] = dict(us=t * 1e6, gbs=by / t / 1e9, frac=by / t / 1e9 / pk["hbm"])
    vs, w, adv = ops.lambda_return(r, v, d, 0.95)
    gvs = torch.randn_like(vs)
    t = timed(lambda: ops.lambda_return_bwd(gvs, v, d, vs, 0.95))
    by = (15 + 16 + 16 + 15 + 3 * 16) * 4 * N   # read g_vs, v, d, vs; write g_r, g_v, g_d
    print(f"K2 bwd time-major  N={N:8d}: {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s = {by/t/1e9/pk['hbm']*100:5.1f} %")
    out[f"k2_bwd_{N}"] = dict(us=t * 1e6, gbs=by / t / 1e9, frac=by / t / 1e9 / pk["hbm"])
    rb, vb, db = r.t().contiguous(), v.t().contiguous(), d.t().contiguous()
    t = timed(lambda: ops.lambda_return(rb, vb, db, 0.95, batch_major=True))
    by = 372 * N
    print(f"K2 fwd batch-major N={N:8d}: {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s = {by/t/1e9/pk['hbm']*100:5.1f} %")
    out[f"k2_fwd_bm_{N}"] = dict(us=t * 1e6, gbs=by / t / 1e9, frac=by / t / 1e9 / pk["hbm"])
    del r, v, d, rb, vb, db, vs, w, adv, gvs

# K3: slot attention forward, config_slotted shape (4 slots, 384 dims, 196 tokens, 2 iterations)
from torch.profiler import profile, ProfilerActivity
for B in (800, 6400):
    eng = ops.SlotAttentionEngine(4, 384, 196, 2)
    sd = {k: torch.randn(*shape, device=dev) * 0.05 for k, shape in {
        "inputs_norm.weight": (384,), "inputs_norm.bias": (384,), "inputs_proj.weight": (768, 384),
        "slots_norm.weight": (384,), "slots_norm.bias": (384,), "slots_proj.weight": (384, 384),
        "slots_reccur.weight_ih": (1152, 384), "slots_reccur.weight_hh": (1152, 384), "slots_reccur.bias_ih": (1152,),
        "slots_reccur.bias_hh": (1152,), "slots_norm_2.weight": (384,), "slots_norm_2.bias": (384,),
        "slots_proj_2.0.weight": (1536, 384), "slots_proj_2.0.bias": (1536,), "slots_proj_2.2.weight": (384, 1536),
        "slots_proj_2.2.bias": (384,)}.items()}
    eng.pack(sd)
    X = torch.randn(B, 196, 384, device=dev)
    S = torch.randn(B, 4, 384, device=dev)
    t_all = timed(lambda: eng.forward(X, S))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            flush.zero_()
            eng.forward(X, S)
        torch.cuda.synchronize()
    ka = {k.key: k for k in prof.key_averages()}
    att = [k for n, k in ka.items() if "slot_attn_kernel" in n]
    t_att = sum(k.device_time_total for k in att) / sum(k.count for k in att) * 1e-6
    by = B * (2 * 196 * 384 * 2 + 2 * 4 * 384 * 4 + 4 * 196 * 4)   # k, v bf16 once; q in, updates out (fp32); attention out
    print(f"K3 forward B={B}: whole call {t_all*1e3:7.3f} ms; slot_attn_kernel {t_att*1e6:7.1f} us per iteration, "
          f"{by/t_att/1e9:7.1f} GB/s = {by/t_att/1e9/pk['hbm']*100:5.1f} % of {pk['hbm']:.1f}")
    out[f"k3_attn_{B}"] = dict(us=t_att * 1e6, gbs=by / t_att / 1e9, frac=by / t_att / 1e9 / pk["hbm"], whole_ms=t_all * 1e3)
json.dump(out, open("gpurun_out/hbm_kernels.json", "w"), indent=1)
