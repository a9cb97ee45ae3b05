"""hidden head layer shape with the trivial full-row epilogue (bias -> bf16 packed image), for ncu source-level captures"""
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
bp = torch.zeros(rb, device=dev)
outp = torch.empty(m_pad * ops.round_up(N, 64), device=dev, dtype=torch.bfloat16)
for _ in range(5):
    _lib.check(lib.rlsb_gemm_ln_act(xp.data_ptr(), kp, wp.data_ptr(), rb, bp.data_ptr(), m, N, None, None, 1e-5, 0, outp.data_ptr(),
                                    ops.round_up(N, 64), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
