"""CPU tests of the boundary: the C-ABI library builds, loads, and exports every symbol that
include/rlsb.h declares.  No compute calls (no GPU here)."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

from rl_sandbox_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 14
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in include/rlsb.h but not exported: {missing}"
    assert lib.rlsb_abi_version() == _lib.ABI_VERSION == 6


def test_no_torch_types_in_abi():
    import re
    hdr = (ROOT / "include" / "rlsb.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)       # declarations only (comments cite the reference)
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code
    assert 'extern "C"' in code


def test_sass_contains_blackwell_instructions():
    sass = subprocess.check_output(["cuobjdump", "-sass", str(_lib.LIB_PATH)], text=True)
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTCHMMA.2CTA" in sass, "tcgen05.mma.cta_group::2 (CTA-pair MMA of the wide contractions) missing from SASS"
    assert "UBLKCP" in sass, "bulk (TMA engine) copies missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "sm_100a" in subprocess.check_output(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], text=True)


def test_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rl_sandbox_b200 import ops
    with pytest.raises(_lib.RlsbError):
        ops.lambda_return(torch.zeros(4, 8), torch.zeros(4, 8), torch.ones(4, 8), 0.95)
    assert _lib.load().rlsb_check_device() != 0


def test_argument_errors_are_negative_codes():
    lib = _lib.load()
    assert lib.rlsb_lambda_return_fwd(None, None, None, 16, 8, 0.95, None, None, None, 0, None) < 0
    cfg = _lib.ImagineCfg(1024, 32, 31, 17, 400, 1, 1, 1, 1, 15, 1)   # classes != 32 -> unsupported
    assert lib.rlsb_imagine_packed_bytes(C.byref(cfg)) == 0
    cfg = _lib.ImagineCfg(1024, 32, 32, 17, 400, 1, 1, 1, 1, 15, 1)
    assert lib.rlsb_imagine_packed_bytes(C.byref(cfg)) > 20_000_000
    assert lib.rlsb_imagine_workspace_bytes(C.byref(cfg), 800) > 0
    # split-operand mode (cfg.parity): three weight blocks per input segment, extra residual images in the workspace;
    # flat RSSM and forward only
    plain = lib.rlsb_imagine_packed_bytes(C.byref(cfg))
    par = _lib.ImagineCfg(1024, 32, 32, 17, 400, 1, 1, 1, 1, 15, 1, parity=1)
    assert 2.9 * plain < lib.rlsb_imagine_packed_bytes(C.byref(par)) < 3.1 * plain
    assert lib.rlsb_imagine_workspace_bytes(C.byref(par), 800) > lib.rlsb_imagine_workspace_bytes(C.byref(cfg), 800)
    assert lib.rlsb_imagine_packed_bytes(C.byref(_lib.ImagineCfg(200, 32, 32, 1, 400, 0, 1, 0, 1, 15, 1, slots=4, parity=1))) == 0
    assert lib.rlsb_imagine_packed_bytes(C.byref(_lib.ImagineCfg(200, 32, 32, 12, 400, 0, 0, 0, 1, 15, 1, with_backward=1, parity=1))) == 0


@pytest.mark.gpu   # the workspace plan sizes the weight-gradient partials by the SM count of the device
def test_actor_slots_lie_inside_the_update_workspace():
    """rlsb_ac_actor_slots is pointer arithmetic over the K4 workspace plan: the twelve slices the rollout fills (layer outputs,
    x_hat, 1/std of the actor's four hidden layers) are disjoint, ordered and inside rlsb_ac_workspace_bytes."""
    lib = _lib.load()
    H, N = 15, 800
    cfg = _lib.AcCfg(1024, 32, 32, 17, 400, 1, 1, H, 1.0, 3e-3, 128, 0)
    total = lib.rlsb_ac_workspace_bytes(C.byref(cfg), N)
    assert total > 0
    base = 1 << 40   # a made-up device address: nothing is dereferenced
    slots = _lib.ActorSlots()
    assert lib.rlsb_ac_actor_slots(C.byref(cfg), N, C.c_void_p(base), C.byref(slots)) == 0
    assert slots.m_pad == 896 and slots.Hp == 448 and slots.steps == H
    img = H * slots.m_pad * slots.Hp * 2        # one group's packed bf16 image of a layer
    spans = []
    for l in range(4):
        spans += [(slots.x[l], img), (slots.pre[l], img), (slots.rstd[l], H * slots.m_pad * 4)]
    spans.sort()
    for (a, n), (b, _) in zip(spans, spans[1:]):
        assert a + n <= b, "actor slices overlap"
    assert spans[0][0] >= base and spans[-1][0] + spans[-1][1] <= base + total
    assert lib.rlsb_ac_actor_slots(None, N, C.c_void_p(base), C.byref(slots)) < 0


def test_staged_output_switch_is_queryable_without_a_device():
    lib = _lib.load()
    cur = lib.rlsb_set_staged_output(-1)
    assert cur in (0, 1)
    assert lib.rlsb_set_staged_output(0) == 0 and lib.rlsb_set_staged_output(1) == 1
    lib.rlsb_set_staged_output(cur)


def test_persistent_rollout_plan_sizes_without_a_device():
    """rlsb_rollout_packed_bytes is host arithmetic (per-CTA weight slabs of the persistent rollout kernel): every cluster
    size yields a blob close to the chained one's weights (same matrices, re-ordered and padded per CTA); unsupported
    combinations are refused with 0 — slotted RSSM, split-operand mode, GRU slices wider than the epilogue's register plan."""
    lib = _lib.load()
    assert lib.rlsb_rollout_cluster_size() in (4, 8, 16)
    c1 = dict(D=1024, groups=32, classes=32, A=17, hidden=400, discrete=1, layer_norm=1, predict_discount=1, with_critic=1, H=15,
              discount_nan_on_tie=1)
    c2 = dict(D=200, groups=32, classes=32, A=12, hidden=400, discrete=0, layer_norm=0, predict_discount=0, with_critic=1, H=15,
              discount_nan_on_tie=1, with_backward=1)
    chained = lib.rlsb_imagine_packed_bytes(C.byref(_lib.ImagineCfg(**c1)))
    for cs in (8, 16):
        n = lib.rlsb_rollout_packed_bytes(C.byref(_lib.ImagineCfg(**c1, rollout_cluster=cs)))
        assert 0.95 * chained < n < 1.15 * chained, (cs, n, chained)
    assert lib.rlsb_rollout_packed_bytes(C.byref(_lib.ImagineCfg(**c1, rollout_cluster=4))) == 0   # 256 hidden units per CTA
    sizes = [lib.rlsb_rollout_packed_bytes(C.byref(_lib.ImagineCfg(**dict(c2, with_backward=0), rollout_cluster=cs))) for cs in (4, 8, 16)]
    assert all(s > 7_000_000 for s in sizes) and max(sizes) < 1.1 * min(sizes)         # same matrices, a little padding
    with_bwd = [lib.rlsb_rollout_packed_bytes(C.byref(_lib.ImagineCfg(**c2, rollout_cluster=cs))) for cs in (4, 8, 16)]
    assert with_bwd[0] == sizes[0] and all(1.5 * a < b < 2.2 * a for a, b in zip(sizes[1:], with_bwd[1:]))   # + transposed slabs
    assert lib.rlsb_rollout_packed_bytes(C.byref(_lib.ImagineCfg(**c1, rollout_cluster=0))) > 0  # 0 = library default
    # the persistent backward kernel: continuous-action configs whose slices fit its epilogue's register plan
    assert [lib.rlsb_rollout_bwd_supported(C.byref(_lib.ImagineCfg(**c2, rollout_cluster=cs))) for cs in (4, 8, 16)] == [0, 1, 1]
    assert lib.rlsb_rollout_bwd_supported(C.byref(_lib.ImagineCfg(**c1, rollout_cluster=16))) == 0    # no with_backward
    assert lib.rlsb_rollout_bwd(None, None, 0, None, None, None, None, None, None) < 0
    assert lib.rlsb_rollout_packed_bytes(C.byref(_lib.ImagineCfg(**c1, parity=1))) == 0
    assert lib.rlsb_rollout_packed_bytes(C.byref(_lib.ImagineCfg(**dict(c2, with_backward=0), slots=4, attention_blocks=3))) == 0
    assert lib.rlsb_rollout_fwd(None, None, 0, None, None, None, None, None, None, None) < 0


def test_fused_rssm_switch_and_gru_cell_sizes():
    """host logic of the round-2 entry points (no GPU): the fused-epilogue switch is queryable / settable and decides whether
    the workspace carries the fp32 pre-activation scratch; the stand-alone GRU cell reports its blob / workspace sizes and
    rejects the sizes its kernel does not cover"""
    from rl_sandbox_b200 import ops
    lib = _lib.load()
    before = lib.rlsb_set_fused_rssm(-1)
    try:
        cfg = ops.ImagineConfig(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True, H=15).to_c()
        assert lib.rlsb_set_fused_rssm(1) == 1
        fused = lib.rlsb_imagine_workspace_bytes(C.byref(cfg), 32768)
        packed = lib.rlsb_imagine_packed_bytes(C.byref(cfg))
        assert lib.rlsb_set_fused_rssm(0) == 0 and lib.rlsb_set_fused_rssm(7) == 0      # other values only query
        unfused = lib.rlsb_imagine_workspace_bytes(C.byref(cfg), 32768)
        assert unfused - fused > 32768 * 3 * 1024 * 4 * 0.9          # the 3 D fp32 scratch per row is gone
        assert lib.rlsb_imagine_packed_bytes(C.byref(cfg)) == packed  # same blob size: the GRU rows are only permuted
        assert lib.rlsb_set_fused_rssm(2) == 2
        small = ops.ImagineConfig(D=200, A=6, discrete=True, layer_norm=True, predict_discount=True, H=15).to_c()
        a = lib.rlsb_imagine_workspace_bytes(C.byref(small), 800)
        lib.rlsb_set_fused_rssm(1)
        assert lib.rlsb_imagine_workspace_bytes(C.byref(small), 800) == a   # D % 64 != 0: the unfused GRU either way
    finally:
        lib.rlsb_set_fused_rssm(before)
    assert lib.rlsb_gru_cell_packed_bytes(1024, 1024) >= 3 * 1024 * 2048 * 2 + 3 * 3 * 1024 * 4
    assert lib.rlsb_gru_cell_packed_bytes(70, 192) >= 3 * 192 * (128 + 192) * 2       # x padded to 128 columns
    assert lib.rlsb_gru_cell_packed_bytes(200, 200) == 0 and lib.rlsb_gru_cell_packed_bytes(64, 128) == 0
    assert lib.rlsb_gru_cell_workspace_bytes(1024, 300) == 16 * 384 * 16               # [D / 64][m_pad] x two 64-bit words
    assert lib.rlsb_gru_cell_workspace_bytes(1000, 300) == 0
    assert lib.rlsb_gru_cell_fwd(None, 1024, 1024, None, None, None, 1, -1.0, 1e-5, None, None, None, None) < 0
