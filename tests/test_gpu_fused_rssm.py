"""The chained rollout's fused RSSM epilogues (rlsb_set_fused_rssm, default on): LayerNorm + ELU of img_in / prior1 and the whole
GRU cell (common.py:69-81) run inside their contractions even though a row spans several n-blocks — the blocks' CTAs exchange
their row statistics through global memory as tagged 64-bit words (GemmParams::xstats) — against the unfused
chain (contraction -> fp32 pre-activations -> ln_act_kernel / gru_gate_kernel) and the oracle.

Same arithmetic, different summation order of the LayerNorm statistics (and 192- instead of 256-column blocks for the GRU):
the two agree to fp32 rounding before each bf16 re-quantisation, so trajectories are compared on the rows whose draws coincide.
The reference-golden parity tests of tests/test_gpu_imagine.py (c1, c1_long: D = 1024) run the fused path as the default.
"""
import pytest
import torch

from oracle import oracle_port as orc
from tests._golden import load_case
from tests.test_gpu_imagine import engine, rel_rms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from rl_sandbox_b200 import ops as _ops
    return _ops


@pytest.fixture()
def lib(cuda):
    from rl_sandbox_b200 import _lib
    lib = _lib.load()
    before = lib.rlsb_set_fused_rssm(-1)
    yield lib
    lib.rlsb_set_fused_rssm(before)


def _run(ops, lib, cuda, c, H, n, fused, h0, z0, lat, act, **kw):
    assert lib.rlsb_set_fused_rssm(1 if fused else 0) == (1 if fused else 0)
    eng = engine(ops, c["meta"], cuda, c, H, "chained")
    args = (h0.to(cuda), z0.to(cuda), None, lat.to(cuda), act.to(cuda))
    eng.rollout(*args, horizon=H, **kw)     # builds the chained blob under the switch in effect (lazy pack)
    l0 = lib.rlsb_launch_count(0)
    out = eng.rollout(*args, horizon=H, **kw)
    torch.cuda.synchronize()
    launches = lib.rlsb_launch_count(0) - l0
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}, launches


@pytest.mark.parametrize("n", [1, 128, 129, 300, 800, 4096])
def test_fused_matches_unfused_and_oracle(ops, lib, cuda, n):
    """one row, an exact row block, an odd number of row blocks (a padding CTA in the last CTA pair), the configured 800
    start states and 32 row blocks (more work items than CTA pairs: the self-resetting counters across work rounds)"""
    c = load_case("c1")
    m = c["meta"]
    H = 4
    h0, z0 = orc.make_start(n + 11, n, m["D"])
    g = torch.Generator().manual_seed(n)
    lat, act = torch.rand(H, n, 1024, generator=g), torch.rand(H, n, m["A"], generator=g)
    a, la = _run(ops, lib, cuda, c, H, n, False, h0, z0, lat, act)
    b, lb = _run(ops, lib, cuda, c, H, n, True, h0, z0, lat, act)
    print(f"[fused rssm] n={n}: launches per rollout unfused {la}, fused {lb}")
    assert lb <= la - 3 * H      # gru_gate_kernel + 2 x ln_act_kernel per step are gone
    own = orc.sample_categorical(b["logits"][1:].cpu().view(H, n, 32, 32), lat.view(H, n, 32, 32))
    assert torch.equal(own, b["stoch_idx"][1:].cpu().long())
    same = (a["stoch_idx"] == b["stoch_idx"]).all(-1) & (a["actions"].argmax(-1) == b["actions"].argmax(-1))
    alive = same.cumprod(0).bool()
    frac = alive[-1].float().mean().item()
    print(f"[fused rssm] n={n}: rows with identical draws over {H} steps: {frac:.3f}")
    assert frac > 0.9 or n == 1
    # the first transition runs every fused layer once from identical inputs
    assert rel_rms(b["determ"][1], a["determ"][1], f"n={n} determ[1] fused vs unfused") < 1e-5
    assert rel_rms(b["logits"][1], a["logits"][1], f"n={n} logits[1] fused vs unfused") < 2e-3
    for k, lim in (("determ", 2e-4), ("logits", 2e-3), ("rewards", 1e-2), ("values", 1e-2)):
        if alive.any():
            assert rel_rms(b[k][alive], a[k][alive], f"n={n} {k} fused vs unfused") < lim
    assert torch.isfinite(b["determ"]).all() and torch.isfinite(b["logits"]).all()
    if n <= 300:
        ref = orc.imagine(c["wm"], c["actor"], c["critic"], h0, z0, H=H, A=m["A"], discrete=True, predict_discount=True,
                          latent_uniforms=lat, action_noise=act, bf16=True)
        same = ((b["stoch_idx"].cpu().long() == ref["stoch_idx"]).all(-1)
                & (b["actions"].cpu().argmax(-1) == ref["actions"].argmax(-1))).cumprod(0).bool()
        assert same[-1].float().mean() > 0.9 or n == 1
        assert rel_rms(b["determ"].cpu()[same], ref["determ"][same], f"n={n} determ fused vs bf16 oracle") < 1e-3
        assert rel_rms(b["logits"].cpu()[same], ref["logits"][same], f"n={n} logits fused vs bf16 oracle") < 2e-3


def test_fused_keeps_packed_state_images(ops, lib, cuda):
    """the per-step packed bf16 images of h the actor-critic update reads afterwards (keep_packed) come from the GRU epilogue's
    bulk stores: identical to the unfused gate kernel's up to the rounding of a bf16 value at a rounding boundary"""
    from tests.test_gpu_rollout import unpack_image
    c = load_case("c1")
    m = c["meta"]
    H, n = 3, 300
    h0, z0 = orc.make_start(5, n, m["D"])
    g = torch.Generator().manual_seed(6)
    lat, act = torch.rand(H, n, 1024, generator=g), torch.rand(H, n, m["A"], generator=g)
    a, _ = _run(ops, lib, cuda, c, H, n, False, h0, z0, lat, act, keep_packed=True)
    b, _ = _run(ops, lib, cuda, c, H, n, True, h0, z0, lat, act, keep_packed=True)
    kpad = b["determ_packed"].shape[-1]
    for t in range(H + 1):
        img = unpack_image(b["determ_packed"][t], n, kpad)
        # the image is the bf16 rounding of the fp32 state the same kernel wrote
        assert torch.equal(img[:, :m["D"]], b["determ"][t].bfloat16().float())
        rows = b["determ_packed"].shape[1]
        if rows > n:     # padding rows of the last row block stay zero
            full = unpack_image(b["determ_packed"][t], rows, kpad)
            assert not full[n:].any()
    alive = (a["stoch_idx"] == b["stoch_idx"]).all(-1).cumprod(0).bool()
    x = torch.stack([unpack_image(a["determ_packed"][t], n, kpad) for t in range(H + 1)])
    y = torch.stack([unpack_image(b["determ_packed"][t], n, kpad) for t in range(H + 1)])
    assert rel_rms(y[alive], x[alive], "packed h fused vs unfused") < 5e-3


@pytest.mark.parametrize("M,Dx,D", [(1, 1024, 1024), (300, 1024, 1024), (4096, 1024, 1024), (333, 200, 256), (129, 70, 192)])
def test_gru_cell_op_matches_reference_module_math(ops, cuda, M, Dx, D):
    """`GRUCell.forward` (common.py:69-81) as one launch through the C ABI (rlsb_gru_cell_fwd) against the oracle's restatement:
    with bf16-rounded contraction operands (the tensor cores' arithmetic) to fp32 rounding, and against the fp32 module math
    within the bf16-contraction tolerance"""
    g = torch.Generator().manual_seed(M + D)
    sd = {"c._layer.weight": torch.randn(3 * D, Dx + D, generator=g) / (Dx + D) ** 0.5,
          "c._layer.bias": 0.1 * torch.randn(3 * D, generator=g),
          "c._norm.weight": 1.0 + 0.1 * torch.randn(3 * D, generator=g),
          "c._norm.bias": 0.1 * torch.randn(3 * D, generator=g)}
    x, h = torch.randn(M, Dx, generator=g), torch.tanh(torch.randn(M, D, generator=g))
    op = ops.GRUCellOp(Dx, D).pack(*(sd[k].to(cuda) for k in ("c._layer.weight", "c._layer.bias", "c._norm.weight", "c._norm.bias")))
    out = op.forward(x.to(cuda), h.to(cuda))
    again = op.forward(x.to(cuda), h.to(cuda))          # a second launch on the same statistics slots (tags advance)
    torch.cuda.synchronize()
    assert torch.equal(out, again)
    ref16 = orc.gru_cell(x, h, sd, "c.", bf16=True)
    ref32 = orc.gru_cell(x, h, sd, "c.", bf16=False)
    assert rel_rms(out, ref16, f"GRU cell M={M} Dx={Dx} D={D} vs bf16-operand oracle") < 2e-5
    assert rel_rms(out, ref32, f"GRU cell M={M} Dx={Dx} D={D} vs fp32 module math") < 3e-3
    from rl_sandbox_b200.ops import unpack_rows
    img = unpack_rows(op.h_packed, M, D, k_pad=D)
    assert torch.equal(img, out.bfloat16().float())


@pytest.mark.parametrize("D,ln,discrete", [(1024, False, True), (576, True, False), (640, True, True), (256, True, True)])
def test_fused_other_widths_and_no_layer_norm(ops, lib, cuda, D, ln, discrete):
    """layer_norm = False (the GRU's own LayerNorm stays; img_in / prior1 run NB > 1 blocks without an exchange), D = 576
    (three 192-column blocks per LayerNorm layer, nine GRU blocks), D = 640 (the LayerNorm layers do not tile into 64-column
    multiples: they stay unfused next to the fused GRU) and D = 256 (single-block LayerNorm layers, four GRU blocks)"""
    A, H, n = 7, 3, 300
    wm, actor, critic = orc.make_params(1000 + D, D=D, A=A, discrete=discrete, layer_norm=ln, predict_discount=False)
    h0, z0 = orc.make_start(D, n, D)
    g = torch.Generator().manual_seed(D)
    lat = torch.rand(H, n, 1024, generator=g)
    act = torch.rand(H, n, A, generator=g) if discrete else torch.randn(H, n, A, generator=g)
    to = lambda sd: {k: v.to(cuda) for k, v in sd.items()}
    outs = []
    for fused in (0, 1):
        assert lib.rlsb_set_fused_rssm(fused) == fused
        eng = ops.ImaginationEngine(ops.ImagineConfig(D=D, A=A, discrete=discrete, layer_norm=ln, predict_discount=False, H=H))
        eng.persistent_max_rows = 0
        eng.pack(to(wm), to(actor), to(critic))
        eng.rollout(h0.to(cuda), z0.to(cuda), None, lat.to(cuda), act.to(cuda), horizon=H)   # (lazy pack happens here)
        l0 = lib.rlsb_launch_count(0)
        o = eng.rollout(h0.to(cuda), z0.to(cuda), None, lat.to(cuda), act.to(cuda), horizon=H)
        torch.cuda.synchronize()
        outs.append(({k: (v.clone() if torch.is_tensor(v) else v) for k, v in o.items()}, lib.rlsb_launch_count(0) - l0))
    (a, la), (b, lb) = outs
    print(f"[fused rssm] D={D} ln={ln}: launches unfused {la}, fused {lb}")
    assert lb < la
    ref = orc.imagine(wm, actor, critic, h0, z0, H=H, A=A, discrete=discrete, predict_discount=False,
                      latent_uniforms=lat, action_noise=act, bf16=True)
    for name, o in (("unfused", a), ("fused", b)):
        same = (o["stoch_idx"].cpu().long() == ref["stoch_idx"]).all(-1)
        if discrete:
            same &= o["actions"].cpu().argmax(-1) == ref["actions"].argmax(-1)
        alive = same.cumprod(0).bool()
        assert alive[-1].float().mean() > 0.9
        assert rel_rms(o["determ"].cpu()[alive], ref["determ"][alive], f"D={D} ln={ln} determ {name} vs bf16 oracle") < 1e-3
        assert rel_rms(o["logits"].cpu()[alive], ref["logits"][alive], f"D={D} ln={ln} logits {name} vs bf16 oracle") < 2e-3
    assert rel_rms(b["determ"][1], a["determ"][1], f"D={D} determ[1] fused vs unfused") < 1e-5


def test_switch_change_after_pack_repacks(ops, lib, cuda):
    """the switch decides the row order of the packed GRU weight: the engine builds the chained blob lazily and rebuilds it when
    the switch changed since — a rollout never contracts rows packed in the other order"""
    c = load_case("c1")
    assert lib.rlsb_set_fused_rssm(1) == 1
    eng = engine(ops, c["meta"], cuda, c, 2, "chained")
    h0, z0 = orc.make_start(1, 8, c["meta"]["D"])
    a = eng.rollout(h0.to(cuda), z0.to(cuda), None, None, None, horizon=2, seed=1)["determ"].clone()
    lib.rlsb_set_fused_rssm(0)
    b = eng.rollout(h0.to(cuda), z0.to(cuda), None, None, None, horizon=2, seed=1)["determ"].clone()
    lib.rlsb_set_fused_rssm(1)
    c2 = eng.rollout(h0.to(cuda), z0.to(cuda), None, None, None, horizon=2, seed=1)["determ"].clone()
    assert torch.equal(a, c2) and rel_rms(b[1], a[1], "determ[1] unfused vs fused, same engine") < 1e-5
