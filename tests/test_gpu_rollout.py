"""The persistent rollout kernel (rlsb_rollout_fwd: ONE launch per rollout, a thread-block cluster per 128 start states)
against the chained rollout (rlsb_imagine_fwd: one launch per layer) and the oracle.

Both paths run the same arithmetic (bf16 tensor-core contractions, fp32 accumulation, fp32 LayerNorm / gates) but sum the
LayerNorm statistics in a different order, so they agree to fp32 rounding before each bf16 re-quantisation: trajectories
are compared on the rows whose draws coincide, and the draws themselves are checked exactly against the oracle sampler on
each kernel's own logits.  The reference-golden parity tests of tests/test_gpu_imagine.py run in both modes as well.
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


def unpack_image(img, n, kpad):
    """rows 0..n-1 of a packed row-block-128 bf16 tile image (rlsb_ptx.cuh::packed_index) as a (n, kpad) matrix"""
    r = torch.arange(n, device=img.device).view(n, 1)
    k = torch.arange(kpad, device=img.device).view(1, kpad)
    idx = ((r // 128) * (kpad // 64) + k // 64) * (128 * 64) + (r % 128) * 64 + ((((k % 64) // 8) ^ (r % 8)) * 8) + k % 8
    return img.reshape(-1)[idx.reshape(-1)].view(n, kpad).float()


def _both(ops, cuda, c, H, n=None, **kw):
    m = c["meta"]
    eng = engine(ops, m, cuda, c, H)
    n = n or m["N"]
    args = (c["h0"][:n].to(cuda), c["z0"][:n].to(cuda), None, c["lat"][:H, :n].contiguous().to(cuda),
            c["act"][:H, :n].contiguous().to(cuda))
    clone = lambda o: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in o.items()}
    from rl_sandbox_b200 import _lib
    lib = _lib.load()
    l0 = lib.rlsb_launch_count(0)
    chained = clone(eng.rollout(*args, horizon=H, persistent=False, **kw))
    l1 = lib.rlsb_launch_count(0)
    pers = clone(eng.rollout(*args, horizon=H, persistent=True, **kw))
    l2 = lib.rlsb_launch_count(0)
    torch.cuda.synchronize()
    return chained, pers, l1 - l0, l2 - l1


@pytest.mark.parametrize("name", ["c1", "c2", "c1_long", "c2_long"])
def test_persistent_matches_chained(ops, cuda, name):
    c = load_case(name)
    m = c["meta"]
    H, N = m["H"], m["N"]
    chained, pers, n_chained, n_pers = _both(ops, cuda, c, H, want_actor_raw=True)
    print(f"[rollout] {name}: launches chained {n_chained}, persistent {n_pers}")
    assert n_pers <= 12 and n_pers < H + 12 and n_chained > 8 * H      # start-state prep + ONE rollout kernel
    # exact: every draw is the oracle sampler's on the kernel's own logits
    own = orc.sample_categorical(pers["logits"][1:].cpu().view(H, N, 32, 32), c["lat"][:H].view(H, N, 32, 32))
    assert torch.equal(own, pers["stoch_idx"][1:].cpu().long())
    same = (pers["stoch_idx"] == chained["stoch_idx"]).all(-1)
    if m["discrete"]:
        same &= pers["actions"].argmax(-1) == chained["actions"].argmax(-1)
    alive = same.cumprod(0).bool()
    frac = alive[-1].float().mean().item()
    print(f"[rollout] {name}: rows with identical draws over {H} steps: {frac:.3f}")
    assert frac > (0.8 if H > 5 else 0.9)
    for k, lim in (("determ", 2e-4), ("logits", 2e-3), ("rewards", 1e-2), ("values", 1e-2)):
        e = rel_rms(pers[k][alive], chained[k][alive], f"{name}.{k} persistent vs chained")
        assert e < lim, (k, e)
    # the first step is a single pass through every layer from identical inputs: near fp32 agreement before the draws
    e1 = rel_rms(pers["logits"][1], chained["logits"][1], f"{name}.logits[1] persistent vs chained")
    assert e1 < 2e-3
    for k in ("rewards", "values"):
        assert rel_rms(pers[k][0], chained[k][0]) < 2e-3
    assert torch.equal(pers["determ"][0], chained["determ"][0]) and torch.equal(pers["stoch"][0], chained["stoch"][0])
    assert not pers["actions"][0].any() and bool((pers["discounts"][0] == 1).all())
    st = pers["stoch"].view(H + 1, N, 32, 32)
    assert torch.equal(st.sum(-1), torch.ones_like(st.sum(-1))) and torch.equal(st.argmax(-1), pers["stoch_idx"].long())
    if not m["discrete"]:
        assert rel_rms(pers["actor_raw"][0], chained["actor_raw"][0]) < 2e-3


@pytest.mark.parametrize("n", [1, 127, 129, 800, 1024, 2048])
def test_persistent_ragged_row_counts_vs_oracle(ops, cuda, n):
    """row blocks with padding rows, a single row, the configured 16 x 50 = 800 start states (7 clusters of 16), the last size
    with clusters of 16 (1024: 8 row blocks) and the engine's upper limit (2048: 16 clusters of 8)"""
    c = load_case("c2")
    m = c["meta"]
    H = 3
    eng = engine(ops, m, cuda, c, H, "persistent")
    h0, z0 = orc.make_start(n + 5, n, m["D"])
    g = torch.Generator().manual_seed(n)
    lat, act = torch.rand(H, n, 1024, generator=g), torch.randn(H, n, m["A"], generator=g)
    out = eng.rollout(h0.to(cuda), z0.to(cuda), None, lat.to(cuda), act.to(cuda), horizon=H)
    blocks = (n + 127) // 128
    assert eng.last_rollout_persistent and eng.rollout_cluster_for(n) == (16 if blocks <= eng.rollout_max_clusters(16) else 8)
    ref = orc.imagine(c["wm"], c["actor"], c["critic"], h0, z0, H=H, A=m["A"], discrete=False, predict_discount=False,
                      latent_uniforms=lat, action_noise=act, bf16=True)
    same = (out["stoch_idx"].cpu().long() == ref["stoch_idx"]).all(-1).cumprod(0).bool()
    assert same[-1].float().mean() > 0.9
    assert rel_rms(out["determ"].cpu()[same], ref["determ"][same], f"n={n} determ vs bf16 oracle") < 1e-3
    assert rel_rms(out["logits"].cpu()[same], ref["logits"][same], f"n={n} logits vs bf16 oracle") < 2e-3
    assert torch.isfinite(out["determ"]).all() and torch.isfinite(out["logits"]).all()


def test_engine_picks_the_persistent_kernel_only_for_one_wave(ops, cuda):
    """a cluster lives inside one GPC, so the device keeps fewer clusters resident than #SMs / cluster; row blocks beyond that
    would run as a second wave at twice the time, and the engine takes the chained rollout instead"""
    c = load_case("c2")
    m = c["meta"]
    eng = engine(ops, m, cuda, c, 2)
    c16, c8 = eng.rollout_max_clusters(16), eng.rollout_max_clusters(8)
    print(f"[rollout] resident clusters: {c16} of 16 CTAs, {c8} of 8 CTAs")
    assert 1 <= c16 <= 148 // 16 and c16 <= c8 <= 148 // 8
    for n in (128 * c16, 128 * c16 + 1, 128 * c8, 128 * c8 + 1):
        if n > eng.persistent_max_rows:
            continue
        h0, z0 = orc.make_start(n, n, m["D"])
        eng.rollout(h0.to(cuda), z0.to(cuda), None, None, None, horizon=2, seed=3)
        blocks = (n + 127) // 128
        assert eng.rollout_cluster_for(n) == (16 if blocks <= c16 else 8)
        assert eng.last_rollout_persistent == (blocks <= c8), n      # (D = 200: clusters of 8 still pay)
    big = engine(ops, load_case("c1")["meta"], cuda, load_case("c1"), 2)   # D = 1024: only with clusters of 16
    assert big.would_run_persistent(128 * c16) and not big.would_run_persistent(128 * c16 + 1)


@pytest.mark.parametrize("name", ["c2", "c2_ln"])
def test_persistent_keeps_packed_states_and_tape_like_chained(ops, cuda, name):
    """what the update (K4) and the backward pass (rlsb_imagine_bwd) read afterwards: the per-step packed state images and
    the activation tape (continuous-action configs).  With identical draws the images must agree to bf16 rounding, and
    d loss / d actions computed from the persistent kernel's tape must match the one from the chained rollout's."""
    c = load_case(name)
    m = c["meta"]
    assert not m["discrete"]
    H, N = 5, 300
    cfg = ops.ImagineConfig(D=m["D"], A=m["A"], discrete=False, layer_norm=m["layer_norm"], predict_discount=False, H=H,
                            with_backward=True)
    eng = ops.ImaginationEngine(cfg)
    to = lambda sd: {k: v.to(cuda) for k, v in sd.items()}
    eng.pack(to(c["wm"]), to(c["actor"]), to(c["critic"]))
    h0, z0 = orc.make_start(77, N, m["D"])
    g = torch.Generator().manual_seed(78)
    lat, act = torch.rand(H, N, 1024, generator=g), torch.randn(H, N, m["A"], generator=g)
    args = (h0.to(cuda), z0.to(cuda), None, lat.to(cuda), act.to(cuda))
    outs = {}
    for mode in (False, True):
        o = eng.rollout(*args, horizon=H, keep_packed=True, tape=True, persistent=mode)
        assert eng.last_rollout_persistent == mode
        outs[mode] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in o.items()}
    a, b = outs[False], outs[True]
    alive = (a["stoch_idx"] == b["stoch_idx"]).all(-1).cumprod(0).bool()
    frac = alive[-1].float().mean().item()
    print(f"[rollout] {name} (tape): rows with identical draws over {H} steps: {frac:.3f}")
    assert frac > 0.9
    for k, kpad in (("determ_packed", a["determ_packed"].shape[-1]), ("stoch_packed", a["stoch_packed"].shape[-1])):
        x = torch.stack([unpack_image(a[k][t], N, kpad) for t in range(H + 1)])
        y = torch.stack([unpack_image(b[k][t], N, kpad) for t in range(H + 1)])
        e = rel_rms(y[alive], x[alive], f"{k} persistent vs chained")
        assert e < 5e-3
        rows = a[k].shape[1]
        if rows > N:   # padding rows of the images are contraction indices of the update's weight gradients: zero
            pad = torch.stack([unpack_image(b[k][t], rows, kpad)[N:] for t in range(H + 1)])
            assert not pad.any(), k
    hh = torch.stack([unpack_image(b["determ_packed"][t], N, b["determ_packed"].shape[-1])[:, :m["D"]] for t in range(H + 1)])
    assert rel_rms(hh, b["determ"], "packed h image vs fp32 determ (bf16 rounding)") < 4e-3
    zz = torch.stack([unpack_image(b["stoch_packed"][t], N, 1024) for t in range(H + 1)])
    assert torch.equal(zz, b["stoch"])
    gen = torch.Generator(device="cuda").manual_seed(5)
    g_r = torch.randn(H + 1, N, device="cuda", generator=gen)
    g_v = torch.randn(H + 1, N, device="cuda", generator=gen)
    ga = eng.backward(a, g_r, g_v, persistent=False).clone()
    gb = eng.backward(b, g_r, g_v, persistent=False).clone()
    rows = alive[-1]
    e = rel_rms(gb[:, rows], ga[:, rows], "d loss / d actions from the persistent tape vs the chained tape")
    assert e < 2e-2 and torch.isfinite(gb).all()
    # the backward rollout as ONE persistent kernel (rlsb_rollout_bwd) on either tape vs the chained rlsb_imagine_bwd
    from rl_sandbox_b200 import _lib
    lib = _lib.load()
    for tag, o, want in (("chained tape", a, ga), ("persistent tape", b, gb)):
        l0 = lib.rlsb_launch_count(0)
        got = eng.backward(o, g_r, g_v, persistent=True).clone()
        launches = lib.rlsb_launch_count(0) - l0
        assert eng.last_backward_persistent and launches <= 4, launches
        e = rel_rms(got, want, f"d loss / d actions, persistent backward vs chained backward ({tag})")
        assert e < 1e-2 and torch.isfinite(got).all()
    # reproducible
    assert torch.equal(eng.backward(b, g_r, g_v, persistent=True), eng.backward(b, g_r, g_v, persistent=True))


def test_persistent_training_rollout_skips_heads_of_last_state(ops, cuda):
    c = load_case("c1_long")
    m = c["meta"]
    H = 5
    eng = engine(ops, m, cuda, c, H, "persistent")
    args = (c["h0"].to(cuda), c["z0"].to(cuda), None, c["lat"][:H].to(cuda), c["act"][:H].to(cuda))
    full = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in eng.rollout(*args, horizon=H).items()}
    lean = eng.rollout(*args, horizon=H, last_step_value_only=True)
    for k in ("determ", "logits", "stoch_idx", "actions", "values"):
        assert torch.equal(lean[k], full[k]), k
    for k in ("rewards", "discounts"):
        assert torch.equal(lean[k][:H], full[k][:H]), k
    assert not lean["rewards"][H].any() and bool((lean["discounts"][H] == 1).all())


def test_persistent_rollout_is_reproducible_and_cluster_size_is_reported(ops, cuda):
    from rl_sandbox_b200 import _lib
    assert _lib.load().rlsb_rollout_cluster_size() in (4, 8, 16)
    c = load_case("c1_long")
    m = c["meta"]
    H = 4
    eng = engine(ops, m, cuda, c, H, "persistent")
    h0, z0 = c["h0"].to(cuda), c["z0"].to(cuda)
    a = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in eng.rollout(h0, z0, None, None, None, seed=7, horizon=H).items()}
    b = eng.rollout(h0, z0, None, None, None, seed=7, horizon=H)
    for k in ("determ", "logits", "stoch_idx", "actions", "rewards", "discounts", "values"):
        assert torch.equal(a[k], b[k]), k
    d = eng.rollout(h0, z0, None, None, None, seed=8, horizon=H)
    assert not torch.equal(a["stoch_idx"], d["stoch_idx"])
