"""GPU parity at the north-star tolerance: latents, lambda-returns and losses within rtol 1e-3 of the REFERENCE's own
fp32 PyTorch path (tests/golden, produced by running the unmodified reference), with the rollout in the split-operand
contraction mode (``ImagineConfig(parity=True)``: every Linear is one tcgen05 contraction over [hi.Whi | hi.Wlo | lo.Whi]).

What is compared, free-running over all H steps from the fixture's start states and noise:
  * categorical indices / actions: identical to the reference's (a row stays "alive" while its whole history is);
  * determ, logits, rewards, values on alive rows: rel-RMS <= 1e-3 AND element-wise |x - ref| <= 1e-3 |ref| + 5e-4 rms(ref);
  * lambda-returns (K2 on the kernel's rewards / values / discounts): same bound; cumprod weights: exact;
  * critic / actor losses: rlsb_ac_losses (the loss kernel of K4) on head outputs evaluated by the rollout kernels in
    parity mode, against the reference's scalars when every row is alive, and ALWAYS against the oracle port (pinned to
    the reference by tests/test_oracle.py) re-run on exactly the alive rows.
No skips, no guards: a fixture row that diverges (a draw decided by less than the arithmetic's 1e-5) is removed from
both sides of every comparison, and the fraction of alive rows is asserted.
"""
import pytest
import torch

from oracle import oracle_port as orc
from tests._golden import load_case

pytestmark = pytest.mark.gpu

RTOL = 1e-3          # BASELINE.json north star
ALL = ["c1", "c2", "c2_ln", "c2_long", "c1_long"]


def rel_rms(x, r):
    x, r = x.double().cpu(), r.double().cpu()
    return ((x - r).pow(2).mean().sqrt() / r.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def assert_close(x, r, tag):
    x, r = x.double().cpu(), r.double().cpu()
    rms = r.pow(2).mean().sqrt().item()
    e = rel_rms(x, r)
    worst = ((x - r).abs() / (RTOL * r.abs() + 5e-4 * rms)).max().item()
    print(f"[parity-mode] {tag}: rel-RMS {e:.3e}, worst element at {worst:.3f} of the bound, {r.numel()} values")
    assert e <= RTOL, f"{tag}: rel-RMS {e:.3e} > {RTOL}"
    assert worst <= 1.0, f"{tag}: an element misses |x - ref| <= 1e-3 |ref| + 5e-4 rms by a factor {worst:.2f}"
    return e


def parity_engine(ops, m, c, H, cuda, target_prefix="target_critic."):
    cfg = ops.ImagineConfig(D=m["D"], A=m["A"], discrete=m["discrete"], layer_norm=m["layer_norm"],
                            predict_discount=m["predict_discount"], H=H, parity=True)
    eng = ops.ImaginationEngine(cfg)
    to = lambda sd: {k: v.to(cuda) for k, v in sd.items()}
    eng.pack(to(c["wm"]), to(c["actor"]), to(c["critic"]), target_prefix=target_prefix)
    return eng, cfg


@pytest.fixture(scope="module")
def ops(cuda):
    from rl_sandbox_b200 import ops as _ops
    return _ops


def run_parity(ops, cuda, name):
    c = load_case(name)
    m, gold = c["meta"], c["gold"]
    H, N = m["H"], m["N"]
    eng, cfg = parity_engine(ops, m, c, H, cuda)
    out = eng.rollout(c["h0"].to(cuda), c["z0"].to(cuda), None, c["lat"].to(cuda), c["act"].to(cuda), want_actor_raw=True)
    torch.cuda.synchronize()
    same = (out["stoch_idx"].cpu().long() == gold["stoch_idx"].long()).all(-1)        # (H+1, N)
    if m["discrete"]:
        same &= out["actions"].cpu().argmax(-1) == gold["actions"].argmax(-1)
    alive = same.cumprod(0).bool()
    return c, m, gold, eng, cfg, out, alive


@pytest.mark.parametrize("name", ALL)
def test_rollout_and_lambda_returns_within_1e3_of_reference(ops, cuda, name):
    c, m, gold, eng, cfg, out, alive = run_parity(ops, cuda, name)
    H, N = m["H"], m["N"]
    frac = alive[-1].float().mean().item()
    print(f"[parity-mode] {name}: rows whose draws equal the reference's over all {H} steps: {frac:.4f} ({N} rows)")
    assert frac >= 0.95, f"{name}: only {frac:.3f} of the trajectories follow the reference's draws"
    # indices are the sampler's on the kernel's own logits (bit-exact, always)
    own = orc.sample_categorical(out["logits"][1:].cpu().view(H, N, 32, 32), c["lat"].view(H, N, 32, 32))
    assert torch.equal(own, out["stoch_idx"][1:].cpu().long())
    for k in ("determ", "logits"):
        R = gold[k].shape[1]                                   # large tensors may cover the first rows only
        a = alive[:, :R]
        assert_close(out[k].cpu()[:, :R][a], gold[k][a], f"{name}.{k}")
    for k in ("rewards", "values"):
        assert_close(out[k].cpu()[alive], gold[k][alive], f"{name}.{k}")
    if not m["discrete"]:
        assert_close(out["actions"].cpu()[alive], gold["actions"][alive], f"{name}.actions")
    d_ours, d_ref = out["discounts"].cpu()[alive], gold["discounts"][alive]
    assert torch.equal(torch.nan_to_num(d_ours, nan=-1.0), torch.nan_to_num(d_ref, nan=-1.0)), "discount modes differ"
    # K2 on the kernel's outputs; a lambda-return looks H steps ahead, so compare rows alive to the end
    vs, w, adv = ops.lambda_return(out["rewards"], out["values"], out["discounts"], m["lam"])
    rows = alive[-1]
    assert_close(vs.cpu()[:, rows], gold["vs"][:, rows], f"{name}.lambda_returns")
    assert torch.equal(w.cpu()[:, rows], gold["w"][:, rows]), "cumprod weights differ"


@pytest.mark.parametrize("name", ALL)
def test_losses_within_1e3_of_reference(ops, cuda, name):
    """ImaginativeCritic.calculate_loss / ImaginativeActor.calculate_loss (ac.py:68-81,113-146) from kernel outputs only:
    actor outputs from the parity rollout, online-critic values from the parity head kernels run on the rollout's states,
    lambda-returns / weights from K2, the loss formulas by rlsb_ac_losses."""
    from rl_sandbox_b200 import _lib
    c, m, gold, eng, cfg, out, alive = run_parity(ops, cuda, name)
    H, N, D = m["H"], m["N"], m["D"]
    rows = alive[-1]
    sub = rows.nonzero().flatten()
    n = sub.numel()
    assert n >= 0.95 * N
    # online critic (critic.*, not target_critic.*) on states 0..H-1: heads of a one-step call on rows (t, n)
    eng_c, _ = parity_engine(ops, m, c, 1, cuda, target_prefix="critic.")
    flat_h = out["determ"][:H].reshape(H * N, D)
    flat_z = out["stoch"][:H].reshape(H * N, 1024)
    half = torch.full((1, H * N, 1024), 0.5, device=cuda)
    noise_a = torch.full((1, H * N, m["A"]), 0.5, device=cuda) if m["discrete"] else torch.zeros(1, H * N, m["A"], device=cuda)
    heads = eng_c.rollout(flat_h, flat_z, None, half, noise_a, horizon=1)
    critic_values = heads["values"][0].view(H, N)
    vs, w, _ = ops.lambda_return(out["rewards"], out["values"], out["discounts"], m["lam"])
    ac = ops.ACUpdateEngine(cfg, rho=m["rho"], eta=m["entropy_scale"], metrics_samples=0)
    dsub = sub.to(cuda)
    pick = lambda t: t.index_select(1, dsub).contiguous()
    scal = ac.losses_from_heads(pick(out["actor_raw"]), pick(critic_values), pick(vs), pick(w), pick(out["values"]),
                                pick(out["actions"]), horizon=H).cpu()
    torch.cuda.synchronize()
    idx = _lib.AC_SCALAR_NAMES
    keys = ("loss_critic", "loss_actor", "loss_actor_reinforce", "loss_actor_dynamics_backprop", "loss_actor_entropy")
    # the oracle port on exactly these rows (pinned to the reference on the whole fixture by tests/test_oracle.py)
    ref_traj = orc.imagine(c["wm"], c["actor"], c["critic"], c["h0"][sub], c["z0"][sub], H=H, A=m["A"],
                           discrete=m["discrete"], predict_discount=m["predict_discount"],
                           latent_uniforms=c["lat"][:, sub], action_noise=c["act"][:, sub])
    assert torch.equal(ref_traj["stoch_idx"], out["stoch_idx"].cpu().long()[:, sub]), "alive rows must share every draw"
    ref = orc.ac_losses(ref_traj, c["actor"], c["critic"], lam=m["lam"], discrete=m["discrete"], rho=m["rho"],
                        eta=m["entropy_scale"])
    for k in keys:
        got, want = scal[idx[k]].item(), float(ref[k])
        print(f"[parity-mode] {name}.{k} on {n}/{N} rows: kernels {got:.7f} oracle {want:.7f}"
              + (f" reference {gold[k].item():.7f}" if n == N else ""))
        assert abs(got - want) <= RTOL * abs(want) + 1e-6, (name, k, got, want)
        if n == N:   # every row alive: the reference's own scalar
            assert abs(got - gold[k].item()) <= RTOL * abs(gold[k].item()) + 1e-6, (name, k, got, gold[k].item())
    for k, g in (("critic/avg_lambda_value", "critic_avg_lambda_value"), ("critic/avg_predicted_value", "critic_avg_predicted_value"),
                 ("critic/avg_target_value", "critic_avg_target_value")):
        if n == N:
            got, want = scal[idx[k]].item(), gold[g].item()
            assert abs(got - want) <= RTOL * abs(want) + 1e-5, (name, k, got, want)


def test_parity_mode_matches_fast_mode_layouts_and_philox(ops, cuda):
    """Same outputs, shapes and noise streams as the fast path: Philox mode is shard-invariant, the packed state images
    kept for the update are those of the fast path up to the arithmetic (bf16-rounded fp32 states)."""
    c = load_case("c2_long")
    m = c["meta"]
    H, N = 4, m["N"]
    eng, cfg = parity_engine(ops, m, c, H, cuda)
    h0, z0 = c["h0"].to(cuda), c["z0"].to(cuda)
    full = eng.rollout(h0, z0, None, None, None, seed=5, row_offset=0, horizon=H, keep_packed=True, want_stoch=True)
    full = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in full.items()}
    a, b = 7, 31
    part = eng.rollout(h0[a:b].contiguous(), z0[a:b].contiguous(), None, None, None, seed=5, row_offset=a, horizon=H)
    for k in ("determ", "logits", "stoch_idx", "actions", "rewards", "discounts", "values"):
        assert torch.equal(part[k], full[k][:, a:b]), k
    # the kept hi images are bf16(determ) / the one-hot
    hp = ops.unpack_rows(full["determ_packed"][1].flatten(), N, m["D"])
    assert torch.equal(hp, full["determ"][1].bfloat16().float())
    zp = ops.unpack_rows(full["stoch_packed"][2].flatten(), N, 1024)
    assert torch.equal(zp, full["stoch"][2])
    # fast mode on the same Philox key: same draws on (nearly) all rows, values within the bf16 band
    fast_cfg = ops.ImagineConfig(D=m["D"], A=m["A"], discrete=True, layer_norm=True, predict_discount=True, H=H)
    fast = ops.ImaginationEngine(fast_cfg)
    to = lambda sd: {k: v.to(cuda) for k, v in sd.items()}
    fast.pack(to(c["wm"]), to(c["actor"]), to(c["critic"]))
    fo = fast.rollout(h0, z0, None, None, None, seed=5, row_offset=0, horizon=H)
    same = (fo["stoch_idx"] == full["stoch_idx"]).all(-1).cumprod(0).bool()
    assert same[1].float().mean() > 0.9
    assert rel_rms(fo["determ"][1][same[1]], full["determ"][1][same[1]]) < 3e-3


def test_parity_mode_rejects_unsupported_requests(ops, cuda):
    from rl_sandbox_b200 import _lib
    with pytest.raises(_lib.RlsbError):
        ops.ImaginationEngine(ops.ImagineConfig(D=200, A=1, discrete=False, layer_norm=True, predict_discount=False,
                                                slots=4, parity=True))
    with pytest.raises(_lib.RlsbError):
        ops.ImaginationEngine(ops.ImagineConfig(D=200, A=12, discrete=False, layer_norm=False, predict_discount=False,
                                                with_backward=True, parity=True))
