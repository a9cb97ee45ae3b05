"""GPU parity of K4 (rlsb_ac_update): critic / actor losses and parameter gradients computed by the
tcgen05 kernel chain vs torch autograd of the mirror modules (themselves verified identical to the
reference's ImaginativeCritic / ImaginativeActor.calculate_loss on CPU, oracle/check_host_mirror.py)
on the same K1 rollout.  Also the weight-gradient contraction alone against a matmul."""
import copy

import pytest
import torch

from tests._golden import load_case
from tests.test_gpu_agent import load_params, make_agent

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,n,k", [(128, 64, 64), (1000, 448, 512), (4096, 400, 2112), (700, 17, 448)])
def test_wgrad_contraction(cuda, M, n, k):
    """dW = dY^T X with both operands read as MN-major UMMA operands from the packed row images."""
    from rl_sandbox_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + n + k)
    dy = torch.randn(M, n, device="cuda", generator=g)
    x = torch.randn(M, k, device="cuda", generator=g)
    n_pad, k_pad = ops.round_up(n, 64), ops.round_up(k, 64)
    out = ops.gemm_wgrad(ops.pack_rows(dy, k_pad=n_pad), n_pad, ops.pack_rows(x, k_pad=k_pad), k_pad, M)
    ref = dy.bfloat16().double().t() @ x.bfloat16().double()
    err = (out[:n, :k].double() - ref).abs().max().item()
    assert err < 2e-5 * M ** 0.5 * 16, err           # fp32 accumulation of exact bf16 products
    assert out[n:].abs().max().item() == 0 if n < n_pad else True
    assert (out[:, k:] == 0).all()


def _reference_grads(agent, k1, vs, w, metrics_samples=128):
    """critic.calculate_loss / actor.calculate_loss + backward with torch autograd (the reference's op sequence)."""
    zs = torch.cat([k1["determ"], torch.nn.functional.one_hot(k1["stoch_idx"].long(), 32).float().flatten(-2)], -1)
    values = k1["values"].unsqueeze(-1)
    actions = k1["actions"]
    vs3, w3 = vs.unsqueeze(-1), w.unsqueeze(-1)
    for p in list(agent.actor.parameters()) + list(agent.critic.parameters()):
        p.grad = None
    lc, mc = agent.critic.calculate_loss(zs[:-1], vs3, w3[:-1], target_values=values[:-1])
    la, ma = agent.actor.calculate_loss(zs[:-2], vs3[1:], values[:-2], w3[:-2], actions[1:-1], metrics_samples=metrics_samples)
    lc["loss_critic"].backward()
    la["loss_actor"].backward()
    grads = {"actor." + n: p.grad.clone() for n, p in agent.actor.actor.named_parameters()}
    grads |= {"critic." + n: p.grad.clone() for n, p in agent.critic.critic.named_parameters()}
    return lc | la, mc | ma, grads


@pytest.mark.parametrize("name,N", [("c1", None), ("c1", 1000)])
def test_ac_update_matches_autograd(cuda, name, N):
    from rl_sandbox.agents.dreamer.rssm import State
    from rl_sandbox_b200 import _lib, ops
    c = load_case(name)
    m = c["meta"]
    H = m["H"]
    agent = make_agent(m, "cuda")
    load_params(agent, c)
    torch.manual_seed(5)
    # the golden case has untrained (near-zero output) heads: perturb the actor / critic so the losses have signal
    with torch.no_grad():
        for p in list(agent.actor.parameters()) + list(agent.critic.critic.parameters()):
            p.add_(0.05 * torch.randn_like(p))
    agent.mark_weights_changed()
    if N is None:
        N = m["N"]
        h0, z0 = c["h0"].cuda(), c["z0"].cuda()
    else:
        g = torch.Generator(device="cuda").manual_seed(3)
        h0 = 0.5 * torch.randn(N, m["D"], device="cuda", generator=g)
        z0 = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device="cuda", generator=g), 32).float().view(N, 1024)
    init = State(h0.unsqueeze(0), torch.zeros(1, N, 32, 32, device="cuda"), z0.unsqueeze(0))
    with torch.no_grad():
        agent.imagine_trajectory(init, noise={"seed": 11}, keep_packed=True)
    k1 = agent.last_rollout
    assert k1["determ_packed"] is not None
    # random (but valid) discounts / rewards so every loss term is exercised
    k1["discounts"] = (torch.rand_like(k1["discounts"]) > 0.1).float()
    k1["discounts"][0] = 1.0
    k1["rewards"] = torch.randn_like(k1["rewards"])
    vs, w, _ = ops.lambda_return(k1["rewards"], k1["values"], k1["discounts"], agent.critic.lambda_)

    ref_losses, ref_metrics, ref_grads = _reference_grads(agent, k1, vs, w)
    for p in list(agent.actor.parameters()) + list(agent.critic.parameters()):
        p.grad = None
    eng = agent._get_ac_engine()
    scal = eng.update(k1, vs, w, agent.actor.actor, agent.critic.critic, seed=1, horizon=H).cpu()
    torch.cuda.synchronize()
    idx = _lib.AC_SCALAR_NAMES
    for k in ("loss_critic", "loss_actor_reinforce", "loss_actor_entropy", "loss_actor"):
        got, ref = scal[idx[k]].item(), ref_losses[k].item()
        print(f"[parity] K4 {name} N={N} {k}: ours {got:.6f} torch {ref:.6f}")
        # same tolerance as test_losses_match_reference: the reinforce term is a signed sum (advantages cancel)
        assert abs(got - ref) <= 5e-3 * abs(ref) + 2e-4, (k, got, ref)
    for k in ("critic/avg_target_value", "critic/avg_lambda_value", "critic/avg_predicted_value", "actor/mean_val",
              "actor/avg_val", "actor/min_val", "actor/max_val"):
        got, ref = scal[idx[k]].item(), ref_metrics[k].item()
        assert abs(got - ref) <= 5e-3 * abs(ref) + 5e-3, (k, got, ref)   # means of O(1) bf16-contraction outputs
    got, ref = scal[idx["actor/avg_sd"]].item(), ref_metrics["actor/avg_sd"].item()
    assert abs(got - ref) <= 2e-2 * abs(ref), ("actor/avg_sd", got, ref)   # different random draws, same statistic
    ours = {"actor." + n: p.grad for n, p in agent.actor.actor.named_parameters()}
    ours |= {"critic." + n: p.grad for n, p in agent.critic.critic.named_parameters()}
    worst = 0.0
    for n, gref in ref_grads.items():
        g = ours[n]
        assert g is not None and g.shape == gref.shape, n
        rel = ((g - gref).norm() / gref.norm().clamp_min(1e-12)).item()
        cos = torch.nn.functional.cosine_similarity(g.flatten(), gref.flatten(), dim=0).item()
        print(f"[parity] K4 grad {n}: rel-L2 {rel:.3e} cos {cos:.6f} |g| {gref.norm().item():.3e}")
        worst = max(worst, rel)
        assert rel < 3e-2 and cos > 0.999, (n, rel, cos)
    print(f"[parity] K4 {name} N={N}: worst gradient rel-L2 error {worst:.3e}")


def test_fused_behaviour_update_trains_like_autograd(cuda):
    """Two agents with identical parameters and noise: one fused (K4) update vs one torch-autograd update give
    the same parameters after AdamW to within the bf16 gradient error."""
    from rl_sandbox.agents.dreamer.rssm import State
    c = load_case("c1")
    m = c["meta"]
    N = m["N"]
    a1 = make_agent(m, "cuda")
    load_params(a1, c)
    with torch.no_grad():
        torch.manual_seed(2)
        for p in list(a1.actor.parameters()) + list(a1.critic.critic.parameters()):
            p.add_(0.05 * torch.randn_like(p))
    a2 = make_agent(m, "cuda")
    a2.world_model.load_state_dict(a1.world_model.state_dict())
    a2.actor.load_state_dict(a1.actor.state_dict())
    a2.critic.load_state_dict(a1.critic.state_dict())
    a1.mark_weights_changed(); a2.mark_weights_changed()
    a2.fused_ac_update = False
    before = copy.deepcopy(a1.actor.state_dict())
    init = State(c["h0"].unsqueeze(0).cuda(), torch.zeros(1, N, 32, 32, device="cuda"), c["z0"].unsqueeze(0).cuda())
    l1, m1 = a1.behaviour_update(init, noise={"seed": 3})
    l2, m2 = a2.behaviour_update(init, noise={"seed": 3})
    for k in ("loss_actor", "loss_critic"):
        assert abs(l1[k].item() - l2[k].item()) <= 2e-3 * abs(l2[k].item()) + 1e-5, (k, l1[k].item(), l2[k].item())
    moved = agree = 0.0
    for (n, p1), (_, p2) in zip(list(a1.actor.named_parameters()) + list(a1.critic.critic.named_parameters()),
                                list(a2.actor.named_parameters()) + list(a2.critic.critic.named_parameters())):
        moved += (p2 - before.get(n, p2)).abs().sum().item() if n in before else 0.0
        agree += (p1 - p2).abs().sum().item()
    # AdamW's first step moves every weight by ~lr regardless of gradient scale; sign agreement dominates
    same_sign = []
    for (n, p1), (_, p2) in zip(a1.actor.named_parameters(), a2.actor.named_parameters()):
        d1, d2 = p1 - before[n], p2 - before[n]
        same_sign.append(((d1 * d2) > 0).float().mean().item())
    print(f"[parity] fused vs autograd AdamW step: mean fraction of weights moving the same way {sum(same_sign)/len(same_sign):.4f}")
    assert sum(same_sign) / len(same_sign) > 0.97


def _engine_for(m, H, with_backward):
    from rl_sandbox_b200 import ops
    cfg = ops.ImagineConfig(D=m["D"], A=m["A"], discrete=m["discrete"], layer_norm=m["layer_norm"],
                            predict_discount=m["predict_discount"], H=H, with_backward=with_backward)
    return ops.ImaginationEngine(cfg), cfg


def _modules_for(m, actor_sd, critic_sd):
    from rl_sandbox.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
    actor = ImaginativeActor(latent_dim=m["D"] + 1024, actions_num=m["A"], is_discrete=m["discrete"],
                             layer_norm=m["layer_norm"], reinforce_fraction=None, entropy_scale=m["entropy_scale"]).cuda()
    critic = ImaginativeCritic(discount_factor=m["gamma"], update_interval=100, soft_update_fraction=1,
                               value_target_lambda=m["lam"], latent_dim=m["D"] + 1024, layer_norm=m["layer_norm"]).cuda()
    actor.load_state_dict(actor_sd)
    critic.load_state_dict(critic_sd)
    return actor, critic


def _fused_update(m, wm, actor_sd, critic_sd, h0, z0, lat, act, H):
    """K1 (tape) -> K2 -> K2 bwd -> K1 bwd -> K4 exactly as DreamerV2._behaviour_update_fused strings them."""
    from rl_sandbox_b200 import ops
    dev = "cuda"
    to = lambda sd: {k: v.to(dev) for k, v in sd.items()}
    dyn = not m["discrete"]
    eng, cfg = _engine_for(m, H, with_backward=dyn)
    eng.pack(to(wm), to(actor_sd), to(critic_sd))
    k1 = eng.rollout(h0.to(dev), z0.to(dev), None, lat.to(dev), act.to(dev), keep_packed=True, tape=dyn, want_stoch=False)
    vs, w, _ = ops.lambda_return(k1["rewards"], k1["values"], k1["discounts"], m["lam"])
    n = h0.shape[0]
    g_actions = None
    if dyn:
        g_vs = torch.zeros_like(vs)
        g_vs[1:] = w[:H - 1] * (-(1.0 - m["rho"]) / ((H - 1) * n))
        g_r, g_v, _ = ops.lambda_return_bwd(g_vs, k1["values"], k1["discounts"], vs, m["lam"])
        g_actions = eng.backward(k1, g_r, g_v)
    actor, critic = _modules_for(m, actor_sd, critic_sd)
    ac = ops.ACUpdateEngine(cfg, rho=m["rho"], eta=m["entropy_scale"], metrics_samples=128)
    ac.pack(actor.state_dict(), critic.state_dict())
    scal = ac.update(k1, vs, w, actor.actor, critic.critic, seed=3, horizon=H, g_actions=g_actions).cpu()
    grads = {"actor." + k: p.grad for k, p in actor.actor.named_parameters()}
    grads |= {"critic." + k: p.grad for k, p in critic.critic.named_parameters()}
    torch.cuda.synchronize()
    return k1, vs, g_actions, scal, grads


def _teacher_forced_update(m, c, gold, H):
    """K4 on the REFERENCE's own trajectory: the state images are packed from the fixture's determ / stoch, the
    lambda-returns, weights, values and actions are the reference's — nothing depends on the rollout's draws."""
    from rl_sandbox_b200 import ops
    dev = "cuda"
    N = m["N"]
    rows = ops.round_up(N, 128)
    z = torch.nn.functional.one_hot(gold["stoch_idx"].long(), 32).float().reshape(H + 1, N, 1024)
    k1 = {"determ": gold["determ"].to(dev), "values": gold["values"].to(dev), "actions": gold["actions"].to(dev),
          "determ_packed": torch.stack([ops.pack_rows(gold["determ"][t].to(dev), rows_pad=rows) for t in range(H + 1)]),
          "stoch_packed": torch.stack([ops.pack_rows(z[t].to(dev), rows_pad=rows) for t in range(H + 1)])}
    _, cfg = _engine_for(m, H, with_backward=False)
    actor, critic = _modules_for(m, c["actor"], c["critic"])
    ac = ops.ACUpdateEngine(cfg, rho=m["rho"], eta=m["entropy_scale"], metrics_samples=128)
    ac.pack(actor.state_dict(), critic.state_dict())
    scal = ac.update(k1, gold["vs"].to(dev), gold["w"].to(dev), actor.actor, critic.critic, seed=3, horizon=H).cpu()
    grads = {"actor." + k: p.grad for k, p in actor.actor.named_parameters()}
    grads |= {"critic." + k: p.grad for k, p in critic.critic.named_parameters()}
    torch.cuda.synchronize()
    return scal, grads


def _parity_rollout_update(m, c, H):
    """K4 on the states of a rollout in the split-operand mode (within ~1e-5 of the reference's trajectory)."""
    from rl_sandbox_b200 import ops
    dev = "cuda"
    to = lambda sd: {k: v.to(dev) for k, v in sd.items()}
    cfg = ops.ImagineConfig(D=m["D"], A=m["A"], discrete=True, layer_norm=m["layer_norm"],
                            predict_discount=m["predict_discount"], H=H, parity=True)
    eng = ops.ImaginationEngine(cfg)
    eng.pack(to(c["wm"]), to(c["actor"]), to(c["critic"]))
    k1 = eng.rollout(c["h0"].to(dev), c["z0"].to(dev), None, c["lat"].to(dev), c["act"].to(dev), keep_packed=True,
                     want_stoch=False)
    vs, w, _ = ops.lambda_return(k1["rewards"], k1["values"], k1["discounts"], m["lam"])
    actor, critic = _modules_for(m, c["actor"], c["critic"])
    ac = ops.ACUpdateEngine(cfg, rho=m["rho"], eta=m["entropy_scale"], metrics_samples=128)
    ac.pack(actor.state_dict(), critic.state_dict())
    scal = ac.update(k1, vs, w, actor.actor, critic.critic, seed=3, horizon=H).cpu()
    grads = {"actor." + k: p.grad for k, p in actor.actor.named_parameters()}
    grads |= {"critic." + k: p.grad for k, p in critic.critic.named_parameters()}
    torch.cuda.synchronize()
    return k1, scal, grads


@pytest.mark.parametrize("name", ["c1", "c2_long", "c1_long", "c2", "c2_ln"])
def test_fused_update_matches_reference_gradients(cuda, name):
    """The fused update vs the gradients the REFERENCE's autograd produced on the same parameters, start states and
    noise (tests/golden: norms + 64 probed entries per tensor; d loss_actor / d a_t for the continuous cases).

    Discrete actors (rho = 1: nothing differentiates through the rollout): K4 is driven from the reference's own
    trajectory (teacher-forced; c1_long, whose fixture stores the states of 40 of its 128 rows, from a split-operand
    rollout that reproduces the reference's trajectory), so a draw the bf16 rollout would flip cannot make the
    comparison vacuous.  Continuous actors: the whole chain K1 (tape) -> K2 -> K2 bwd -> K1 bwd -> K4; their fixtures
    (6 rows x 3 steps) are reproduced draw for draw, which is asserted."""
    from oracle.gen_golden import grad_probe_indices
    from rl_sandbox_b200 import _lib
    c = load_case(name)
    m, gold = c["meta"], c["gold"]
    H, N = m["H"], m["N"]
    g_actions = None
    tol_rows = 0
    if m["discrete"] and gold["determ"].shape[1] == N:
        scal, grads = _teacher_forced_update(m, c, gold, H)
    elif m["discrete"]:
        k1, scal, grads = _parity_rollout_update(m, c, H)
        same = ((k1["stoch_idx"].cpu() == gold["stoch_idx"]).all(-1) &
                (k1["actions"].argmax(-1).cpu() == gold["actions"].argmax(-1))).all(0)
        tol_rows = int((~same).sum())
        print(f"[parity] {name}: rows of the split-operand rollout that left the reference's trajectory: {tol_rows} of {N}")
        assert tol_rows <= 2
    else:
        k1, vs, g_actions, scal, grads = _fused_update(m, c["wm"], c["actor"], c["critic"], c["h0"], c["z0"], c["lat"],
                                                       c["act"], H)
        same = (k1["stoch_idx"].cpu() == gold["stoch_idx"]).all(-1)
        assert bool(same.all()), f"{name}: the rollout no longer reproduces the fixture's draws: regenerate it with another seed"
    idx = _lib.AC_SCALAR_NAMES
    for k in ("loss_critic", "loss_actor", "loss_actor_dynamics_backprop", "loss_actor_entropy"):
        got, ref = scal[idx[k]].item(), gold[k].item()
        print(f"[parity] {name}.{k}: fused {got:.6f} reference {ref:.6f}")
        # means over H x N head outputs (18 for the small fixtures), each a 5-deep bf16 contraction chain (4-9e-3 per
        # element); a diverged row moves a mean by at most its share
        assert abs(got - ref) <= (1e-2 + 2.0 * tol_rows / N) * abs(ref) + 2e-4, (k, got, ref)
    if not m["discrete"]:
        ga, gref = g_actions.cpu(), gold["grad_actions"]
        rel = ((ga - gref).norm() / gref.norm()).item()
        print(f"[parity] {name}: d loss / d actions rel-L2 vs reference {rel:.3e} (|g| {gref.norm().item():.3e})")
        assert rel < 3e-2, rel
    worst = 0.0
    for i, n in enumerate(m["grad_names"]):
        g = grads[n].cpu()
        nref = gold["grad_norms"][i].item()
        probe = g.flatten()[grad_probe_indices(g.numel())]
        perr = ((probe - gold["grad_probes"][i]).norm() / gold["grad_probes"][i].norm().clamp_min(1e-12)).item()
        nerr = abs(g.norm().item() - nref) / max(nref, 1e-12)
        worst = max(worst, perr)
        assert nerr < 3e-2 + 2.0 * tol_rows / N and perr < 6e-2 + 2.0 * tol_rows / N, (n, nerr, perr)
    print(f"[parity] {name}: worst probed-gradient rel-L2 vs reference {worst:.3e}")


@pytest.mark.parametrize("layer_norm", [False, True])
def test_continuous_update_matches_oracle(cuda, layer_norm):
    """config-2 dims at N = 300 (ragged last tile), H = 5: K1 backward + continuous K4 vs the oracle's autograd
    with bf16-rounded contraction operands (same arithmetic, so the draws coincide)."""
    from oracle import oracle_port as orc
    m = dict(D=200, A=12, discrete=False, layer_norm=layer_norm, predict_discount=False, lam=0.95, rho=0.0,
             entropy_scale=1e-3, gamma=0.99)
    H, N = 5, 300
    wm, actor, critic = orc.make_params(77, D=200, A=12, discrete=False, layer_norm=layer_norm, predict_discount=False)
    h0, z0 = orc.make_start(78, N, 200)
    g = torch.Generator().manual_seed(79)
    lat, act = torch.rand(H, N, 1024, generator=g), torch.randn(H, N, 12, generator=g)
    ref = orc.continuous_update_grads(wm, actor, critic, h0, z0, lat, act, H=H, A=12, lam=0.95, rho=0.0, eta=1e-3, bf16=True)
    k1, vs, g_actions, scal, grads = _fused_update(m, wm, actor, critic, h0, z0, lat, act, H)
    same = (k1["stoch_idx"].cpu().long() == ref["traj"]["stoch_idx"]).all(-1).all(0)
    print(f"[parity] continuous ln={layer_norm}: rows with identical draws {same.float().mean().item():.4f}")
    assert same.float().mean().item() > 0.97
    ga, gref = g_actions.cpu()[:, same], ref["g_actions"][:, same]
    rel = ((ga - gref).norm() / gref.norm()).item()
    print(f"[parity] continuous ln={layer_norm}: d loss / d actions rel-L2 {rel:.3e}")
    assert rel < 3e-2, rel
    if same.all():
        for n, gr in ref["grads"].items():
            gg = grads[n].cpu()
            r = ((gg - gr).norm() / gr.norm().clamp_min(1e-12)).item()
            assert r < 5e-2, (n, r)


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_cuda_graph_replay_matches_eager(cuda, name):
    """The captured update (pack + K1 + K2 [+ bwd] + K4 in one CUDA graph, Philox key in device memory) reproduces the
    eagerly launched one bit for bit, step after step (same seeds, parameters evolving under AdamW)."""
    from rl_sandbox.agents.dreamer.rssm import State
    c = load_case(name)
    m = c["meta"]
    N = 200
    g = torch.Generator(device="cuda").manual_seed(3)
    h0 = 0.5 * torch.randn(N, m["D"], device="cuda", generator=g)
    z0 = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device="cuda", generator=g), 32).float().view(N, 1024)
    agents = []
    for graphed in (True, False):
        a = make_agent(m, "cuda", H=5)
        load_params(a, c)
        a.cuda_graph = graphed
        agents.append(a)
    assert agents[0]._can_fuse_ac()
    for step in range(3):
        outs = []
        for a in agents:
            init = State(h0.unsqueeze(0) + 0.01 * step, torch.zeros(1, N, 32, 32, device="cuda"), z0.unsqueeze(0))
            losses, _ = a.behaviour_update(init, noise={"seed": 40 + step})
            outs.append((losses, {k: v.clone() for k, v in a.last_rollout.items() if torch.is_tensor(v) and k != "tape"}))
        assert agents[0]._graphs and not agents[1]._graphs
        for k in outs[0][0]:
            torch.testing.assert_close(outs[0][0][k], outs[1][0][k], rtol=5e-2, atol=1e-4)   # a flipped draw is possible after step 0
        for (n1, p1), (_, p2) in zip(agents[0].actor.named_parameters(), agents[1].actor.named_parameters()):
            # LayerNorm gamma / beta gradients are summed with shared-memory float atomics (order-dependent rounding):
            # everything downstream of the first AdamW step agrees to rounding, not bit for bit
            if step == 0 and p1.dim() == 2:
                assert torch.equal(p1, p2), (step, n1)
            torch.testing.assert_close(p1, p2, rtol=1e-3, atol=2e-4, msg=lambda s_: f"{step} {n1}: {s_}")   # AdamW moves each weight by <= lr per step
        if step == 0:
            for k in outs[0][0]:
                assert torch.equal(outs[0][0][k], outs[1][0][k]), (step, k)
            for k in ("determ", "stoch_idx", "actions", "rewards", "values"):
                assert torch.equal(outs[0][1][k], outs[1][1][k]), (step, k)
    print(f"[parity] {name}: graph replays track eager steps (step 0 bit-identical; later steps to fp32 rounding)")


@pytest.mark.parametrize("graphed", [False, True])
def test_actor_forward_reuse_matches_recompute(cuda, graphed):
    """K1 leaves the actor head's activations (layer outputs, x_hat, 1/std of steps 0..H-1) in the update's workspace
    (rlsb_imagine_out::actor_slots) and K4 skips the actor's forward: same kernel, same operands -> the losses and the
    weight / bias gradients are bit-identical to the update that recomputes the actor forward itself."""
    from rl_sandbox.agents.dreamer.rssm import State
    c = load_case("c1")
    m = c["meta"]
    N = 300   # not a multiple of 128: padding rows of the last tile stay zero in the reused images too
    g = torch.Generator(device="cuda").manual_seed(5)
    h0 = 0.5 * torch.randn(N, m["D"], device="cuda", generator=g)
    z0 = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device="cuda", generator=g), 32).float().view(N, 1024)
    res = []
    for reuse in (True, False):
        a = make_agent(m, "cuda", H=5)
        load_params(a, c)
        a.cuda_graph = graphed
        a.reuse_actor_forward = reuse
        a.reuse_actor_min_rows = 0
        a._get_engine().persistent_max_rows = 0   # both runs through the chained rollout (the persistent kernel has no actor slots)
        init = State(h0.unsqueeze(0), torch.zeros(1, N, 32, 32, device="cuda"), z0.unsqueeze(0))
        from rl_sandbox_b200 import _lib
        before = _lib.load().rlsb_launch_count(0)
        losses, _ = a.behaviour_update(init, noise={"seed": 91})
        launches = _lib.load().rlsb_launch_count(0) - before
        grads = {n: p.grad.clone() for n, p in list(a.actor.named_parameters()) + list(a.critic.critic.named_parameters())}
        res.append((losses, grads, launches))
    for k in res[0][0]:
        assert torch.equal(res[0][0][k], res[1][0][k]), k
    for n, ga in res[0][1].items():
        gb = res[1][1][n]
        if ga.dim() == 2:
            assert torch.equal(ga, gb), n
        else:   # biases through the ones tile are deterministic too; LayerNorm gamma / beta use smem float atomics
            torch.testing.assert_close(ga, gb, rtol=1e-4, atol=1e-7, msg=lambda s_: f"{n}: {s_}")
    print(f"[parity] actor forward reuse (graphed={graphed}): losses and weight gradients bit-identical")


def test_ac_update_full_size_properties(cuda):
    """BASELINE sweep size (32768 start states x H = 15), size-independent properties of K4:
    weight / bias gradients are bitwise reproducible run to run (fixed-order split reductions), everything is finite,
    and the actor's last-layer bias gradient sums to zero (sum_k d loss / d logit_k = 0 for both the reinforce and
    the entropy term of a softmax policy)."""
    from rl_sandbox.agents.dreamer.rssm import State
    c = load_case("c1")
    m = c["meta"]
    agent = make_agent(m, "cuda", H=15)
    load_params(agent, c)
    torch.manual_seed(1)
    with torch.no_grad():
        for p in list(agent.actor.parameters()) + list(agent.critic.critic.parameters()):
            p.add_(0.05 * torch.randn_like(p))
    agent.mark_weights_changed()
    N = 32768
    g = torch.Generator(device="cuda").manual_seed(3)
    h0 = 0.5 * torch.randn(N, m["D"], device="cuda", generator=g)
    z0 = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device="cuda", generator=g), 32).float().view(N, 1024)
    init = State(h0.unsqueeze(0), torch.zeros(1, N, 32, 32, device="cuda"), z0.unsqueeze(0))
    from rl_sandbox_b200 import ops
    with torch.no_grad():
        agent.imagine_trajectory(init, noise={"seed": 5}, keep_packed=True)
    k1 = agent.last_rollout
    vs, w, _ = ops.lambda_return(k1["rewards"], k1["values"], k1["discounts"], agent.critic.lambda_)
    eng = agent._get_ac_engine()
    runs = []
    for _ in range(2):
        for p in list(agent.actor.parameters()) + list(agent.critic.parameters()):
            p.grad = None
        scal = eng.update(k1, vs, w, agent.actor.actor, agent.critic.critic, seed=1, horizon=15).clone()
        runs.append((scal, {n: p.grad.clone() for n, p in list(agent.actor.actor.named_parameters()) +
                            [("c." + n, p) for n, p in agent.critic.critic.named_parameters()]}))
    assert torch.isfinite(runs[0][0]).all()
    for n, g0 in runs[0][1].items():
        assert torch.isfinite(g0).all(), n
        if int(n.split(".")[-2]) in (0, 3, 6, 9, 12):   # nn.Linear weights / biases (LayerNorm sits at 1, 4, 7, 10)
            assert torch.equal(g0, runs[1][1][n]), f"{n}: weight / bias gradients must be bitwise reproducible"
    gb = runs[0][1]["12.bias"]
    assert gb.sum().abs().item() < 2e-2 * gb.norm().item(), (gb.sum().item(), gb.norm().item())
    print(f"[parity] K4 full size: actor last-bias gradient sum {gb.sum().item():.3e} vs norm {gb.norm().item():.3e}")


def test_fused_update_in_chunks_matches_one_pass(cuda):
    """More start states than `max_rows_per_pass`: the fused update runs in passes over contiguous chunks.  Philox counters
    are global start-state indices, so the chunked update draws the same noise and produces the same losses and parameter
    gradients as one pass (up to the summation order of the means)."""
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    import bench
    from rl_sandbox_b200.agents.dreamer.rssm import State
    dims = dict(bench.DIMS["config1"], D=256)
    N, H = 800, 6
    res = {}
    for chunked in (False, True):
        torch.manual_seed(0)
        agent = bench.build_agent(dims, H, "cuda", 16)
        agent.cuda_graph = False
        if chunked:
            agent.max_rows_per_pass = 300      # -> 3 passes of 384 / 384 / 32 start states
        g = torch.Generator(device="cuda").manual_seed(1)
        h0 = 0.5 * torch.randn(N, dims["D"], device="cuda", generator=g)
        z0 = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device="cuda", generator=g), 32).float().view(N, 1024)
        state = State(h0.unsqueeze(0), torch.zeros(1, N, 32, 32, device="cuda"), z0.unsqueeze(0))
        losses, metrics = agent.behaviour_update(state, noise={"seed": 77, "row_offset": 1000})
        grads = {n: p.grad.clone() for n, p in list(agent.actor.named_parameters()) + list(agent.critic.critic.named_parameters())}
        res[chunked] = (losses, metrics, grads)
    for k in ("loss_actor", "loss_critic", "loss_actor_entropy", "loss_actor_reinforce"):
        a, b = float(res[True][0][k]), float(res[False][0][k])
        print(f"[parity] chunked vs one pass {k}: {a:.6f} vs {b:.6f}")
        assert abs(a - b) <= 2e-4 * abs(b) + 1e-6, k
    worst = max(((res[True][2][n] - res[False][2][n]).norm() / (res[False][2][n].norm() + 1e-12)).item() for n in res[False][2])
    print(f"[parity] chunked vs one pass: worst parameter-gradient rel-L2 {worst:.3e}")
    assert worst < 2e-2   # the per-pass loss scale 1 / (H n_chunk) changes the bf16 rounding of the dY images (0.4 % per element)
