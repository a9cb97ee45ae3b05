"""GPU tests at the agent boundary (rl_sandbox.agents.DreamerV2): the reference's API driving K1 + K2."""
from functools import partial

import pytest
import torch

from oracle import oracle_port as orc
from tests._golden import load_case

pytestmark = pytest.mark.gpu


def make_agent(m, device, H=None, batch_cluster_size=50):
    # through the ALIAS package: the dotted paths the reference's Hydra configs name
    from rl_sandbox.agents import DreamerV2
    from rl_sandbox.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
    from rl_sandbox.agents.dreamer.world_model import WorldModel
    from rl_sandbox.utils.optimizer import Optimizer
    ln = m["layer_norm"]
    opt = partial(Optimizer, lr=1e-4, eps=1e-5, weight_decay=1e-6, clip=100)
    return DreamerV2(
        obs_space_num=[64, 64, 3], clip_rewards="identity", actions_num=m["A"],
        world_model=partial(WorldModel, batch_cluster_size=batch_cluster_size, latent_dim=32, latent_classes=32,
                            rssm_dim=m["D"], discount_loss_scale=1.0, kl_loss_scale=2, kl_loss_balancing=0.8,
                            kl_free_nats=1.0, discrete_rssm=False, predict_discount=m["predict_discount"], layer_norm=ln,
                            encode_vit=False, decode_vit=False, vit_l2_ratio=0.5, vit_img_size=224),
        actor=partial(ImaginativeActor, layer_norm=ln, reinforce_fraction=None, entropy_scale=m["entropy_scale"]),
        critic=partial(ImaginativeCritic, discount_factor=m["gamma"], update_interval=100, soft_update_fraction=1,
                       value_target_lambda=0.95, layer_norm=ln),
        action_type="discrete" if m["discrete"] else "continuous", imagination_horizon=H or m["H"],
        wm_optim=opt, actor_optim=opt, critic_optim=opt, layer_norm=ln, batch_cluster_size=batch_cluster_size,
        f16_precision=False, device_type=device)


def load_params(agent, c):
    agent.world_model.load_state_dict(c["wm"], strict=False)
    agent.actor.load_state_dict(c["actor"])
    agent.critic.load_state_dict(c["critic"])
    agent.mark_weights_changed()


@pytest.mark.parametrize("name", ["c1", "c2", "c2_long"])
def test_losses_match_reference(cuda, name):
    """imagine_trajectory (K1) -> lambda_return (K2) -> calculate_loss, vs the losses the REFERENCE's
    methods produced on the same parameters, start states and noise (tests/golden)."""
    from rl_sandbox.agents.dreamer.rssm import State
    from rl_sandbox_b200 import ops
    c = load_case(name)
    m, gold = c["meta"], c["gold"]
    H, N = m["H"], m["N"]
    agent = make_agent(m, "cuda")
    load_params(agent, c)
    init = State(c["h0"].unsqueeze(0).cuda(), torch.zeros(1, N, 32, 32, device="cuda"), c["z0"].unsqueeze(0).cuda())
    with torch.no_grad():
        states, actions, rewards, ts = agent.imagine_trajectory(
            init, noise={"latent_uniforms": c["lat"].cuda(), "action_noise": c["act"].cuda()})
    assert states.determ.shape == (H + 1, N, m["D"]) and states.stoch_logits.shape == (H + 1, N, 32, 32)
    assert actions.shape == (H + 1, N, m["A"]) and rewards.shape == (H + 1, N, 1) and ts.shape == (H + 1, N, 1)
    idx = states.stoch.view(H + 1, N, 32, 32).argmax(-1).cpu()
    same = (idx == gold["stoch_idx"].long()).all(-1)
    if m["discrete"]:
        same &= actions.argmax(-1).cpu() == gold["actions"].argmax(-1)
    # a discount is a Bernoulli MODE (0 / 1): a head output within the bf16 band of 0 flips it, and with it the weights
    same &= torch.nan_to_num(ts.squeeze(-1).cpu(), nan=-1.0) == torch.nan_to_num(gold["discounts"], nan=-1.0)
    frac = same.all(0).float().mean().item()
    print(f"[parity] {name}: trajectories with identical draws over all {H} steps: {frac:.3f}")
    zs = states.combined
    values = agent.last_rollout["values"].unsqueeze(-1)
    vs = agent.critic.lambda_return(zs, rewards[:-1], ts, vs=values)
    vs2, w, adv = ops.lambda_return(rewards, values, ts, agent.critic.lambda_)
    assert torch.equal(vs, vs2)
    # bf16 contractions flip a near-tie draw now and then, after which that row imagines another trajectory: every
    # comparison runs on the rows whose draws equal the reference's over all H steps — against the oracle port re-run on
    # exactly those rows (the port is pinned to the reference on the whole fixture, tests/test_oracle.py) and, when no row
    # diverged, against the reference's own scalars.  The 1e-3 north-star bound is asserted in the split-operand mode
    # (tests/test_gpu_parity_mode.py); this is the default bf16 mode, whose band DESIGN.md derives.
    rows = same.all(0)
    sub = rows.nonzero().flatten()
    assert sub.numel() >= 0.4 * N, f"{name}: only {sub.numel()} of {N} rows follow the reference's draws"
    dsub = sub.cuda()
    pk = lambda t: t.index_select(1, dsub)
    losses_c, metrics_c = agent.critic.calculate_loss(pk(zs[:-1]), pk(vs), pk(w[:-1]), target_values=pk(values[:-1]))
    losses_a, metrics_a = agent.actor.calculate_loss(pk(zs[:-2]), pk(vs[1:]), pk(values[:-2]), pk(w[:-2]), pk(actions[1:-1]))
    ref_traj = orc.imagine(c["wm"], c["actor"], c["critic"], c["h0"][sub], c["z0"][sub], H=H, A=m["A"],
                           discrete=m["discrete"], predict_discount=m["predict_discount"],
                           latent_uniforms=c["lat"][:, sub], action_noise=c["act"][:, sub])
    ref = orc.ac_losses(ref_traj, c["actor"], c["critic"], lam=m["lam"], discrete=m["discrete"], rho=m["rho"],
                        eta=m["entropy_scale"])
    for k in ("loss_critic", "loss_actor", "loss_actor_reinforce", "loss_actor_dynamics_backprop", "loss_actor_entropy"):
        got = float((losses_c | losses_a)[k])
        want = float(ref[k])
        print(f"[parity] {name}.{k} on {sub.numel()}/{N} rows: ours {got:.6f} oracle {want:.6f}")
        # losses are means over (H x rows) terms; the reinforce term is a signed sum in which advantages cancel
        assert abs(got - want) <= 5e-3 * abs(want) + 2e-4, (name, k, got, want)
        if sub.numel() == N:
            assert abs(got - gold[k].item()) <= 5e-3 * abs(gold[k].item()) + 2e-4, (name, k, got, gold[k].item())
    e = ((pk(vs).squeeze(-1).cpu() - gold["vs"][:, sub]).pow(2).mean().sqrt() / gold["vs"][:, sub].pow(2).mean().sqrt()).item()
    print(f"[parity] {name}.lambda_returns rel-RMS vs reference on alive rows: {e:.3e}")
    assert e < 2e-2
    assert torch.equal(pk(w).squeeze(-1).cpu(), gold["w"][:, sub])


def test_train_step_end_to_end(cuda):
    """DreamerV2.train(RolloutChunks) — world-model half (torch) + hot path (K1/K2) — returns the
    reference's keys, finite values, and updates actor / critic parameters."""
    from rl_sandbox.utils.replay_buffer import RolloutChunks
    m = dict(D=200, A=5, discrete=True, layer_norm=True, predict_discount=True, entropy_scale=3e-3, gamma=0.999, H=5)
    torch.manual_seed(0)
    agent = make_agent(m, "cuda", batch_cluster_size=6)
    B, T = 3, 6
    obs = agent.preprocess_obs(torch.randint(0, 255, (B * T, 64, 64, 3), dtype=torch.uint8)).cuda()
    chunks = RolloutChunks(obs=obs, actions=torch.randint(0, 5, (B * T, 1)).cuda(), rewards=torch.randn(B * T).cuda(),
                           is_finished=torch.zeros(B * T).cuda(), is_first=torch.zeros(B * T).cuda(), additional_data={})
    before = [p.detach().clone() for p in agent.actor.parameters()]
    out = agent.train(chunks)
    expected = {"loss_wm", "loss_reconstruction", "loss_reconstruction_img", "loss_reward_pred", "loss_discount_pred",
                "loss_kl_reg", "loss_actor", "loss_actor_reinforce", "loss_actor_dynamics_backprop", "loss_actor_entropy",
                "loss_critic", "total", "reward_mean", "reward_std", "reward_sae", "prior_entropy", "posterior_entropy",
                "actor/avg_val", "actor/mean_val", "actor/avg_sd", "actor/min_val", "actor/max_val",
                "critic/avg_target_value", "critic/avg_lambda_value", "critic/avg_predicted_value"}
    assert expected <= set(out), expected - set(out)
    import numpy as np
    assert all(np.isfinite(v).all() for v in out.values())
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, agent.actor.parameters()))
    out2 = agent.train(chunks)   # second step repacks the changed weights
    assert np.isfinite(out2["total"]).all()
    # acting path and the metrics caller of imagine_trajectory(state, precomp_actions, horizon)
    agent.reset()
    a = agent.get_action(torch.randint(0, 255, (64, 64, 3), dtype=torch.uint8).numpy())
    assert 0 <= int(a) < 5
    with torch.no_grad():
        st, acts, rew, ts = agent.imagine_trajectory(agent._state, precomp_actions=torch.zeros(3, 1, 5).cuda(), horizon=3)
    assert st.determ.shape == (4, 1, 200) and rew.shape == (4, 1, 1)


def test_world_model_update_graph_replay_matches_eager(cuda):
    """The world-model half of train() replayed from a CUDA graph (forward + backward captured; clip + AdamW eager)
    is the eager computation: same Philox keys for the observe scan -> same losses and gradients on the first step,
    the same parameters after it, and the same loss trajectory over the following steps."""
    from rl_sandbox.utils.replay_buffer import RolloutChunks
    import numpy as np
    m = dict(D=200, A=5, discrete=True, layer_norm=True, predict_discount=True, entropy_scale=3e-3, gamma=0.999, H=5)
    B, T = 3, 6
    g = torch.Generator().manual_seed(5)
    obs = (torch.randint(0, 255, (B * T, 64, 64, 3), dtype=torch.uint8, generator=g).float() / 255 - 0.5).permute(0, 3, 1, 2).cuda()
    first = torch.zeros(B * T)
    first[::T] = 1
    chunks = RolloutChunks(obs=obs, actions=torch.randint(0, 5, (B * T, 1), generator=g).cuda(),
                           rewards=torch.randn(B * T, generator=g).cuda(), is_finished=torch.zeros(B * T).cuda(),
                           is_first=first.cuda(), additional_data={})
    runs = {}
    for graphed in (True, False):
        torch.manual_seed(0)
        agent = make_agent(m, "cuda", batch_cluster_size=T)
        agent.cuda_graph_wm = graphed
        p_init = torch.cat([p.detach().flatten() for p in agent.world_model.parameters()]).clone()
        out0 = agent.train(chunks)
        grads = {n: p.grad.detach().clone() for n, p in agent.world_model.named_parameters() if p.grad is not None}
        params = torch.cat([p.detach().flatten() for p in agent.world_model.parameters()]).clone()
        outs = [out0] + [agent.train(chunks) for _ in range(4)]
        assert bool(agent._wm_graphs) == graphed
        assert all(np.isfinite(v).all() for o in outs for v in o.values())
        runs[graphed] = (outs, grads, params)
    (og, gg, pg), (oe, ge, pe) = runs[True], runs[False]
    for k in ("loss_wm", "loss_reconstruction", "loss_reward_pred", "loss_kl_reg", "loss_discount_pred"):
        a, b = float(og[0][k]), float(oe[0][k])
        print(f"[parity] wm graph vs eager, step 1 {k}: {a:.6f} vs {b:.6f}")
        assert abs(a - b) <= 1e-4 * abs(b) + 1e-5, k
    assert set(gg) == set(ge)
    worst = max(((gg[n] - ge[n]).norm() / (ge[n].norm() + 1e-12)).item() for n in ge)
    print(f"[parity] wm graph vs eager, step 1 gradients: worst rel-L2 over {len(ge)} tensors {worst:.3e}")
    assert worst < 1e-2   # cuDNN / float-atomics reduction order only
    cos = torch.nn.functional.cosine_similarity(pg - p_init, pe - p_init, dim=0).item()
    print(f"[parity] wm graph vs eager, step 1 parameter displacement: cos {cos:.5f}")
    assert cos > 0.98 and (pg - pe).abs().max().item() <= 2.1e-4   # one AdamW step of lr 1e-4
    for k in ("loss_wm", "loss_kl_reg"):
        a = np.array([float(o[k]) for o in og])
        b = np.array([float(o[k]) for o in oe])
        print(f"[parity] wm graph vs eager {k}: {a.round(4).tolist()} vs {b.round(4).tolist()}")
        assert np.all(np.abs(a - b) <= 0.02 * np.abs(b)), k
    assert float(og[-1]["loss_wm"]) < float(og[0]["loss_wm"])


def test_acting_path_graph_replay(cuda):
    """get_action(): the batch-1 acting step replayed from a CUDA graph follows the eager op sequence — the first step
    after reset() (zero state, zero action: no randomness upstream of the posterior logits) gives the same recurrent
    state; actions stay valid, reset() reloads the static state, and the replay is several times faster."""
    import time
    m = dict(D=200, A=5, discrete=True, layer_norm=True, predict_discount=True, entropy_scale=3e-3, gamma=0.999, H=5)
    torch.manual_seed(0)
    agent = make_agent(m, "cuda", batch_cluster_size=6)
    g = torch.Generator().manual_seed(3)
    frames = [torch.randint(0, 255, (64, 64, 3), dtype=torch.uint8, generator=g).numpy() for _ in range(6)]
    res, times = {}, {}
    for graphed in (False, True):
        agent.cuda_graph_act = graphed
        agent.reset()
        a0 = agent.get_action(frames[0])
        res[graphed] = (agent._state.determ.clone(), agent._state.stoch_logits.clone())
        acts = [int(a0)] + [int(agent.get_action(f)) for f in frames[1:]]
        assert all(0 <= a < 5 for a in acts)
        assert torch.isfinite(agent._state.determ).all() and abs(agent._action_probs.sum().item() - 6) < 1e-3
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in frames * 5:
            agent.get_action(f)
        torch.cuda.synchronize()
        times[graphed] = (time.perf_counter() - t0) / 30 * 1e3
    assert agent._act_graph is not None
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(res[True][1], res[False][1], rtol=1e-4, atol=1e-4)
    print(f"[perf] get_action: eager {times[False]:.3f} ms, graph replay {times[True]:.3f} ms per call")
    assert times[True] < times[False]
    # reset() reloads the zero state into the static buffers
    agent.reset()
    agent.get_action(frames[0])
    torch.testing.assert_close(agent._state.determ, res[True][0], rtol=1e-4, atol=1e-5)


def test_train_step_continuous_actor_runs_fused(cuda):
    """config_dino-shaped agent (continuous actions, rho = 0): train() drives K1 (+tape) -> K2 -> K2 bwd -> K1 bwd -> K4;
    no torch autograd on the behaviour half.  Gradient parity is in test_gpu_ac_update.py."""
    from rl_sandbox.utils.replay_buffer import RolloutChunks
    import numpy as np
    m = dict(D=200, A=12, discrete=False, layer_norm=False, predict_discount=False, entropy_scale=1e-4, gamma=0.99, H=5)
    torch.manual_seed(0)
    agent = make_agent(m, "cuda", batch_cluster_size=6)
    assert agent._can_fuse_ac()
    B, T = 3, 6
    obs = agent.preprocess_obs(torch.randint(0, 255, (B * T, 64, 64, 3), dtype=torch.uint8)).cuda()
    chunks = RolloutChunks(obs=obs, actions=torch.randn(B * T, 12).cuda(), rewards=torch.randn(B * T).cuda(),
                           is_finished=torch.zeros(B * T).cuda(), is_first=torch.zeros(B * T).cuda(), additional_data={})
    before = [p.detach().clone() for p in agent.actor.parameters()]
    called = []
    orig = agent._imagine_autograd
    agent._imagine_autograd = lambda *a, **k: called.append(1) or orig(*a, **k)
    out = agent.train(chunks)
    assert not called, "the torch replay of the rollout must not run on the fused path"
    assert all(np.isfinite(v).all() for v in out.values())
    assert out["loss_actor_dynamics_backprop"] != 0
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, agent.actor.parameters()))
    assert np.isfinite(agent.train(chunks)["total"]).all()


def test_checkpoint_roundtrip_uses_reference_key_format(cuda, tmp_path, monkeypatch):
    m = dict(D=200, A=5, discrete=True, layer_norm=False, predict_discount=False, entropy_scale=3e-3, gamma=0.99, H=3)
    agent = make_agent(m, "cuda", batch_cluster_size=4)
    monkeypatch.chdir(tmp_path)
    agent.save_ckpt(7, {"total": 1.5})
    ck = torch.load(tmp_path / "dreamerV2-7-1.5.ckpt", weights_only=False)
    assert all(k.startswith("_orig_mod.") for k in ck["world_model_state_dict"])      # dreamer_v2.py:54,226
    assert all(k.startswith("_orig_mod.") for k in ck["critic_state_dict"]) and "actor.0.weight" in ck["actor_state_dict"]
    other = make_agent(m, "cuda", batch_cluster_size=4)
    assert other.load_ckpt(tmp_path / "dreamerV2-7-1.5.ckpt") == 7
    for a, b in zip(agent.actor.parameters(), other.actor.parameters()):
        assert torch.equal(a, b)
