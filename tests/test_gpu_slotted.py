"""GPU parity of the slotted imagination step (K1 with slots > 1: rlsb_imagine.cu + rlsb_mixer.cu) against the
reference's slotted world model (tests/golden/imagine_slotted.npz) and against the oracle at a ragged size."""
import pytest
import torch

from oracle import oracle_port as orc
from tests.test_oracle import _load_slotted

pytestmark = pytest.mark.gpu


def _engine(m, H, blocks=3, coeff=1.0):
    from rl_sandbox_b200 import ops
    cfg = ops.ImagineConfig(D=m["D"], A=m["A"], discrete=m["discrete"], layer_norm=m["layer_norm"],
                            predict_discount=m["predict_discount"], H=H, slots=m["K"], attention_blocks=blocks,
                            mixer_coeff=coeff)
    return ops.ImaginationEngine(cfg)


def _rel(a, b):
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def test_slotted_rollout_matches_reference(cuda):
    m, gold, wm, actor, critic, h0, z0, lat, act = _load_slotted()
    H, N, K = m["H"], m["N"], m["K"]
    eng = _engine(m, H, m["blocks"])
    to = lambda sd: {k: v.cuda() for k, v in sd.items()}
    eng.pack(to(wm), to(actor), to(critic))
    out = eng.rollout(h0.cuda(), z0.cuda(), None, lat.cuda(), act.cuda())
    torch.cuda.synchronize()
    assert out["determ"].shape == (H + 1, N, K, m["D"]) and out["logits"].shape == (H + 1, N, K, 1024)
    assert out["stoch_idx"].shape == (H + 1, N, K, 32) and out["rewards"].shape == (H + 1, N)
    # (1) indices are bit-exact given the kernel's own logits and the uniforms
    own = orc.sample_categorical(out["logits"][1:].cpu().view(H, N, K, 32, 32), lat.view(H, N, K, 32, 32))
    assert torch.equal(own, out["stoch_idx"][1:].cpu().long())
    # (2) first transition from the reference's start state (no accumulated divergence)
    for k, tol in (("determ", 3e-3), ("logits", 8e-3)):
        e = _rel(out[k][1].cpu(), gold[k][1])
        print(f"[parity] slotted step 1 {k}: rel-RMS vs reference {e:.3e}")
        assert e < tol, (k, e)
    for k in ("rewards", "values"):
        e = _rel(out[k][:2].cpu(), gold[k][:2])
        print(f"[parity] slotted {k}[0:2]: rel-RMS vs reference {e:.3e}")
        assert e < 2e-2, (k, e)
    # (3) whole trajectories where every draw coincides with the reference's
    same = (out["stoch_idx"].cpu() == gold["stoch_idx"]).all(-1).all(-1).cumprod(0).bool()
    frac = same[-1].float().mean().item()
    print(f"[parity] slotted: start states with identical draws over all {H} steps: {frac:.2f}")
    if same[-1].any():
        e = _rel(out["determ"].cpu()[:, same[-1]], gold["determ"][:, same[-1]])
        ea = _rel(out["actions"].cpu()[:, same[-1]], gold["actions"][:, same[-1]])
        print(f"[parity] slotted trajectories: determ rel-RMS {e:.3e}, actions {ea:.3e}")
        assert e < 5e-3 and ea < 2e-2


@pytest.mark.parametrize("N,K,blocks,coeff", [(300, 4, 3, 1.0), (77, 2, 1, 0.4)])
def test_slotted_rollout_matches_bf16_oracle(cuda, N, K, blocks, coeff):
    """ragged sizes / other slot counts against the oracle with bf16-rounded contraction operands"""
    m = dict(D=200, A=3, K=K, discrete=True, layer_norm=True, predict_discount=True)
    H = 4
    wm, actor, critic = orc.make_params_slotted(5, D=200, A=3, K=K, discrete=True, layer_norm=True, predict_discount=True)
    h0, z0 = orc.make_start(6, N * K, 200)
    h0, z0 = h0.view(N, K, 200), z0.view(N, K, 1024)
    g = torch.Generator().manual_seed(7)
    lat, act = torch.rand(H, N, K, 1024, generator=g), torch.rand(H, N, 3, generator=g)
    ref = orc.imagine_slotted(wm, actor, critic, h0, z0, H=H, A=3, K=K, discrete=True, predict_discount=True,
                              latent_uniforms=lat, action_noise=act, blocks=blocks, coeff=coeff, bf16=True)
    eng = _engine(m, H, blocks, coeff)
    to = lambda sd: {k: v.cuda() for k, v in sd.items()}
    eng.pack(to(wm), to(actor), to(critic))
    out = eng.rollout(h0.cuda(), z0.cuda(), None, lat.cuda(), act.cuda())
    same = ((out["stoch_idx"].cpu().long() == ref["stoch_idx"]).all(-1).all(-1) &
            (out["actions"].cpu().argmax(-1) == ref["actions"].argmax(-1))).cumprod(0).bool()
    frac = same[-1].float().mean().item()
    e = _rel(out["determ"].cpu()[:, same[-1]], ref["determ"][:, same[-1]])
    el = _rel(out["logits"].cpu()[1:, same[-1]], ref["logits"][1:, same[-1]])
    print(f"[parity] slotted N={N} K={K}: identical draws {frac:.3f}, determ rel-RMS {e:.2e}, logits {el:.2e}")
    assert frac > 0.85 and e < 1e-3 and el < 3e-3   # a flipped draw (2e-4 per draw) ends a trajectory comparison


def test_slotted_agent_end_to_end(cuda):
    """config_slotted-shaped agent through the alias package: train() (world-model loss with K3-backed slot attention
    under torch autograd, behaviour half through the torch replay because the actor is continuous) and the K1 slotted
    rollout behind imagine_trajectory under no_grad (metrics / acting callers)."""
    from functools import partial
    import numpy as np
    from rl_sandbox.agents import DreamerV2
    from rl_sandbox.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
    from rl_sandbox.agents.dreamer.world_model_slots_attention import WorldModel
    from rl_sandbox.utils.optimizer import Optimizer
    from rl_sandbox.utils.replay_buffer import RolloutChunks
    torch.manual_seed(0)
    opt = partial(Optimizer, lr=8e-5, eps=1e-5, weight_decay=1e-6, clip=100)
    agent = DreamerV2(
        obs_space_num=[64, 64, 3], clip_rewards="tanh", actions_num=1,
        world_model=partial(WorldModel, batch_cluster_size=4, latent_dim=32, latent_classes=32, rssm_dim=200, slots_num=4,
                            slots_iter_num=2, kl_loss_scale=1000, kl_loss_balancing=0.8, kl_free_nats=5e-4,
                            discrete_rssm=False, decode_vit=False, vit_l2_ratio=0.75, use_prev_slots=False,
                            encode_vit=False, predict_discount=False, layer_norm=True),
        actor=partial(ImaginativeActor, layer_norm=True, reinforce_fraction=None, entropy_scale=1e-4),
        critic=partial(ImaginativeCritic, discount_factor=0.999, update_interval=100, soft_update_fraction=1,
                       value_target_lambda=0.95, layer_norm=True),
        action_type="continuous", imagination_horizon=4, wm_optim=opt, actor_optim=opt, critic_optim=opt,
        layer_norm=True, batch_cluster_size=4, f16_precision=False, device_type="cuda")
    B, T = 2, 4
    obs = agent.preprocess_obs(torch.randint(0, 255, (B * T, 64, 64, 3), dtype=torch.uint8)).cuda()
    chunks = RolloutChunks(obs=obs, actions=torch.randn(B * T, 1).cuda(), rewards=torch.randn(B * T).cuda(),
                           is_finished=torch.zeros(B * T).cuda(), is_first=torch.zeros(B * T).cuda(), additional_data={})
    out = agent.train(chunks)
    assert all(np.isfinite(v).all() for v in out.values()), {k: v for k, v in out.items() if not np.isfinite(v).all()}
    assert out["loss_actor_dynamics_backprop"] != 0
    # both halves of train() replay from CUDA graphs once the mixer schedule has settled (the first step is eager):
    # losses stay finite, the world-model loss falls, actor and critic keep moving, and the graphed steps agree with
    # eager ones on the same batch to within the step-to-step change
    before = [p.detach().clone() for p in list(agent.actor.parameters()) + list(agent.critic.critic.parameters())]
    outs = [agent.train(chunks) for _ in range(4)]
    assert len(agent._wm_graphs) == 2, list(agent._wm_graphs)
    assert all(np.isfinite(v).all() for o in outs for v in o.values())
    assert float(outs[-1]["loss_wm"]) < float(out["loss_wm"])
    after = list(agent.actor.parameters()) + list(agent.critic.critic.parameters())
    assert all(not torch.equal(a, b.detach()) for a, b in zip(before, after))
    agent.cuda_graph_wm = False
    eager = agent.train(chunks)
    for k in ("loss_wm", "loss_critic", "loss_actor_dynamics_backprop"):
        a, b, c = float(outs[-2][k]), float(outs[-1][k]), float(eager[k])
        print(f"[parity] slotted train() {k}: graphed {a:.5f}, {b:.5f} -> eager {c:.5f}")
        assert abs(c - b) <= 3 * abs(b - a) + 0.05 * abs(b) + 1e-3, k
    agent.cuda_graph_wm = True
    # K1 (slots = 4) behind the reference's method surface
    state, _ = agent.world_model.get_initial_state(batch_size=6)
    with torch.no_grad():
        states, actions, rewards, ts = agent.imagine_trajectory(state, horizon=4)
    assert states.determ.shape == (5, 6, 4, 200) and states.stoch_logits.shape == (5, 6, 4, 32, 32)
    assert states.combined.shape == (5, 6, 4 * 1224) and actions.shape == (5, 6, 1) and rewards.shape == (5, 6, 1)
    assert agent.last_rollout is not None and torch.isfinite(rewards).all()
    # the same rollout replayed with torch ops on the K1 actions reproduces the K1 rewards (bf16 tolerance)
    with torch.no_grad():
        prev = state
        prev.stoch_ = states.stoch[0:1]
        rs = []
        for t in range(4):
            prev.stoch_ = states.stoch[t:t + 1]           # K1's own draws
            prior, r, _ = agent.world_model.predict_next(prev, actions[t + 1:t + 2])
            prior.stoch_ = states.stoch[t + 1:t + 2]
            rs.append(agent.world_model.reward_predictor(prior.combined).mode)
            prev = prior
        e = _rel(rewards[1:], torch.cat(rs))
    print(f"[parity] slotted agent: K1 rewards vs torch replay on the same draws: rel-RMS {e:.3e}")
    assert e < 3e-2
