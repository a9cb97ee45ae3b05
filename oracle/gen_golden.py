"""gen_golden.py — TEST INFRASTRUCTURE.  Generates tests/golden/*.npz|json by RUNNING THE REFERENCE
(imported unmodified from /root/reference through oracle/ref_harness.py) on seeded inputs and
explicit noise.  Run in the build container only:  python -m oracle.gen_golden

Fixtures
  lambda_known_answers.json   the five known-answer vectors of the reference's own
                              test/dreamer/test_critic.py:14-62, repaired for today's signature
                              (SURVEY 4): the code reads vs[i+1], so the value sequence is shifted
                              by one (a dummy v_0 is prepended) and gamma is folded into ds.  Each
                              expected vector is the literal from the reference test AND is
                              re-checked here against the reference's ImaginativeCritic._lambda_return.
  slot_attention.npz          SlotAttention.forward of the reference (vision/slot_attention.py:52-77),
                              4 slots x 384, 196 tokens, 2 iterations, explicit prev_slots.
                              Also the GRADIENTS the reference's autograd produces for
                              loss_actor.backward() / loss_critic.backward() (optimizer.py:55-57):
                              d loss_actor / d actions (continuous actors: the quantity rlsb_imagine_bwd
                              returns), and per parameter tensor its L2 norm and 64 probed entries
                              (indices from grad_probe_indices()).
  imagine_<case>.npz          DreamerV2.imagine_trajectory (dreamer_v2.py:68-96) outputs, the
                              target-critic values, lambda-returns (ac.py:64-66), cumprod weights
                              (dreamer_v2.py:192-197) and the critic / actor losses
                              (ac.py:68-81, 113-146) computed by the reference's own methods.
  acting.npz                  DreamerV2.get_action (dreamer_v2.py:139-154) over four frames from reset(): recurrent state,
                              actor probabilities and the action of every step (torch CPU generator seeded).
Parameters are NOT stored: they are regenerated from a seed by oracle_port.make_params (the same
tensors are loaded into the reference modules here), start states by oracle_port.make_start.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from . import oracle_port as orc
from . import ref_harness as rh

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

CASES = {
    # config 1 dims (agent/dreamer_v2_crafter.yaml): D=1024, discrete A=17, layer_norm, discount head
    "c1": dict(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True, N=6, H=3,
               entropy_scale=3e-3, gamma=0.999, param_seed=11, start_seed=12, noise_seed=13),
    # config 2 dims (agent/dreamer_v2.yaml): D=200, continuous A=12, no layer_norm, no discount head
    "c2": dict(D=200, A=12, discrete=False, layer_norm=False, predict_discount=False, N=6, H=3,
               entropy_scale=1e-5, gamma=0.99, param_seed=21, start_seed=22, noise_seed=23),
    # config-2 dims with layer_norm on: pins the LayerNorm-backward branches of the rollout's backward pass
    "c2_ln": dict(D=200, A=12, discrete=False, layer_norm=True, predict_discount=False, N=6, H=3,
                  entropy_scale=1e-5, gamma=0.99, param_seed=51, start_seed=52, noise_seed=53),
    # longer horizon / more rows, config-2 dims (cheap to store)
    "c2_long": dict(D=200, A=6, discrete=True, layer_norm=True, predict_discount=True, N=40, H=15,
                    entropy_scale=1e-4, gamma=0.99, param_seed=31, start_seed=32, noise_seed=33),
    # config 1 dims at the configured horizon, one full M tile of start states (N = 128, H = 15): the BASELINE shape's
    # arithmetic (D = 1024 split-row LayerNorms, 12 n-blocks of the GRU contraction) against the reference itself
    # (the two large tensors, determ and logits, are stored for the first `store_rows` start states only; rows are
    # independent, everything else covers all 128)
    "c1_long": dict(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True, N=128, H=15,
                    entropy_scale=3e-3, gamma=0.999, param_seed=81, start_seed=82, noise_seed=83, store_rows=40),
}


def make_noise(case):
    g = torch.Generator().manual_seed(case["noise_seed"])
    H, N, A = case["H"], case["N"], case["A"]
    lat = torch.rand(H, N, 1024, generator=g)
    act = torch.rand(H, N, A, generator=g) if case["discrete"] else torch.randn(H, N, A, generator=g)
    return lat, act


def known_answers():
    # literals from the reference's test/dreamer/test_critic.py:14-62
    t = lambda *a: [float(x) for x in a]
    ar = list(range(1, 11))
    cases = [
        dict(name="discount_0", lam=0.0, gamma=0.0, rs=t(*ar), vs_old=[1.0] * 10, expected=t(*ar), ref="test_critic.py:14-23"),
        dict(name="lambda_0", lam=0.0, gamma=1.0, rs=[1.0] * 10, vs_old=t(*ar), expected=t(*range(2, 12)), ref="test_critic.py:25-34"),
        dict(name="lambda_0_gamma_0_5", lam=0.0, gamma=0.5, rs=[1.0] * 10, vs_old=t(2, 2, 4, 4, 6, 6, 8, 8, 10, 10),
             expected=t(2, 2, 3, 3, 4, 4, 5, 5, 6, 6), ref="test_critic.py:36-45"),
        dict(name="lambda_1", lam=1.0, gamma=1.0, rs=[1.0] * 10, vs_old=t(*ar), expected=t(*range(20, 10, -1)), ref="test_critic.py:47-56"),
        dict(name="lambda_1_gamma_0_5", lam=1.0, gamma=0.5, rs=[0.0] * 10, vs_old=t(*[2 ** k for k in range(1, 11)]),
             expected=t(*[2 ** k for k in range(0, 10)]), ref="test_critic.py:58-62"),
    ]
    rh._import_reference()
    from rl_sandbox.agents.dreamer.ac import ImaginativeCritic
    critic = ImaginativeCritic(discount_factor=1, update_interval=100, soft_update_fraction=1,
                               value_target_lambda=0.95, latent_dim=10, layer_norm=False)
    for c in cases:
        c["vs"] = [0.0] + c["vs_old"]               # state values v_0..v_10 (v_0 is never read)
        c["ds"] = [c["gamma"]] * 11                   # gamma * ts
        critic.lambda_ = c["lam"]
        got = critic._lambda_return(torch.tensor(c["vs"]), torch.tensor(c["rs"]), torch.tensor(c["ds"]))
        assert torch.equal(got, torch.tensor(c["expected"])), (c["name"], got)
    (OUT / "lambda_known_answers.json").write_text(json.dumps(cases, indent=1))
    print("lambda_known_answers.json: 5 vectors verified against the reference")


def grad_probe_indices(numel: int, k: int = 64) -> torch.Tensor:
    """the probed flat indices of a parameter gradient (shared by the generator and the tests)"""
    g = torch.Generator().manual_seed(1000003 + numel)
    return torch.randint(0, numel, (k,), generator=g)


def run_case(name, case):
    D, A, H, N = case["D"], case["A"], case["H"], case["N"]
    wm_sd, actor_sd, critic_sd = orc.make_params(case["param_seed"], D=D, A=A, discrete=case["discrete"],
                                                 layer_norm=case["layer_norm"], predict_discount=case["predict_discount"])
    h0, z0 = orc.make_start(case["start_seed"], N, D)
    lat, act = make_noise(case)
    agent = rh.build_agent(D=D, A=A, discrete=case["discrete"], layer_norm=case["layer_norm"],
                           predict_discount=case["predict_discount"], H=H, entropy_scale=case["entropy_scale"],
                           gamma=case["gamma"])
    rh.load_params(agent, wm_sd, actor_sd, critic_sd)
    state = rh.ref_state(agent, h0, z0)
    q = rh.NoiseQueue(lat, act)
    # observation seam (no arithmetic changed): remember the action tensor each predict_next receives so that
    # d loss_actor / d a_t can be read off after backward
    seen_actions = []
    orig_predict_next = agent.world_model.predict_next

    def spy_predict_next(prev_state, a):
        if a.requires_grad:
            a.retain_grad()
        seen_actions.append(a)
        return orig_predict_next(prev_state, a)

    agent.world_model.predict_next = spy_predict_next
    with rh.injected_noise(q):
        # agents/dreamer_v2.py:182-207 (second half of DreamerV2.train), reference methods only
        states, actions, rewards, discounts = agent.imagine_trajectory(state)
        zs = states.combined
        rewards = agent.world_model.reward_normalizer(rewards.float())
        discounts = discounts.float()
        values = agent.critic.target_critic(zs).mode
        vs = agent.critic.lambda_return(zs, rewards[:-1], discounts)
        w = torch.cumprod(torch.cat([torch.ones_like(discounts[:1]), discounts[:-1]], dim=0), dim=0).detach()
        losses_c, metrics_c = agent.critic.calculate_loss(zs[:-1], vs, w[:-1])
        losses_a, metrics_a = agent.actor.calculate_loss(zs[:-2], vs[1:], agent.critic.target_critic(zs[:-2]).mode,
                                                         w[:-2], actions[1:-1])
        # the gradients of Optimizer.step (utils/optimizer.py:55-57), reference autograd
        for p_ in list(agent.actor.parameters()) + list(agent.critic.parameters()):
            p_.grad = None
        losses_a["loss_actor"].backward(retain_graph=True)
        losses_c["loss_critic"].backward()
    grad_names, grad_norms, grad_probes = [], [], []
    for prefix, mod in (("actor.", agent.actor.actor), ("critic.", agent.critic.critic)):
        for n_, p_ in mod.named_parameters():
            g_ = p_.grad if p_.grad is not None else torch.zeros_like(p_)
            grad_names.append(prefix + n_)
            grad_norms.append(g_.norm().item())
            grad_probes.append(g_.flatten()[grad_probe_indices(g_.numel())].numpy())
    if seen_actions and seen_actions[0].grad is not None:
        g_actions = torch.cat([a_.grad for a_ in seen_actions]).detach().numpy()   # (H, N, A): d loss_actor / d a_t
    else:
        g_actions = np.zeros((0,), np.float32)
    assert q.log[:2] == (["action", "latent"] if case["discrete"] else ["action_normal", "latent"]), q.log[:4]
    assert not q.latent and not q.action, "noise not fully consumed"
    f = lambda x: x.detach().squeeze(-1).numpy().astype(np.float32) if x.dim() == 3 and x.shape[-1] == 1 else x.detach().numpy().astype(np.float32)
    R = case.get("store_rows", N)
    out = dict(
        determ=states.determ.detach()[:, :R].numpy(), logits=states.stoch_logits.detach().reshape(H + 1, N, 1024)[:, :R].numpy(),
        stoch_idx=states.stoch.detach().reshape(H + 1, N, 32, 32).argmax(-1).numpy().astype(np.uint8),
        actions=actions.detach().numpy(), rewards=f(rewards), discounts=f(discounts), values=f(values), vs=f(vs), w=f(w),
        loss_critic=np.float32(losses_c["loss_critic"].item()), loss_actor=np.float32(losses_a["loss_actor"].item()),
        loss_actor_reinforce=np.float32(float(losses_a["loss_actor_reinforce"])),
        loss_actor_dynamics_backprop=np.float32(float(losses_a["loss_actor_dynamics_backprop"])),
        loss_actor_entropy=np.float32(float(losses_a["loss_actor_entropy"])),
        critic_avg_target_value=np.float32(metrics_c["critic/avg_target_value"].item()),
        critic_avg_lambda_value=np.float32(metrics_c["critic/avg_lambda_value"].item()),
        critic_avg_predicted_value=np.float32(metrics_c["critic/avg_predicted_value"].item()),
        grad_actions=g_actions.astype(np.float32), grad_norms=np.asarray(grad_norms, np.float32),
        grad_probes=np.stack(grad_probes).astype(np.float32),
        meta=json.dumps({**case, "lam": 0.95, "rho": 1.0 if case["discrete"] else 0.0, "grad_names": grad_names}),
    )
    np.savez_compressed(OUT / f"imagine_{name}.npz", **out)
    print(f"imagine_{name}.npz written:", {k: getattr(v, 'shape', None) for k, v in out.items() if k != 'meta'})


SLOTTED_CASE = dict(D=200, A=1, K=4, discrete=False, layer_norm=True, predict_discount=False, N=5, H=3,
                    entropy_scale=1e-4, gamma=0.999, param_seed=61, start_seed=62, noise_seed=63, blocks=3)


def slotted_inputs(case=SLOTTED_CASE):
    D, A, K, N, H = case["D"], case["A"], case["K"], case["N"], case["H"]
    h0, z0 = orc.make_start(case["start_seed"], N * K, D)
    g = torch.Generator().manual_seed(case["noise_seed"])
    lat = torch.rand(H, N, K, 1024, generator=g)
    act = torch.rand(H, N, A, generator=g) if case["discrete"] else torch.randn(H, N, A, generator=g)
    return h0.view(N, K, D), z0.view(N, K, 1024), lat, act


def run_slotted():
    """config_slotted dims: DreamerV2.imagine_trajectory over world_model_slots_attention.WorldModel (reference)."""
    c = SLOTTED_CASE
    wm_sd, actor_sd, critic_sd = orc.make_params_slotted(c["param_seed"], D=c["D"], A=c["A"], K=c["K"],
                                                         discrete=c["discrete"], layer_norm=c["layer_norm"],
                                                         predict_discount=c["predict_discount"])
    h0, z0, lat, act = slotted_inputs()
    agent = rh.build_agent_slotted(D=c["D"], A=c["A"], K=c["K"], discrete=c["discrete"], layer_norm=c["layer_norm"],
                                   predict_discount=c["predict_discount"], H=c["H"], entropy_scale=c["entropy_scale"],
                                   gamma=c["gamma"], attention_block_num=c["blocks"])
    rh.load_params(agent, {k: v for k, v in wm_sd.items() if k != "pos_enc"}, actor_sd, critic_sd)
    wm = getattr(agent.world_model, "_orig_mod", agent.world_model)
    assert torch.allclose(wm.pos_enc, wm_sd["pos_enc"], atol=1e-6), "pos_enc restatement differs"
    state = rh.ref_state_slotted(agent, h0, z0)
    q = rh.NoiseQueue(list(lat), list(act))
    with rh.injected_noise(q), torch.no_grad():
        states, actions, rewards, discounts = agent.imagine_trajectory(state)
        values = agent.critic.target_critic(states.combined).mode
    assert not q.latent and not q.action, "noise not fully consumed"
    H, N, K = c["H"], c["N"], c["K"]
    np.savez_compressed(
        OUT / "imagine_slotted.npz", determ=states.determ.numpy(), logits=states.stoch_logits.reshape(H + 1, N, K, 1024).numpy(),
        stoch_idx=states.stoch.reshape(H + 1, N, K, 32, 32).argmax(-1).numpy().astype(np.uint8), actions=actions.numpy(),
        rewards=rewards.squeeze(-1).numpy(), discounts=discounts.squeeze(-1).numpy(), values=values.squeeze(-1).numpy(),
        meta=json.dumps(c))
    print("imagine_slotted.npz written:", states.determ.shape, states.stoch_logits.shape, actions.shape)


OBSERVE_CASE = dict(D=200, A=5, layer_norm=True, B=3, T=4, E=1536, param_seed=71, input_seed=72)


def observe_inputs(case=OBSERVE_CASE):
    g = torch.Generator().manual_seed(case["input_seed"])
    T, B = case["T"], case["B"]
    embed = torch.randn(T, B, case["E"], generator=g)
    actions = torch.randn(T, B, case["A"], generator=g)
    uniforms = torch.rand(T, B, 1024, generator=g)
    weights = {k: torch.randn(T, B, n, generator=g) / n ** 0.5
               for k, n in (("prior_logits", 1024), ("post_logits", 1024), ("determ", case["D"]), ("stoch", 1024))}
    return embed, actions, uniforms, weights


def run_observe():
    """The observe loop of WorldModel.calculate_loss (world_model.py:187-202) run with the reference's RSSM.forward and
    State objects; outputs + the reference autograd's gradients of a seeded probe loss (pins rlsb_observe_fwd / _bwd)."""
    c = OBSERVE_CASE
    wm_sd, actor_sd, critic_sd = orc.make_params(c["param_seed"], D=c["D"], A=c["A"], discrete=False,
                                                 layer_norm=c["layer_norm"], predict_discount=False)
    agent = rh.build_agent(D=c["D"], A=c["A"], discrete=False, layer_norm=c["layer_norm"], predict_discount=False, H=3)
    rh.load_params(agent, wm_sd, actor_sd, critic_sd)
    wm = getattr(agent.world_model, "_orig_mod", agent.world_model)
    embed, actions, uniforms, weights = observe_inputs()
    embed = embed.clone().requires_grad_()
    from rl_sandbox.agents.dreamer.rssm import State
    q = rh.NoiseQueue(list(uniforms), [])
    for p_ in wm.parameters():
        p_.grad = None
    with rh.injected_noise(q):
        prev = wm.get_initial_state(c["B"])
        priors, posts = [], []
        for t in range(c["T"]):
            prior, post, _ = wm.recurrent_model.forward(prev, embed[t].unsqueeze(0), actions[t].unsqueeze(0))
            prev = post
            priors.append(prior)
            posts.append(post)
        posterior, prior = State.stack(posts), State.stack(priors)
        o = dict(prior_logits=prior.stoch_logits.reshape(c["T"], c["B"], 1024),
                 post_logits=posterior.stoch_logits.reshape(c["T"], c["B"], 1024), determ=posterior.determ,
                 stoch=posterior.stoch)
        orc.observe_probe_loss(o, weights).backward()
    assert not q.latent, "noise not fully consumed"
    names = [n for n in orc.OBSERVE_PARAM_KEYS if "recurrent_model." + n in dict(wm.named_parameters())]
    params = dict(wm.named_parameters())
    np.savez_compressed(
        OUT / "observe.npz", prior_logits=o["prior_logits"].detach().numpy(), post_logits=o["post_logits"].detach().numpy(),
        determ=o["determ"].detach().numpy(), stoch_idx=o["stoch"].detach().reshape(c["T"], c["B"], 32, 32).argmax(-1).numpy().astype(np.uint8),
        grad_embed=embed.grad.numpy(),
        grad_norms=np.asarray([params["recurrent_model." + n].grad.norm().item() for n in names], np.float32),
        grad_probes=np.stack([params["recurrent_model." + n].grad.flatten()[grad_probe_indices(params["recurrent_model." + n].numel())].numpy()
                              for n in names]).astype(np.float32),
        meta=json.dumps({**c, "grad_names": names}))
    print("observe.npz written:", o["determ"].shape, len(names), "parameter gradients")


SLOT_CASE = dict(B=3, tokens=196, dim=384, slots=4, iters=2, param_seed=41, input_seed=42)


def slot_inputs(case=SLOT_CASE):
    g = torch.Generator().manual_seed(case["input_seed"])
    X = torch.randn(case["B"], case["tokens"], case["dim"], generator=g)
    prev = torch.randn(case["B"], case["slots"], case["dim"], generator=g)
    return X, prev


def slot_grad_weights(case=SLOT_CASE):
    g = torch.Generator().manual_seed(case["input_seed"] + 1)
    return torch.randn(case["B"], case["slots"], case["dim"], generator=g)


def run_slot_attention():
    """SlotAttention.forward of the reference (vision/slot_attention.py:52-77) with explicit prev_slots."""
    rh._import_reference()
    from rl_sandbox.vision.slot_attention import SlotAttention
    c = SLOT_CASE
    sd = orc.make_slot_params(c["param_seed"], c["dim"], c["slots"])
    mod = SlotAttention(c["slots"], c["dim"], c["iters"], use_prev_slots=False)
    mod.load_state_dict(sd, strict=True)
    X, prev = slot_inputs()
    with torch.no_grad():
        out = mod(X, prev)
    # gradients of the reference's autograd for the scalar sum(out * G), G seeded (pins rlsb_slot_attention_bwd)
    Xg, pg = X.clone().requires_grad_(), prev.clone().requires_grad_()
    G = slot_grad_weights()
    (mod(Xg, pg) * G).sum().backward()
    names = [n for n, _ in mod.named_parameters() if not n.startswith("slots_mu") and not n.startswith("slots_logsigma")]
    params = dict(mod.named_parameters())
    np.savez_compressed(OUT / "slot_attention.npz", slots=out.numpy(), attn=mod.last_attention.detach().numpy(),
                        grad_X=Xg.grad.numpy(), grad_prev=pg.grad.numpy(),
                        grad_norms=np.asarray([params[n].grad.norm().item() for n in names], np.float32),
                        grad_probes=np.stack([params[n].grad.flatten()[grad_probe_indices(params[n].numel())].numpy()
                                              for n in names]).astype(np.float32),
                        meta=json.dumps({**c, "grad_names": names}))
    print("slot_attention.npz written:", out.shape, mod.last_attention.shape, len(names), "parameter gradients")


ACT_CASE = dict(D=200, A=6, discrete=True, layer_norm=True, predict_discount=True, steps=4, param_seed=91, enc_seed=92,
                frame_seed=93, torch_seed=94)


def acting_frames(case=ACT_CASE):
    g = torch.Generator().manual_seed(case["frame_seed"])
    return torch.randint(0, 256, (case["steps"], 64, 64, 3), generator=g, dtype=torch.uint8)


def run_acting():
    """DreamerV2.get_action of the reference (dreamer_v2.py:139-154: preprocess, conv encoder on one frame, one RSSM.forward
    step, actor, action draw) over a few frames from reset(), torch CPU generator seeded: the recurrent state after every
    step, the actor's probabilities and the returned action (pins the acting path of the host mirror)."""
    c = ACT_CASE
    wm_sd, actor_sd, critic_sd = orc.make_params(c["param_seed"], D=c["D"], A=c["A"], discrete=c["discrete"],
                                                 layer_norm=c["layer_norm"], predict_discount=c["predict_discount"])
    agent = rh.build_agent(D=c["D"], A=c["A"], discrete=c["discrete"], layer_norm=c["layer_norm"],
                           predict_discount=c["predict_discount"], H=3)
    rh.load_params(agent, wm_sd, actor_sd, critic_sd)
    wm = getattr(agent.world_model, "_orig_mod", agent.world_model)
    wm.encoder.load_state_dict(orc.seeded_module_params(wm.encoder, c["enc_seed"]))
    frames = acting_frames()
    agent.reset()
    torch.manual_seed(c["torch_seed"])
    out = {k: [] for k in ("determ", "post_logits", "stoch_idx", "probs", "action")}
    with torch.no_grad():
        for f in frames:
            a = agent.get_action(f.numpy())
            st = agent._state
            out["determ"].append(st.determ.reshape(-1).numpy().copy())
            out["post_logits"].append(st.stoch_logits.reshape(-1).numpy().copy())
            out["stoch_idx"].append(st.stoch.reshape(32, 32).argmax(-1).numpy().astype(np.uint8))
            out["probs"].append(agent.actor.get_action(st).probs.reshape(-1).numpy().copy())
            out["action"].append(int(a))
    np.savez_compressed(OUT / "acting.npz", **{k: np.stack(v) for k, v in out.items()},
                        action_probs_sum=agent._action_probs.numpy().copy(), meta=json.dumps(c))
    print("acting.npz written:", out["action"])


def main(argv=None):
    """no arguments: every fixture; otherwise the named imagine_<case> fixtures only (python -m oracle.gen_golden c1_long)"""
    import sys
    only = list(sys.argv[1:] if argv is None else argv)
    assert rh.available(), "reference checkout not found"
    OUT.mkdir(parents=True, exist_ok=True)
    if not only:
        known_answers()
        run_slot_attention()
        run_slotted()
        run_observe()
        run_acting()
    if only == ["acting"]:
        run_acting()
        return
    for name, case in CASES.items():
        if not only or name in only:
            run_case(name, case)


if __name__ == "__main__":
    main()
