"""CPU tests: the oracle (oracle/) against the golden vectors generated from the reference itself."""
import numpy as np
import pytest
import torch

from oracle import oracle_port as orc
from tests._golden import known_answers, load_case


def test_det_logf_accuracy():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(200000).astype(np.float32) * 0.9999 + 1e-7,
                        np.float32(10.0) ** rng.uniform(-19, 3, 50000).astype(np.float32)])
    y = orc.det_logf(x)
    ref = np.log(x.astype(np.float64))
    ulp = np.abs(y - ref) / np.spacing(np.abs(ref).astype(np.float32))
    assert ulp.max() < 1.5


def test_gumbel_distribution():
    g = torch.Generator().manual_seed(0)
    x = orc.gumbel(torch.rand(400000, generator=g))
    assert abs(x.mean().item() - 0.5772) < 0.01 and abs(x.var().item() - np.pi ** 2 / 6) < 0.03


def test_sampler_matches_softmax_and_argmax_ties():
    g = torch.Generator().manual_seed(1)
    lg = torch.tensor([[0.0, 1.0, 2.0, -1.0]]).repeat(200000, 1)
    idx = orc.sample_categorical(lg, torch.rand(200000, 4, generator=g))
    f = torch.bincount(idx, minlength=4) / 200000.0
    assert torch.allclose(f, torch.softmax(lg[0], 0), atol=5e-3)
    # identical scores -> lowest index, like torch.argmax
    assert orc.sample_categorical(torch.zeros(3, 5), torch.full((3, 5), 0.5)).tolist() == [0, 0, 0]


def test_philox_known_answer():
    # Random123 known-answer test for philox4x32-10 (kat_vectors): counter/key all ones-complement
    import ctypes as C
    out = (C.c_uint32 * 4)()
    orc.clib().orc_philox_raw(C.c_uint32(0xffffffff), C.c_uint32(0xffffffff), C.c_uint32(0xffffffff),
                              C.c_uint32(0xffffffff), C.c_uint32(0xffffffff), C.c_uint32(0xffffffff), out)
    assert [hex(v) for v in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    orc.clib().orc_philox_raw(C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), out)
    assert [hex(v) for v in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    u = orc.philox_uniform(42, 0, 0, 0, 1024, 64)
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01


@pytest.mark.parametrize("case", known_answers(), ids=lambda c: c["name"])
def test_lambda_known_answers(case):
    """The reference's own test vectors (test/dreamer/test_critic.py:14-62), repaired."""
    vs, rs, ds = (np.array(case[k], np.float32).reshape(-1, 1) for k in ("vs", "rs", "ds"))
    rs11 = np.concatenate([rs, np.zeros((1, 1), np.float32)])
    out, w, adv = orc.lambda_return_c(rs11, vs, ds, case["lam"])
    assert out[:, 0].tolist() == case["expected"]
    loop = orc.lambda_return_loop(torch.tensor(case["vs"]), torch.tensor(case["rs"]), torch.tensor(case["ds"]), case["lam"])
    assert loop.tolist() == case["expected"]


def test_lambda_c_oracle_is_bit_identical_to_reference_loop():
    g = torch.Generator().manual_seed(3)
    T, N = 16, 777
    r, v = torch.randn(T, N, generator=g), torch.randn(T, N, generator=g)
    # imagined discounts are Bernoulli modes, i.e. exactly 0 or 1 (world_model.py:136-139)
    d = (torch.rand(T, N, generator=g) > 0.15).float()
    vs, w, adv = orc.lambda_return_c(r.numpy(), v.numpy(), d.numpy(), 0.95)
    assert torch.equal(torch.from_numpy(vs), orc.lambda_return_loop(v, r[:-1], d, 0.95))
    w_ref = torch.cumprod(torch.cat([torch.ones_like(d[:1]), d[:-1]]), 0)
    assert torch.equal(torch.from_numpy(w), w_ref)
    assert torch.equal(torch.from_numpy(adv), torch.from_numpy(vs)[1:] - v[:-2])
    # general (non-binary) discounts: torch's CPU cumprod accumulates in double, the oracle in fp32
    d2 = d * 0.999
    vs2, w2, _ = orc.lambda_return_c(r.numpy(), v.numpy(), d2.numpy(), 0.95)
    assert torch.equal(torch.from_numpy(vs2), orc.lambda_return_loop(v, r[:-1], d2, 0.95))
    torch.testing.assert_close(torch.from_numpy(w2), torch.cumprod(torch.cat([torch.ones_like(d2[:1]), d2[:-1]]), 0),
                               rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["c1", "c2", "c2_long", "c1_long"])
def test_oracle_port_matches_reference_rollout(name):
    """oracle_port.imagine / ac_losses vs the tensors the reference's modules produced."""
    c = load_case(name)
    m, gold = c["meta"], c["gold"]
    out = orc.imagine(c["wm"], c["actor"], c["critic"], c["h0"], c["z0"], H=m["H"], A=m["A"], discrete=m["discrete"],
                      predict_discount=m["predict_discount"], latent_uniforms=c["lat"], action_noise=c["act"])
    assert torch.equal(out["stoch_idx"], gold["stoch_idx"].long()), "categorical indices must be bit-exact"
    for k in ("determ", "logits", "actions", "rewards", "values"):
        R = gold[k].shape[1]   # the large tensors of some fixtures cover the first `store_rows` start states only
        torch.testing.assert_close(out[k][:, :R], gold[k], rtol=1e-4, atol=2e-5, msg=lambda s: f"{k}: {s}")
    assert torch.equal(torch.nan_to_num(out["discounts"], nan=-1), torch.nan_to_num(gold["discounts"], nan=-1))
    losses = orc.ac_losses(out, c["actor"], c["critic"], lam=m["lam"], discrete=m["discrete"], rho=m["rho"],
                           eta=m["entropy_scale"])
    torch.testing.assert_close(losses["vs"], gold["vs"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(losses["w"], gold["w"], rtol=0, atol=0)
    for k in ("loss_critic", "loss_actor", "loss_actor_reinforce", "loss_actor_dynamics_backprop", "loss_actor_entropy"):
        torch.testing.assert_close(losses[k].float(), gold[k], rtol=1e-4, atol=1e-6, msg=lambda s: f"{k}: {s}")


def test_oracle_slot_attention_matches_reference():
    """oracle_port.slot_attention vs the reference's SlotAttention.forward (tests/golden/slot_attention.npz)."""
    import json
    from tests._golden import GOLDEN
    from oracle.gen_golden import SLOT_CASE, slot_inputs
    from oracle.gen_golden import slot_grad_weights
    z = np.load(GOLDEN / "slot_attention.npz")
    meta = json.loads(str(z["meta"]))
    assert {k: meta[k] for k in SLOT_CASE} == SLOT_CASE, "fixture is stale: re-run python -m oracle.gen_golden"
    sd = orc.make_slot_params(SLOT_CASE["param_seed"], SLOT_CASE["dim"], SLOT_CASE["slots"])
    X, prev = slot_inputs()
    out, attn = orc.slot_attention(X, prev, sd, SLOT_CASE["iters"])
    torch.testing.assert_close(out, torch.from_numpy(z["slots"]), rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(attn, torch.from_numpy(z["attn"]), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(attn.sum(-1), torch.ones(3, 4), rtol=1e-5, atol=1e-5)
    # the port's autograd reproduces the reference's gradients of sum(out * G) w.r.t. the inputs
    Xg, pg = X.clone().requires_grad_(), prev.clone().requires_grad_()
    (orc.slot_attention(Xg, pg, sd, SLOT_CASE["iters"])[0] * slot_grad_weights()).sum().backward()
    torch.testing.assert_close(Xg.grad, torch.from_numpy(z["grad_X"]), rtol=2e-3, atol=1e-5)
    torch.testing.assert_close(pg.grad, torch.from_numpy(z["grad_prev"]), rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize("name", ["c2", "c2_ln"])
def test_oracle_continuous_gradients_match_reference(name):
    """The differentiable rollout of the port (imagine_st + ac_losses) reproduces the gradients the REFERENCE's
    autograd produced (tests/golden: d loss_actor / d a_t, per-parameter norms and probed entries)."""
    from oracle.gen_golden import grad_probe_indices
    c = load_case(name)
    m, gold = c["meta"], c["gold"]
    r = orc.continuous_update_grads(c["wm"], c["actor"], c["critic"], c["h0"], c["z0"], c["lat"], c["act"], H=m["H"],
                                    A=m["A"], lam=m["lam"], rho=m["rho"], eta=m["entropy_scale"])
    assert torch.equal(r["traj"]["stoch_idx"].to(torch.uint8), gold["stoch_idx"])
    torch.testing.assert_close(r["losses"]["loss_actor"], gold["loss_actor"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(r["g_actions"], gold["grad_actions"], rtol=2e-3, atol=1e-7)
    for i, n in enumerate(m["grad_names"]):
        g = r["grads"][n]
        assert abs(g.norm().item() - gold["grad_norms"][i].item()) <= 1e-3 * gold["grad_norms"][i].item() + 1e-9, n
        torch.testing.assert_close(g.flatten()[grad_probe_indices(g.numel())], gold["grad_probes"][i], rtol=5e-3,
                                   atol=1e-6 * max(1.0, gold["grad_norms"][i].item()), msg=lambda s_: f"{n}: {s_}")


def _load_slotted():
    import json
    import numpy as np
    from oracle.gen_golden import SLOTTED_CASE, slotted_inputs
    from tests._golden import GOLDEN
    z = np.load(GOLDEN / "imagine_slotted.npz")
    m = json.loads(str(z["meta"]))
    assert m == SLOTTED_CASE, "fixture is stale: re-run python -m oracle.gen_golden"
    gold = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if k != "meta"}
    wm, actor, critic = orc.make_params_slotted(m["param_seed"], D=m["D"], A=m["A"], K=m["K"], discrete=m["discrete"],
                                                layer_norm=m["layer_norm"], predict_discount=m["predict_discount"])
    h0, z0, lat, act = slotted_inputs()
    return m, gold, wm, actor, critic, h0, z0, lat, act


def test_oracle_slotted_rollout_matches_reference():
    """oracle_port.imagine_slotted vs the reference's DreamerV2.imagine_trajectory over the slotted world model
    (rssm_slots_attention.py:166-209: mixer blocks, un-mixed determ in the state, pos_enc on the head input)."""
    m, gold, wm, actor, critic, h0, z0, lat, act = _load_slotted()
    out = orc.imagine_slotted(wm, actor, critic, h0, z0, H=m["H"], A=m["A"], K=m["K"], discrete=m["discrete"],
                              predict_discount=m["predict_discount"], latent_uniforms=lat, action_noise=act,
                              blocks=m["blocks"])
    assert torch.equal(out["stoch_idx"], gold["stoch_idx"].long()), "categorical indices must be bit-exact"
    for k in ("determ", "logits", "actions", "rewards", "values"):
        torch.testing.assert_close(out[k], gold[k], rtol=1e-4, atol=3e-5, msg=lambda s: f"{k}: {s}")


def _load_observe():
    import json
    from oracle.gen_golden import OBSERVE_CASE, observe_inputs
    from tests._golden import GOLDEN
    z = np.load(GOLDEN / "observe.npz")
    meta = json.loads(str(z["meta"]))
    assert {k: meta[k] for k in OBSERVE_CASE} == OBSERVE_CASE, "fixture is stale: re-run python -m oracle.gen_golden"
    gold = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if k != "meta"}
    wm, _, _ = orc.make_params(meta["param_seed"], D=meta["D"], A=meta["A"], discrete=False, layer_norm=meta["layer_norm"],
                               predict_discount=False)
    return meta, gold, wm, observe_inputs()


def test_oracle_observe_scan_matches_reference():
    """oracle_port.observe_scan (+ autograd) vs the reference's observe loop (RSSM.forward x T, world_model.py:187-202)
    and the gradients its autograd produced for the seeded probe loss."""
    from oracle.gen_golden import grad_probe_indices
    meta, gold, wm, (embed, actions, uniforms, weights) = _load_observe()
    rp = "recurrent_model."
    wm = {k: (v.clone().requires_grad_() if k.startswith(rp) else v) for k, v in wm.items()}
    embed = embed.clone().requires_grad_()
    o = orc.observe_scan(wm, embed, actions, uniforms)
    assert torch.equal(o["stoch_idx"].to(torch.uint8), gold["stoch_idx"])
    for k in ("prior_logits", "post_logits", "determ"):
        torch.testing.assert_close(o[k], gold[k], rtol=1e-4, atol=2e-5, msg=lambda s: f"{k}: {s}")
    orc.observe_probe_loss(o, weights).backward()
    torch.testing.assert_close(embed.grad, gold["grad_embed"], rtol=2e-3, atol=1e-6)
    for i, n in enumerate(meta["grad_names"]):
        g = wm[rp + n].grad
        nref = gold["grad_norms"][i].item()
        assert abs(g.norm().item() - nref) <= 1e-3 * nref + 1e-8, n
        torch.testing.assert_close(g.flatten()[grad_probe_indices(g.numel())], gold["grad_probes"][i], rtol=5e-3,
                                   atol=1e-5 * max(1.0, nref), msg=lambda s: f"{n}: {s}")
