"""GPU parity of K5 (rlsb_observe_fwd / rlsb_observe_bwd): the world-model observe scan and its BPTT."""
import pytest
import torch

from oracle import oracle_port as orc
from tests.test_oracle import _load_observe

pytestmark = pytest.mark.gpu
RP = "recurrent_model."


def _rel(a, b):
    return ((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt().clamp_min(1e-30)).item()


def _run_kernel(wm, embed, actions, uniforms, weights, D, A, E, ln):
    from rl_sandbox_b200 import ops
    T = embed.shape[0]
    eng = ops.ObserveEngine(D, A, E, ln, T)
    eng.pack({k[len(RP):]: v.cuda() for k, v in wm.items() if k.startswith(RP)})
    out = eng.forward(embed.cuda(), actions.cuda(), latent_uniforms=uniforms.cuda())
    g_embed, grads = eng.backward(out, weights["prior_logits"].cuda(), weights["post_logits"].cuda(),
                                  weights["determ"].cuda(), weights["stoch"].cuda())
    torch.cuda.synchronize()
    return out, g_embed, grads


def test_observe_scan_matches_reference(cuda):
    from oracle.gen_golden import grad_probe_indices
    meta, gold, wm, (embed, actions, uniforms, weights) = _load_observe()
    out, g_embed, grads = _run_kernel(wm, embed, actions, uniforms, weights, meta["D"], meta["A"], meta["E"], meta["layer_norm"])
    own = orc.sample_categorical(out["post_logits"].cpu().view(meta["T"], meta["B"], 32, 32), uniforms.view(meta["T"], meta["B"], 32, 32))
    assert torch.equal(own, out["stoch_idx"].cpu().long()), "posterior indices must be bit-exact for the kernel's logits"
    same = (out["stoch_idx"].cpu() == gold["stoch_idx"]).all(-1).cumprod(0).bool()   # (T, B): identical draws so far
    print(f"[parity] observe: sequences with identical posterior draws over all {meta['T']} steps: {same[-1].float().mean().item():.2f}")
    for k, tol in (("determ", 4e-3), ("prior_logits", 1e-2), ("post_logits", 1e-2)):
        e = _rel(out[k].cpu()[same], gold[k][same])
        print(f"[parity] observe {k} vs reference: rel-RMS {e:.3e}")
        assert e < tol, (k, e)
    if same[-1].all():
        e = _rel(g_embed.cpu(), gold["grad_embed"])
        print(f"[parity] observe d loss / d embed vs reference autograd: rel-RMS {e:.3e}")
        assert e < 3e-2
        worst = 0.0
        for i, n in enumerate(meta["grad_names"]):
            g = grads[n].cpu()
            nref = gold["grad_norms"][i].item()
            pr = gold["grad_probes"][i]
            perr = ((g.flatten()[grad_probe_indices(g.numel())] - pr).norm() / pr.norm().clamp_min(1e-12)).item()
            nerr = abs(g.norm().item() - nref) / max(nref, 1e-12)
            print(f"[parity] observe grad {n}: |g| ours {g.norm().item():.4e} ref {nref:.4e} probes rel-L2 {perr:.3e}")
            worst = max(worst, perr)
            assert nerr < 3e-2 and perr < 8e-2, (n, nerr, perr)
        print(f"[parity] observe: worst probed parameter-gradient error vs reference {worst:.3e}")


@pytest.mark.parametrize("D,ln,B,T", [(1024, True, 16, 6), (200, False, 5, 4)])
def test_observe_scan_matches_bf16_oracle(cuda, D, ln, B, T):
    """config-1 dims (LayerNorm over 1024 columns = 4 n-blocks) and config-2 dims (no LayerNorm), ragged B, vs the
    oracle's autograd with bf16-rounded contraction operands"""
    A, E = 7, 1536
    wm, _, _ = orc.make_params(3, D=D, A=A, discrete=False, layer_norm=ln, predict_discount=False)
    g = torch.Generator().manual_seed(4)
    embed, actions, uniforms = torch.randn(T, B, E, generator=g), torch.randn(T, B, A, generator=g), torch.rand(T, B, 1024, generator=g)
    weights = {k: torch.randn(T, B, n, generator=g) / n ** 0.5
               for k, n in (("prior_logits", 1024), ("post_logits", 1024), ("determ", D), ("stoch", 1024))}
    wm_g = {k: (v.clone().requires_grad_() if k.startswith(RP) else v) for k, v in wm.items()}
    e_g = embed.clone().requires_grad_()
    ref = orc.observe_scan(wm_g, e_g, actions, uniforms, bf16=True)
    orc.observe_probe_loss(ref, weights).backward()
    out, g_embed, grads = _run_kernel(wm, embed, actions, uniforms, weights, D, A, E, ln)
    same = (out["stoch_idx"].cpu().long() == ref["stoch_idx"]).all(-1).all(0)
    print(f"[parity] observe D={D} ln={ln}: sequences with identical draws {same.float().mean().item():.2f}")
    for k, tol in (("determ", 1e-3), ("prior_logits", 3e-3), ("post_logits", 3e-3)):
        e = _rel(out[k].cpu()[:, same], ref[k].detach()[:, same])
        print(f"[parity] observe D={D} {k} vs bf16 oracle: rel-RMS {e:.3e}")
        assert e < tol, (k, e)
    if same.all():
        e = _rel(g_embed.cpu(), e_g.grad)
        print(f"[parity] observe D={D} d/d embed: rel-RMS {e:.3e}")
        assert e < 3e-2
        for n, gk in grads.items():
            r = _rel(gk.cpu(), wm_g[RP + n].grad)
            print(f"[parity] observe D={D} d/d {n}: rel-RMS {r:.3e}")
            assert r < 4e-2, (n, r)


def test_world_model_loss_uses_observe_scan(cuda):
    """WorldModel.calculate_loss through the alias package: the observe loop runs in K5 under torch autograd; the
    noise-free first step agrees with the torch-op loop, every loss is finite, and gradients reach the encoder and
    every RSSM parameter."""
    from rl_sandbox.agents.dreamer.world_model import WorldModel
    torch.manual_seed(0)
    wm = WorldModel(batch_cluster_size=6, latent_dim=32, latent_classes=32, rssm_dim=200, actions_num=5,
                    discount_loss_scale=1.0, kl_loss_scale=2, kl_loss_balancing=0.8, kl_free_nats=1.0, discrete_rssm=False,
                    predict_discount=True, layer_norm=True, encode_vit=False, decode_vit=False, vit_l2_ratio=0.5,
                    vit_img_size=224).cuda()
    B, T = 3, 6
    g = torch.Generator().manual_seed(1)
    obs = (torch.rand(B * T, 3, 64, 64, generator=g) - 0.5).cuda()
    a = torch.nn.functional.one_hot(torch.randint(0, 5, (B * T,), generator=g), 5).float().cuda()
    r, disc, first = torch.randn(B * T, generator=g).cuda(), 0.99 * torch.ones(B * T).cuda(), torch.zeros(B * T).cuda()
    first[0] = 1
    res = {}
    for mode in (True, False):
        wm.kernel_observe = mode
        for p in wm.parameters():
            p.grad = None
        losses, post, metrics = wm.calculate_loss(obs, a, r, disc, first, {})
        losses["loss_wm"].backward()
        assert all(torch.isfinite(v).all() for v in losses.values())
        res[mode] = (losses, post, {n: p.grad.clone() for n, p in wm.named_parameters() if p.grad is not None})
    assert wm._observe_calls == 1
    e = _rel(res[True][1].determ[0], res[False][1].determ[0])
    print(f"[parity] world model: first-step determ, K5 vs torch loop: rel-RMS {e:.3e}")
    assert e < 6e-3   # two bf16 contractions deep, against torch fp32
    for n in list(res[False][2]):
        if n.startswith("recurrent_model.") and ("determ_discretizer" not in n and "determ_layer_norm" not in n):
            assert n in res[True][2] and res[True][2][n].abs().sum() > 0, n
    assert res[True][2]["encoder.net.0.weight"].abs().sum() > 0
    # same order of magnitude as the torch path (posterior draws differ: Philox vs torch RNG)
    for k in ("loss_kl_reg", "loss_reward_pred", "loss_reconstruction"):
        a_, b_ = res[True][0][k].item(), res[False][0][k].item()
        print(f"[parity] world model {k}: K5 path {a_:.4f} torch path {b_:.4f}")
        assert abs(a_ - b_) <= 0.2 * abs(b_) + 1e-3
