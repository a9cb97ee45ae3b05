"""check_host_mirror.py — TEST INFRASTRUCTURE (build container only: needs /root/reference).

Builds the reference agent and the rl_sandbox_b200 agent with the same constructor kwargs, copies
the reference's parameters into ours through state_dict (=> key / shape compatibility) and compares
  * WorldModel.calculate_loss (observe loop + losses), same torch seed  -> identical losses
  * ImaginativeCritic.calculate_loss / ImaginativeActor.calculate_loss on a reference trajectory
Prints one line per check and exits non-zero on failure.  Run:  python -m oracle.check_host_mirror
"""
import sys
from functools import partial

import torch

from . import ref_harness as rh


def check_slotted():
    """world_model_slots_attention.WorldModel + rssm_slots_attention.RSSM mirrors vs the reference (config_slotted dims)."""
    rh._import_reference()
    from rl_sandbox.agents.dreamer.world_model_slots_attention import WorldModel as RefWM
    from rl_sandbox_b200.agents.dreamer.world_model_slots_attention import WorldModel as MyWM
    kw = dict(batch_cluster_size=4, latent_dim=32, latent_classes=32, rssm_dim=200, actions_num=1,
              discount_loss_scale=1.0, kl_loss_scale=1000, kl_loss_balancing=0.8, kl_free_nats=5e-4,
              discrete_rssm=False, predict_discount=False, layer_norm=True, encode_vit=False, decode_vit=False,
              vit_l2_ratio=0.75, vit_img_size=224, slots_num=4, slots_iter_num=2, use_prev_slots=False)
    torch.manual_seed(0)
    ref, mine = RefWM(**kw), MyWM(**kw)
    sd = ref.state_dict()
    mk, rk = set(mine.state_dict()), set(sd)
    ok = mk == rk and all(mine.state_dict()[k].shape == sd[k].shape for k in rk)
    print(f"[slotted] state_dict keys/shapes world_model: {'ok' if ok else 'MISMATCH ' + str(sorted(mk ^ rk)[:8])} ({len(rk)} entries)")
    if not ok:
        return False
    mine.load_state_dict(sd)
    B, T = 2, 4
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(B * T, 3, 64, 64, generator=g) - 0.5
    a, r = torch.randn(B * T, 1, generator=g), torch.randn(B * T, generator=g)
    disc, first = 0.99 * torch.ones(B * T), torch.zeros(B * T)
    first[0] = 1
    torch.manual_seed(5)
    l_ref, post_ref, m_ref = ref.calculate_loss(obs, a, r, disc, first, {})
    torch.manual_seed(5)
    l_mine, post_mine, m_mine = mine.calculate_loss(obs, a, r, disc, first, {})
    for k in l_ref:
        same = torch.allclose(l_ref[k].float().reshape(-1), l_mine[k].float().reshape(-1), rtol=1e-5, atol=1e-6)
        print(f"[slotted] calculate_loss {k}: ref {float(l_ref[k]):.6f} ours {float(l_mine[k]):.6f} {'ok' if same else 'MISMATCH'}")
        ok &= same
    for k in m_ref:
        ok &= torch.allclose(torch.as_tensor(m_ref[k]).float(), torch.as_tensor(m_mine[k]).float(), rtol=1e-5, atol=1e-6)
    same = torch.allclose(post_ref.determ, post_mine.determ, atol=1e-6) and torch.equal(post_ref.stoch, post_mine.stoch)
    print(f"[slotted] posterior states: {'ok' if same else 'MISMATCH'}")
    ok &= same
    act = torch.randn(1, B * T, 1)
    torch.manual_seed(6)
    pr, rr, dr = ref.predict_next(post_ref.flatten().detach(), act)
    torch.manual_seed(6)
    pm, rm, dm = mine.predict_next(post_mine.flatten().detach(), act)
    same = (torch.allclose(pr.determ, pm.determ, atol=1e-6) and torch.allclose(pr.stoch_logits, pm.stoch_logits, atol=1e-5)
            and torch.allclose(rr, rm, atol=1e-6) and torch.allclose(pr.determ_updated, pm.determ_updated, atol=1e-6))
    diffs = [(pr.determ - pm.determ).abs().max().item(), (pr.stoch_logits - pm.stoch_logits).abs().max().item(),
             (pr.determ_updated - pm.determ_updated).abs().max().item(), (rr - rm).abs().max().item()]
    print(f"[slotted] predict_next (determ, logits, determ_updated, reward) max abs diff {diffs}: {'ok' if same else 'MISMATCH'}")
    return ok and same


def _patched_hub():
    """torch.hub.load_state_dict_from_url -> seeded random ViT-S weights built by the REFERENCE's vit_small (no network;
    DINO is frozen and only supplies loss targets)."""
    import contextlib

    @contextlib.contextmanager
    def ctx():
        from rl_sandbox.vision.dino import vit_small   # the reference's (ref_harness imported it first)
        orig = torch.hub.load_state_dict_from_url

        def fake(url, *a, **k):
            g = torch.Generator().manual_seed(1234)
            sd = vit_small(patch_size=8 if "small8" in url else 16).state_dict()
            return {k_: (torch.randn(v.shape, generator=g) * 0.05 + (1.0 if k_.endswith("norm1.weight") or k_.endswith("norm2.weight") or k_ == "norm.weight" else 0.0))
                    for k_, v in sd.items()}
        torch.hub.load_state_dict_from_url = fake
        try:
            yield
        finally:
            torch.hub.load_state_dict_from_url = orig
    return ctx()


def check_dino():
    """decode_vit=True (config_dino / config_slotted as shipped): ViTFeat, SpatialBroadcastDecoder, precalc_data and the
    DINO reconstruction loss of both world models vs the reference; parameter ORDER (optimizer-state indices) too."""
    rh._import_reference()
    from rl_sandbox.agents.dreamer.world_model import WorldModel as RefFlat
    from rl_sandbox.agents.dreamer.world_model_slots_attention import WorldModel as RefSlot
    from rl_sandbox_b200.agents.dreamer.world_model import WorldModel as MyFlat
    from rl_sandbox_b200.agents.dreamer.world_model_slots_attention import WorldModel as MySlot
    flat_kw = dict(batch_cluster_size=4, latent_dim=32, latent_classes=32, rssm_dim=200, actions_num=12,
                   discount_loss_scale=1.0, kl_loss_scale=2, kl_loss_balancing=0.8, kl_free_nats=1.0, discrete_rssm=False,
                   predict_discount=False, layer_norm=False, encode_vit=False, decode_vit=True, vit_l2_ratio=0.5,
                   vit_img_size=224)
    slot_kw = dict(batch_cluster_size=4, latent_dim=32, latent_classes=32, rssm_dim=200, actions_num=1,
                   discount_loss_scale=1.0, kl_loss_scale=1000, kl_loss_balancing=0.8, kl_free_nats=5e-4,
                   discrete_rssm=False, predict_discount=False, layer_norm=True, encode_vit=False, decode_vit=True,
                   vit_l2_ratio=0.75, vit_img_size=224, slots_num=4, slots_iter_num=2, use_prev_slots=False)
    ok = True
    for tag, Ref, Mine, kw, A in (("dino-flat", RefFlat, MyFlat, flat_kw, 12), ("dino-slotted", RefSlot, MySlot, slot_kw, 1),
                                  ("dino-slotted-spatial", RefSlot, MySlot, dict(slot_kw, spatial_decoder=True), 1),
                                  ("dino-flat-64", RefFlat, MyFlat, dict(flat_kw, vit_img_size=64), 12)):
        with _patched_hub():
            torch.manual_seed(0)
            ref = Ref(**kw)
            mine = Mine(**kw)
        sd = ref.state_dict()
        names_ref = [n for n, _ in ref.named_parameters()]
        names_mine = [n for n, _ in mine.named_parameters()]
        same = names_ref == names_mine and list(sd) == list(mine.state_dict()) and \
            all(mine.state_dict()[k].shape == sd[k].shape for k in sd)
        print(f"[{tag}] parameter order + state_dict keys/shapes: {'ok' if same else 'MISMATCH'} ({len(sd)} entries, "
              f"{len(names_ref)} parameters)")
        ok &= same
        if not same:
            bad = [(a, b) for a, b in zip(names_ref, names_mine) if a != b][:4]
            print("   first differences:", bad, sorted(set(sd) ^ set(mine.state_dict()))[:6])
            continue
        mine.load_state_dict(sd)
        frozen = all(not p.requires_grad for p in mine.dino_vit.parameters())
        print(f"[{tag}] dino_vit frozen: {'ok' if frozen else 'MISMATCH'}")
        ok &= frozen
        if kw["vit_img_size"] == 64:
            # the reference's 64-pixel branch builds a 14 x 14 decoder for 8 x 8 features: construction + features only
            obs = torch.rand(3, 3, 64, 64) - 0.5
            d_ref, d_mine = ref.precalc_data(obs)["d_features"], mine.precalc_data(obs)["d_features"]
            same = d_ref.shape == d_mine.shape and torch.allclose(d_ref, d_mine, rtol=1e-4, atol=1e-5)
            print(f"[{tag}] precalc_data d_features {tuple(d_ref.shape)} max abs diff {(d_ref - d_mine).abs().max():.2e}: "
                  f"{'ok' if same else 'MISMATCH'}")
            ok &= same
            continue
        B, T = 2, 4
        g = torch.Generator().manual_seed(1)
        obs = torch.rand(B * T, 3, 64, 64, generator=g) - 0.5
        a, r = torch.randn(B * T, A, generator=g), torch.randn(B * T, generator=g)
        disc, first = 0.99 * torch.ones(B * T), torch.zeros(B * T)
        first[0] = 1
        d_ref, d_mine = ref.precalc_data(obs)["d_features"], mine.precalc_data(obs)["d_features"]
        same = d_ref.shape == d_mine.shape and d_ref.device == d_mine.device and \
            torch.allclose(d_ref, d_mine, rtol=1e-4, atol=1e-5)
        print(f"[{tag}] precalc_data d_features {tuple(d_ref.shape)} max abs diff {(d_ref - d_mine).abs().max():.2e}: "
              f"{'ok' if same else 'MISMATCH'}")
        ok &= same
        torch.manual_seed(5)
        l_ref, post_ref, m_ref = ref.calculate_loss(obs, a, r, disc, first, {"d_features": d_ref})
        torch.manual_seed(5)
        l_mine, post_mine, m_mine = mine.calculate_loss(obs, a, r, disc, first, {"d_features": d_ref})
        for k in l_ref:
            same = torch.allclose(l_ref[k].float().reshape(-1), l_mine[k].float().reshape(-1), rtol=1e-5, atol=1e-6)
            print(f"[{tag}] calculate_loss {k}: ref {float(l_ref[k]):.6f} ours {float(l_mine[k]):.6f} {'ok' if same else 'MISMATCH'}")
            ok &= same
        for k in m_ref:
            same = torch.allclose(torch.as_tensor(m_ref[k]).float(), torch.as_tensor(m_mine[k]).float(), rtol=1e-5, atol=1e-6)
            if not same:
                print(f"[{tag}] metric {k}: MISMATCH")
            ok &= same
    return ok


def main():
    ok = True
    for cfg in (dict(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True),
                dict(D=200, A=12, discrete=False, layer_norm=False, predict_discount=False)):
        torch.manual_seed(0)
        ref = rh.build_agent(**cfg, H=3, batch_cluster_size=4)
        from rl_sandbox_b200.agents.dreamer_v2 import DreamerV2
        from rl_sandbox_b200.agents.dreamer.world_model import WorldModel
        from rl_sandbox_b200.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
        from rl_sandbox_b200.utils.optimizer import Optimizer
        ln = cfg["layer_norm"]
        mine = DreamerV2(
            obs_space_num=[64, 64, 3], clip_rewards="identity", actions_num=cfg["A"],
            world_model=partial(WorldModel, batch_cluster_size=4, latent_dim=32, latent_classes=32, rssm_dim=cfg["D"],
                                discount_loss_scale=1.0, kl_loss_scale=2, kl_loss_balancing=0.8, kl_free_nats=1.0,
                                discrete_rssm=False, predict_discount=cfg["predict_discount"], layer_norm=ln,
                                encode_vit=False, decode_vit=False, vit_l2_ratio=0.5, vit_img_size=224),
            actor=partial(ImaginativeActor, layer_norm=ln, reinforce_fraction=None, entropy_scale=1e-5),
            critic=partial(ImaginativeCritic, discount_factor=0.99, update_interval=100, soft_update_fraction=1,
                           value_target_lambda=0.95, layer_norm=ln),
            action_type="discrete" if cfg["discrete"] else "continuous", imagination_horizon=3,
            wm_optim=partial(Optimizer, lr=1e-4, eps=1e-5, weight_decay=1e-6, clip=100),
            actor_optim=partial(Optimizer, lr=1e-4, eps=1e-5, weight_decay=1e-6, clip=100),
            critic_optim=partial(Optimizer, lr=1e-4, eps=1e-5, weight_decay=1e-6, clip=100),
            layer_norm=ln, batch_cluster_size=4, f16_precision=False, device_type="cpu")
        strip = lambda sd: {k.removeprefix("_orig_mod."): v for k, v in sd.items()}
        for name, a, b in (("world_model", ref.world_model, mine.world_model), ("actor", ref.actor, mine.actor),
                           ("critic", ref.critic, mine.critic)):
            sd = strip(a.state_dict())
            mk, rk = set(b.state_dict()), set(sd)
            same = mk == rk and all(b.state_dict()[k].shape == sd[k].shape for k in rk)
            print(f"[{cfg['D']}] state_dict keys/shapes {name}: {'ok' if same else 'MISMATCH ' + str(sorted(mk ^ rk)[:6])} ({len(rk)} entries)")
            ok &= same
            b.load_state_dict(sd)
        B, T, A = 2, 4, cfg["A"]
        g = torch.Generator().manual_seed(1)
        obs = torch.rand(B * T, 3, 64, 64, generator=g) - 0.5
        a = (torch.nn.functional.one_hot(torch.randint(0, A, (B * T,), generator=g), A).float() if cfg["discrete"]
             else torch.randn(B * T, A, generator=g))
        r = torch.randn(B * T, generator=g)
        disc = 0.99 * torch.ones(B * T)
        first = torch.zeros(B * T)
        first[0] = 1
        torch.manual_seed(5)
        l_ref, post_ref, m_ref = ref.world_model.calculate_loss(obs, a, r, disc, first, {})
        torch.manual_seed(5)
        l_mine, post_mine, m_mine = mine.world_model.calculate_loss(obs, a, r, disc, first, {})
        for k in l_ref:
            same = torch.allclose(l_ref[k].float().reshape(-1), l_mine[k].float().reshape(-1), rtol=1e-5, atol=1e-6)
            print(f"[{cfg['D']}] calculate_loss {k}: ref {float(l_ref[k]):.6f} ours {float(l_mine[k]):.6f} {'ok' if same else 'MISMATCH'}")
            ok &= same
        same = torch.allclose(post_ref.determ, post_mine.determ, atol=1e-6) and torch.equal(post_ref.stoch, post_mine.stoch)
        print(f"[{cfg['D']}] posterior states: {'ok' if same else 'MISMATCH'}")
        ok &= same
        for k in m_ref:
            ok &= torch.allclose(m_ref[k].float(), m_mine[k].float(), rtol=1e-5, atol=1e-6)
        # AC losses on a (CPU) reference trajectory: our torch loss code vs the reference's
        H, N = 3, B * T
        zs = torch.randn(H + 1, N, cfg["D"] + 1024, generator=g)
        vs = torch.randn(H, N, 1, generator=g)
        w = torch.rand(H + 1, N, 1, generator=g)
        acts = (torch.nn.functional.one_hot(torch.randint(0, A, (H - 1, N), generator=g), A).float() if cfg["discrete"]
                else torch.randn(H - 1, N, A, generator=g))
        base = torch.randn(H - 1, N, 1, generator=g)
        lc_r, _ = ref.critic.calculate_loss(zs[:-1], vs, w[:-1])
        lc_m, _ = mine.critic.calculate_loss(zs[:-1], vs, w[:-1])
        torch.manual_seed(9)
        la_r, ma_r = ref.actor.calculate_loss(zs[:-2], vs[1:], base, w[:-2], acts)
        torch.manual_seed(9)
        la_m, ma_m = mine.actor.calculate_loss(zs[:-2], vs[1:], base, w[:-2], acts)
        for k, x, y in [("loss_critic", lc_r["loss_critic"], lc_m["loss_critic"])] + [(k, la_r[k], la_m[k]) for k in la_r] + \
                       [(k, ma_r[k], ma_m[k]) for k in ma_r]:
            same = torch.allclose(torch.as_tensor(x).float(), torch.as_tensor(y).float(), rtol=1e-5, atol=1e-6)
            print(f"[{cfg['D']}] {k}: ref {float(x):.6f} ours {float(y):.6f} {'ok' if same else 'MISMATCH'}")
            ok &= same
    ok &= check_slotted()
    ok &= check_dino()
    print("HOST MIRROR", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
