"""The reference's Hydra configs instantiate the drop-in unchanged: every `_target_` dotted path of
config / config_default / config_dino / config_slotted resolves through the alias package `rl_sandbox.*`, takes the
YAML's kwargs, and yields an agent whose state-dict layout is the reference's.  CPU only: construction, preprocess()
(DINO feature targets for decode_vit configs) and the checkpoint key layout; the GPU tests train these agents."""
import json
import warnings
from pathlib import Path

import pytest
import torch

from tests._hydra_lite import build_agent, instantiate, locate

FIXTURE = Path(__file__).resolve().parent / "golden" / "agent_configs.json"
CONFIGS = ["config", "config_default", "config_dino", "config_slotted"]


@pytest.fixture(scope="module")
def fixture():
    return json.loads(FIXTURE.read_text())


@pytest.fixture(autouse=True)
def offline_hub(monkeypatch):
    """No network here: the DINO checkpoint download fails at once and ViTFeat keeps its random initialisation."""
    def no_network(url, *a, **k):
        raise OSError("offline test environment")
    monkeypatch.setattr(torch.hub, "load_state_dict_from_url", no_network)


def test_fixture_matches_reference_yamls(fixture):
    """The committed fixture is what the reference's YAMLs resolve to (build container only)."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference checkout not present")
    from oracle.gen_config_fixture import build
    assert build() == fixture, "stale fixture: python -m oracle.gen_config_fixture"


def test_every_target_resolves(fixture):
    seen = set()

    def walk(node):
        if isinstance(node, dict):
            if "_target_" in node:
                seen.add(node["_target_"])
            for v in node.values():
                walk(v)
        elif isinstance(node, list):
            for v in node:
                walk(v)
    walk(fixture["agents"])
    assert {"rl_sandbox.agents.DreamerV2", "rl_sandbox.agents.dreamer.world_model.WorldModel",
            "rl_sandbox.agents.dreamer.world_model_slots_attention.WorldModel",
            "rl_sandbox.agents.dreamer.ac.ImaginativeActor", "rl_sandbox.agents.dreamer.ac.ImaginativeCritic",
            "rl_sandbox.utils.optimizer.Optimizer", "rl_sandbox.utils.optimizer.WarmupScheduler"} <= seen
    for path in sorted(seen):
        obj = locate(path)
        assert obj.__module__.startswith("rl_sandbox_b200."), (path, obj.__module__)


@pytest.mark.parametrize("name", CONFIGS)
def test_config_instantiates(fixture, name):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        agent = build_agent(fixture, name, "cpu", f16_precision=False)
    top, cfg = fixture["top"][name], fixture["agents"][name]
    wm = agent.world_model
    assert agent.imagination_horizon == cfg["imagination_horizon"] == 15
    assert wm.rssm_dim == cfg["world_model"]["rssm_dim"] and wm.cluster_size == cfg["batch_cluster_size"]
    assert agent.is_discrete == (top["action_type"] == "discrete")
    slotted = "slots_num" in cfg["world_model"]
    assert wm.state_size == (4 if slotted else 1) * (wm.rssm_dim + 1024)
    assert agent.actor.rho == (1.0 if agent.is_discrete else 0.0)          # reinforce_fraction: null (ac.py:90-93)
    assert agent.critic.lambda_ == 0.95
    for opt, key in ((agent.world_model_optimizer, "wm_optim"), (agent.actor_optimizer, "actor_optim")):
        g = opt.optimizer.param_groups[0]
        assert g["eps"] == 1e-5 and g["weight_decay"] == 1e-6 and opt.clip == 100
        assert g["initial_lr" if "initial_lr" in g else "lr"] == pytest.approx(cfg[key]["lr"])
    if cfg["wm_optim"].get("lr_scheduler"):
        assert isinstance(agent.world_model_optimizer.lr_scheduler, torch.optim.lr_scheduler.ChainedScheduler)
    # the reference wraps nothing else: every trainable parameter belongs to exactly one optimizer
    n_opt = sum(len(g["params"]) for o in (agent.world_model_optimizer, agent.actor_optimizer, agent.critic_optimizer)
                for g in o.optimizer.param_groups)
    n_mod = sum(1 for m in (agent.world_model, agent.actor, agent.critic) for _ in m.parameters())
    assert n_opt == n_mod
    state = agent.world_model.get_initial_state()
    state = state[0] if isinstance(state, tuple) else state
    assert state.determ.shape[-1] == wm.rssm_dim


@pytest.mark.parametrize("name", ["config_dino", "config_slotted"])
def test_decode_vit_preprocess(fixture, name):
    """decode_vit: true (config_slotted as shipped; config_dino with the override BASELINE configs[1] describes):
    preprocess() attaches the frozen ViT's key features (B, 384, 196) to the rollout (dreamer_v2.py:103-111)."""
    from rl_sandbox.utils.replay_buffer import Rollout
    over = {}
    if name == "config_dino":
        wm_node = dict(fixture["agents"][name]["world_model"], decode_vit=True)
        over["world_model"] = instantiate(wm_node)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        agent = build_agent(fixture, name, "cpu", **over)
    assert any("ViTFeat keeps its random initialisation" in str(w.message) for w in rec)
    wm = agent.world_model
    assert wm.decode_vit and not any(p.requires_grad for p in wm.dino_vit.parameters())
    assert [n for n, _ in wm.named_children()][:2] == ["recurrent_model", "dino_vit"]     # reference order
    T = 3
    g = torch.Generator().manual_seed(0)
    A = fixture["top"][name]["actions_num"]
    ro = Rollout(obs=torch.randint(0, 256, (T, 64, 64, 3), generator=g, dtype=torch.uint8),
                 actions=torch.randn(T, A, generator=g), rewards=torch.randn(T, generator=g),
                 is_finished=torch.zeros(T), is_first=torch.zeros(T))
    out = agent.preprocess(ro)
    assert out.obs.shape == (T, 3, 64, 64) and out.obs.min() >= -0.5 and out.obs.max() <= 0.5
    d = out.additional_data["d_features"]
    assert d.shape == (T, 384, 196) and torch.isfinite(d).all() and not d.requires_grad
    # optimizer excludes nothing it should hold, and the frozen ViT receives no update
    sd = wm.state_dict()
    assert "dino_vit.model.blocks.11.attn.qkv.weight" in sd and "dino_predictor.convin.weight" in sd


def test_decode_vit_rejects_unknown_size(fixture):
    node = dict(fixture["agents"]["config_dino"]["world_model"], decode_vit=True, vit_img_size=128)
    with pytest.raises(RuntimeError, match="Unknown vit img size"):
        instantiate(node)(actions_num=12)
