"""ref_harness.py — TEST INFRASTRUCTURE ONLY; works only where /root/reference exists (the build
container).  Imports the UNMODIFIED reference (Midren/rl_sandbox) with three import shims
(oracle/shims) and a noise-injection seam, so that the reference's own PyTorch modules can be run
on caller-supplied noise and compared with the oracle port / the CUDA path.

Noise seam: the reference draws through torch's global generator
(aten::multinomial inside OneHotCategorical.sample, aten::normal_ inside Normal.rsample).  We
replace, for the duration of a `with injected_noise(...)` block,
  * ``torch.distributions.OneHotCategoricalStraightThrough`` by a subclass whose ``sample()`` is
    the Gumbel-max draw on the RAW logits with supplied uniforms (oracle_port.sample_categorical),
  * ``torch.distributions.normal._standard_normal`` by a function returning supplied normals.
Everything else that runs is the reference's code.
"""
from __future__ import annotations

import contextlib
import os
import sys
from functools import partial
from pathlib import Path

import torch
import torch.distributions as td

_HERE = Path(__file__).resolve().parent


def _find_reference() -> Path:
    """RLSB_REFERENCE_ROOT, else the read-only checkout of the build container, else the copy oracle/make_ref.py
    stages under oracle/_ref/ (git-ignored; what the GPU box has)."""
    env = os.environ.get("RLSB_REFERENCE_ROOT")
    if env:
        return Path(env)
    for cand in (Path("/root/reference"), _HERE / "_ref"):
        if (cand / "rl_sandbox" / "agents" / "dreamer_v2.py").exists():
            return cand
    return Path("/root/reference")


REFERENCE_ROOT = _find_reference()


def available() -> bool:
    return (REFERENCE_ROOT / "rl_sandbox" / "agents" / "dreamer_v2.py").exists()


def _import_reference():
    """Import the reference package under its own name `rl_sandbox` (shadowing this repo's alias
    package of the same name, which must not be imported in the same process)."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (os.fspath(_HERE / "shims"), os.fspath(REFERENCE_ROOT)):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in [m for m in sys.modules if m == "rl_sandbox" or m.startswith("rl_sandbox.")]:
        mod = sys.modules[name]
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(os.fspath(REFERENCE_ROOT)):
            del sys.modules[name]
    td.Distribution.set_default_validate_args(False)  # as train.py:38
    import rl_sandbox.agents.dreamer_v2 as ref_dv2  # noqa
    return ref_dv2


class NoiseQueue:
    def __init__(self, latent_uniforms=None, action_noise=None, groups=32, classes=32):
        self.latent = list(latent_uniforms) if latent_uniforms is not None else []
        self.action = list(action_noise) if action_noise is not None else []
        self.groups, self.classes = groups, classes
        self.log = []

    def next_uniform(self, shape):
        if len(shape) >= 2 and tuple(shape[-2:]) == (self.groups, self.classes):
            self.log.append("latent")
            return self.latent.pop(0).reshape(shape)
        self.log.append("action")
        return self.action.pop(0).reshape(shape)

    def next_normal(self, shape):
        self.log.append("action_normal")
        return self.action.pop(0).reshape(shape)


@contextlib.contextmanager
def injected_noise(queue: NoiseQueue):
    from . import oracle_port as orc
    import torch.distributions.normal as tdn
    orig_cls = td.OneHotCategoricalStraightThrough
    orig_norm = tdn._standard_normal

    class InjectedST(orig_cls):
        def __init__(self, probs=None, logits=None, validate_args=None):
            super().__init__(probs=probs, logits=logits, validate_args=validate_args)
            self._raw_logits = logits

        def sample(self, sample_shape=torch.Size()):
            if len(sample_shape) != 0 or self._raw_logits is None:
                return super().sample(sample_shape)   # metrics-only draws (ac.py:137)
            raw = self._raw_logits.detach()
            u = queue.next_uniform(raw.shape)
            idx = orc.sample_categorical(raw, u)
            return torch.nn.functional.one_hot(idx, raw.shape[-1]).to(raw.dtype)

    def fake_standard_normal(shape, dtype, device):
        if len(shape) == 3 and queue.action:
            return queue.next_normal(shape).to(dtype)
        return orig_norm(shape, dtype=dtype, device=device)

    td.OneHotCategoricalStraightThrough = InjectedST
    tdn._standard_normal = fake_standard_normal
    try:
        yield queue
    finally:
        td.OneHotCategoricalStraightThrough = orig_cls
        tdn._standard_normal = orig_norm


def build_agent(*, D, A, discrete, layer_norm, predict_discount, H=15, entropy_scale=1e-5, lam=0.95,
                gamma=0.99, lr=1e-4, batch_cluster_size=50, clip_rewards="identity", device_type="cpu"):
    """The reference DreamerV2 with the kwargs of config/agent/dreamer_v2*.yaml."""
    ref = _import_reference()
    from rl_sandbox.agents.dreamer.world_model import WorldModel
    from rl_sandbox.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
    from rl_sandbox.utils.optimizer import Optimizer
    wm = partial(WorldModel, batch_cluster_size=batch_cluster_size, latent_dim=32, latent_classes=32, rssm_dim=D,
                 discount_loss_scale=1.0, kl_loss_scale=2, kl_loss_balancing=0.8, kl_free_nats=1.0,
                 discrete_rssm=False, predict_discount=predict_discount, layer_norm=layer_norm, encode_vit=False,
                 decode_vit=False, vit_l2_ratio=0.5, vit_img_size=224)
    actor = partial(ImaginativeActor, layer_norm=layer_norm, reinforce_fraction=None, entropy_scale=entropy_scale)
    critic = partial(ImaginativeCritic, discount_factor=gamma, update_interval=100, soft_update_fraction=1,
                     value_target_lambda=lam, layer_norm=layer_norm)
    opt = partial(Optimizer, lr=lr, eps=1e-5, weight_decay=1e-6, clip=100)
    agent = ref.DreamerV2(obs_space_num=[64, 64, 3], clip_rewards=clip_rewards, actions_num=A,
                          world_model=wm, actor=actor, critic=critic,
                          action_type="discrete" if discrete else "continuous", imagination_horizon=H,
                          wm_optim=opt, actor_optim=opt, critic_optim=opt, layer_norm=layer_norm,
                          batch_cluster_size=batch_cluster_size, f16_precision=False, device_type=device_type)
    return agent


def build_agent_slotted(*, D, A, K, discrete, layer_norm, predict_discount, H=15, entropy_scale=1e-4, lam=0.95,
                        gamma=0.999, lr=8e-5, batch_cluster_size=50, attention_block_num=3):
    """The reference DreamerV2 over world_model_slots_attention.WorldModel (config/agent/dreamer_v2_slotted_debug.yaml,
    plus the two arguments that YAML forgets; decode_vit off: DINO only shapes the world-model loss)."""
    ref = _import_reference()
    from rl_sandbox.agents.dreamer.world_model_slots_attention import WorldModel
    from rl_sandbox.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
    from rl_sandbox.utils.optimizer import Optimizer
    wm = partial(WorldModel, batch_cluster_size=batch_cluster_size, latent_dim=32, latent_classes=32, rssm_dim=D,
                 discount_loss_scale=1.0, kl_loss_scale=1000, kl_loss_balancing=0.8, kl_free_nats=5e-4,
                 discrete_rssm=False, predict_discount=predict_discount, layer_norm=layer_norm, encode_vit=False,
                 decode_vit=False, vit_l2_ratio=0.75, vit_img_size=224, slots_num=K, slots_iter_num=2,
                 use_prev_slots=False, attention_block_num=attention_block_num)
    actor = partial(ImaginativeActor, layer_norm=layer_norm, reinforce_fraction=None, entropy_scale=entropy_scale)
    critic = partial(ImaginativeCritic, discount_factor=gamma, update_interval=100, soft_update_fraction=1,
                     value_target_lambda=lam, layer_norm=layer_norm)
    opt = partial(Optimizer, lr=lr, eps=1e-5, weight_decay=1e-6, clip=100)
    return ref.DreamerV2(obs_space_num=[64, 64, 3], clip_rewards="identity", actions_num=A, world_model=wm,
                         actor=actor, critic=critic, action_type="discrete" if discrete else "continuous",
                         imagination_horizon=H, wm_optim=opt, actor_optim=opt, critic_optim=opt, layer_norm=layer_norm,
                         batch_cluster_size=batch_cluster_size, f16_precision=False, device_type="cpu")


def ref_state_slotted(agent, h0, z0):
    """h0 (N,K,D), z0 (N,K,1024) -> the reference's slotted State (1, N, K, .) carrying the world model's pos_enc."""
    from rl_sandbox.agents.dreamer.rssm_slots_attention import State
    N, K = h0.shape[:2]
    wm = getattr(agent.world_model, "_orig_mod", agent.world_model)
    return State(h0.unsqueeze(0).clone(), torch.zeros(1, N, K, 32, 32), z0.unsqueeze(0).clone(),
                 wm.pos_enc.unsqueeze(0).unsqueeze(0))


def load_params(agent, wm_sd, actor_sd, critic_sd):
    """Copy oracle_port.make_params tensors into the reference modules (same state-dict names)."""
    wm = getattr(agent.world_model, "_orig_mod", agent.world_model)
    cr = getattr(agent.critic, "_orig_mod", agent.critic)
    missing = wm.load_state_dict(wm_sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    r = agent.actor.load_state_dict(actor_sd, strict=True)
    r2 = cr.load_state_dict(critic_sd, strict=True)
    return agent


def ref_state(agent, h0, z0, logits0=None):
    from rl_sandbox.agents.dreamer.rssm import State
    N = h0.shape[0]
    lg = logits0 if logits0 is not None else torch.zeros(N, 1024)
    return State(h0.unsqueeze(0).clone(), lg.view(1, N, 32, 32).clone(), z0.unsqueeze(0).clone())
