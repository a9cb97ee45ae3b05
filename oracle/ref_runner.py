"""ref_runner.py — BENCH INFRASTRUCTURE: the hot path executed by the UNMODIFIED reference modules.

``ReferenceHotPath.step`` is the second half of ``DreamerV2.train`` (reference agents/dreamer_v2.py:179-211) — imagination,
lambda-return, shifted discount cumprod, critic / actor losses, both ``Optimizer.step`` calls, ``update_target`` and the
host read of the losses (:216-217) — written as calls of the reference's own methods on its own agent object, from
synthetic start states (the first half of ``train``, the world-model update, is not on the path BASELINE.json names).
Nothing is restated: every tensor operation runs inside the reference's code (imported from /root/reference here, from the
staged copy oracle/_ref/ on the GPU box; import shims as in oracle/ref_harness.py).  Used by ``bench.py --impl reference``
(host cores) and by the ``torch_gpu_baseline`` leg (the same object on cuda: the reference's eager PyTorch path, TF32
allowed as its train.py:40).
"""
from __future__ import annotations

import torch

from . import ref_harness as rh


def available() -> bool:
    return rh.available()


class ReferenceHotPath:
    def __init__(self, *, D, A, discrete, layer_norm, predict_discount, H=15, eta=3e-3, lr=1e-4, gamma=0.999,
                 device="cpu", seed=0):
        torch.manual_seed(seed)
        self.agent = rh.build_agent(D=D, A=A, discrete=discrete, layer_norm=layer_norm, predict_discount=predict_discount,
                                    H=H, entropy_scale=eta, gamma=gamma, lr=lr, clip_rewards="tanh", device_type=device)
        self.device, self.D = device, D
        from rl_sandbox.agents.dreamer.rssm import State   # the reference's (ref_harness put it first on sys.path)
        self._State = State

    def state(self, h0: torch.Tensor, z0: torch.Tensor):
        n = h0.shape[0]
        return self._State(h0.unsqueeze(0).to(self.device), torch.zeros(1, n, 32, 32, device=self.device),
                           z0.unsqueeze(0).to(self.device))

    def step(self, initial_states) -> dict:
        ag = self
        self = ag.agent
        # ---- reference agents/dreamer_v2.py:182-217, verbatim call sequence --------------------------------------
        states, actions, rewards, discount_factors = self.imagine_trajectory(initial_states)
        rewards = rewards.float()
        discount_factors = discount_factors.float()
        zs = states.combined
        rewards = self.world_model.reward_normalizer(rewards)
        vs = self.critic.lambda_return(zs, rewards[:-1], discount_factors)
        discount_factors = torch.cat([torch.ones_like(discount_factors[:1]), discount_factors[:-1]], dim=0)
        discount_factors = torch.cumprod(discount_factors, dim=0).detach()
        losses_c, metrics_c = self.critic.calculate_loss(zs[:-1], vs, discount_factors[:-1])
        losses_a, metrics_a = self.actor.calculate_loss(zs[:-2], vs[1:], self.critic.target_critic(zs[:-2]).mode,
                                                        discount_factors[:-2], actions[1:-1])
        metrics_a |= self.actor_optimizer.step(losses_a['loss_actor'])
        metrics_c |= self.critic_optimizer.step(losses_c['loss_critic'])
        self.critic.update_target()
        losses = losses_a | losses_c
        metrics = metrics_a | metrics_c
        losses = {k: v.detach().cpu().numpy() for k, v in losses.items()}
        metrics = {k: v.detach().cpu().numpy() for k, v in metrics.items()}
        return losses | metrics
