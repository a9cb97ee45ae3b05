"""Actor and critic heads of DreamerV2 (reference: rl_sandbox/agents/dreamer/ac.py:11-146).

The modules own the fp32 parameters (AdamW, checkpoints and the gradient all-reduce work on
ordinary tensors).  On the B200 path:
  * ``lambda_return`` / the cumprod weights / the advantage run in librlsb (K2, rlsb_lambda_return_fwd);
  * during imagination the actor MLP and the target critic are evaluated inside K1
    (rlsb_imagine_fwd), which hands back ``values`` so lambda_return does not re-run the target critic;
  * the loss MLP forward+backward is torch autograd (cuBLAS) — SURVEY 8f rank 2 ("next").
"""
import typing as t

import torch
import torch.distributions as td
from torch import nn

from rl_sandbox_b200.utils.dists import DistLayer
from rl_sandbox_b200.utils.fc_nn import fc_nn_generator


def _kernels_usable(x: torch.Tensor) -> bool:
    return x.is_cuda


class ImaginativeCritic(nn.Module):
    def __init__(self, discount_factor: float, update_interval: int, soft_update_fraction: float,
                 value_target_lambda: float, latent_dim: int, layer_norm: bool):
        super().__init__()
        self.gamma = discount_factor
        self.critic_update_interval = update_interval
        self.lambda_ = value_target_lambda
        self.critic_soft_update_fraction = soft_update_fraction
        self._update_num = 0
        make = lambda: fc_nn_generator(latent_dim, 1, 400, 5, intermediate_activation=nn.ELU,
                                       layer_norm=layer_norm, final_activation=DistLayer('mse'))
        self.critic = make()
        self.target_critic = make()
        self.target_critic.requires_grad_(False)

    def update_target(self):
        """Hard copy on calls 0, interval, 2*interval, ... (ac.py:39-47)."""
        if self._update_num == 0:
            self.target_critic.load_state_dict(self.critic.state_dict())
        self._update_num = (self._update_num + 1) % self.critic_update_interval

    def estimate_value(self, z) -> td.Distribution:
        return self.critic(z)

    def _lambda_return(self, vs: torch.Tensor, rs: torch.Tensor, ds: torch.Tensor):
        """V_T = vs[T]; V_i = rs[i] + ds[i]*((1-l)*vs[i+1] + l*V_{i+1})  -> V_0..V_{T-1}  (ac.py:52-62).
        vs has one more row than rs; ds may have T or T+1 rows (rows 0..T-1 are read)."""
        T = rs.shape[0]
        if not _kernels_usable(vs):
            raise RuntimeError("ImaginativeCritic._lambda_return: the B200 kernels need CUDA tensors "
                               "(rl_sandbox_b200 has no CPU fallback)")
        from rl_sandbox_b200 import ops
        shape = vs.shape[1:]
        pad = torch.zeros((1,) + tuple(shape), device=vs.device, dtype=torch.float32)
        r = torch.cat([rs.float(), pad]) if rs.shape[0] == T else rs.float()
        d = ds.float()
        d = torch.cat([d, pad]) if d.shape[0] == T else d[:T + 1]
        if torch.is_grad_enabled() and (r.requires_grad or vs.requires_grad or d.requires_grad):
            return ops.LambdaReturnFn.apply(r.contiguous(), vs.float().contiguous(), d.contiguous(), self.lambda_)
        out, _, _ = ops.lambda_return(r, vs.float(), d, self.lambda_, want_weights=False, want_adv=False)
        return out

    def lambda_return(self, zs, rs, ds, vs: t.Optional[torch.Tensor] = None):
        """``vs`` may be supplied by K1 (target critic evaluated inside the rollout); otherwise it is
        computed here like the reference does (ac.py:64-66)."""
        if vs is None:
            vs = self.target_critic(zs).mode
        return self._lambda_return(vs, rs, ds)

    def calculate_loss(self, zs: torch.Tensor, vs: torch.Tensor, discount_factors: torch.Tensor,
                       target_values: t.Optional[torch.Tensor] = None):
        pred = self.estimate_value(zs.detach())
        losses = {'loss_critic': -(pred.log_prob(vs.detach()).unsqueeze(2) * discount_factors).mean()}
        if target_values is None:
            target_values = self.target_critic(zs).mode
        metrics = {'critic/avg_target_value': target_values.mean(),
                   'critic/avg_lambda_value': vs.mean(),
                   'critic/avg_predicted_value': pred.mode.mean()}
        return losses, metrics


class ImaginativeActor(nn.Module):
    def __init__(self, latent_dim: int, actions_num: int, is_discrete: bool, layer_norm: bool,
                 reinforce_fraction: t.Optional[float], entropy_scale: float):
        super().__init__()
        self.rho = is_discrete if reinforce_fraction is None else reinforce_fraction
        self.eta = entropy_scale
        self.is_discrete = is_discrete
        self.actions_num = actions_num
        self.actor = fc_nn_generator(latent_dim, actions_num if is_discrete else actions_num * 2, 400, 5,
                                     layer_norm=layer_norm, intermediate_activation=nn.ELU,
                                     final_activation=DistLayer('onehot' if is_discrete else 'normal_trunc'))

    def forward(self, z: torch.Tensor) -> td.Distribution:
        return self.actor(z)

    def get_action(self, state) -> td.Distribution:
        if isinstance(state, tuple):  # slotted world models return (State, slots)
            state = state[0]
        return self.actor(state.combined)

    def calculate_loss(self, zs: torch.Tensor, vs: torch.Tensor, baseline: torch.Tensor,
                       discount_factors: torch.Tensor, actions: torch.Tensor, metrics_samples: int = 128):
        dist = self.actor(zs.detach())
        advantage = (vs - baseline).detach()
        losses = {}
        losses['loss_actor_reinforce'] = -(self.rho * dist.log_prob(actions.detach()).unsqueeze(2) *
                                           discount_factors * advantage).mean()
        if self.rho != 1.0:
            losses['loss_actor_dynamics_backprop'] = -((1 - self.rho) * (vs * discount_factors)).mean()
        else:
            losses['loss_actor_dynamics_backprop'] = torch.tensor(0)
        losses['loss_actor_entropy'] = -(self.eta * dist.entropy().unsqueeze(2) * discount_factors).mean()
        losses['loss_actor'] = (losses['loss_actor_reinforce'] + losses['loss_actor_dynamics_backprop'] +
                                losses['loss_actor_entropy'])
        # statistics of the action distribution from `metrics_samples` draws per element (ac.py:137-143)
        metrics = {}
        with torch.no_grad():
            sample = dist.rsample((metrics_samples,))
            avg = sample.mean(0)
            metrics['actor/avg_val'] = avg.mean()
            metrics['actor/mean_val'] = dist.mean.mean()
            metrics['actor/avg_sd'] = ((sample - avg) ** 2).mean(0).sqrt().mean()
            metrics['actor/min_val'] = sample.min()
            metrics['actor/max_val'] = sample.max()
        return losses, metrics
