"""DreamerV2 agent behind the reference's API (reference: rl_sandbox/agents/dreamer_v2.py:19-245).

Same constructor kwargs (the Hydra `_target_`/`_partial_` contract of config/agent/dreamer_v2*.yaml),
same methods and the same keys in the dict ``train`` returns.  What differs is where the second half
of ``train`` (dreamer_v2.py:179-211) executes, for the flat RSSM (configs 1 and 2):

  imagine_trajectory   -> K1: rlsb_rollout_fwd (ONE persistent kernel, a thread-block cluster per 128 start states;
                          up to 2048 start states) or rlsb_imagine_fwd (one tcgen05 GEMM launch per layer; the 16 k - 256 k
                          sweep), with the actor, the RSSM step, the reward / discount / target-critic heads and all draws
  lambda_return,
  cumprod weights,
  advantage            -> K2: ONE kernel, rlsb_lambda_return_fwd
  critic / actor loss,
  both backward passes -> K4: rlsb_ac_update (forward, losses, metrics, backward to parameter gradients); a continuous
                          actor (rho == 0) adds rlsb_lambda_return_bwd + rlsb_imagine_bwd for d loss / d actions
  optimizer steps      -> fused AdamW on one flat actor + critic gradient bucket (a single all-reduce when
                          torch.distributed is up)

No torch autograd runs on that path.  The slotted world model (config 3) has K1's forward (mixer blocks included) but not
its backward / K4: its behaviour half differentiates through ``_imagine_autograd`` / ``_behaviour_losses`` — the reference's
op sequence on CUDA tensors, replayed from a CUDA graph — as does a continuous actor with rssm_dim > 512 (no shipped
config).  Everything raises without a CUDA device and librlsb.so: there is no CPU path.
"""
import os
import typing as t
from pathlib import Path

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from rl_sandbox_b200.agents.rl_agent import RlAgent
from rl_sandbox_b200.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
from rl_sandbox_b200.agents.dreamer.rssm import State
from rl_sandbox_b200.utils.replay_buffer import Action, Observation, Rollout, RolloutChunks, unpack


def _strip_compile_prefix(sd: dict) -> dict:
    """Reference checkpoints carry `_orig_mod.` on world-model / critic keys (torch.compile wrapper)."""
    return {k.removeprefix('_orig_mod.'): v for k, v in sd.items()}


def _lib_handle():
    from rl_sandbox_b200 import _lib
    return _lib.load()


class DreamerV2(RlAgent):

    def __init__(self, obs_space_num: list[int], clip_rewards: str, actions_num: int, world_model: t.Any,
                 actor: t.Any, critic: t.Any, action_type: str, imagination_horizon: int, wm_optim: t.Any,
                 actor_optim: t.Any, critic_optim: t.Any, layer_norm: bool, batch_cluster_size: int,
                 f16_precision: bool, device_type: str = 'cpu', logger=None):
        self.logger = logger
        self.device = device_type
        self.imagination_horizon = imagination_horizon
        self.actions_num = actions_num
        self.is_discrete = (action_type != 'continuous')
        if clip_rewards == 'identity':
            self.reward_clipper = nn.Identity()
        elif clip_rewards == 'tanh':
            self.reward_clipper = nn.Tanh()
        else:
            raise RuntimeError('Invalid reward clipping')
        self.is_f16 = f16_precision

        self.world_model = world_model(actions_num=actions_num).to(device_type)
        self.actor: ImaginativeActor = actor(latent_dim=self.world_model.state_size, actions_num=actions_num,
                                             is_discrete=self.is_discrete).to(device_type)
        self.critic: ImaginativeCritic = critic(latent_dim=self.world_model.state_size).to(device_type)

        self.world_model_optimizer = wm_optim(model=self.world_model, scaler=self.is_f16)
        self.actor_optimizer = actor_optim(model=self.actor)
        self.critic_optimizer = critic_optim(model=self.critic)

        self._engine = None          # lazily built ImaginationEngine (needs a B200)
        self._ac_engine = None       # lazily built ACUpdateEngine (K4: fused critic/actor losses + backward)
        self._ac_packed_version = -1
        self.fused_ac_update = True  # False: torch autograd on the K1 outputs (the reference's op sequence)
        # launch-bound shapes (the configured N = 16 x 50 = 800): capture pack + K1 + K2 (+ bwd) + K4 once per shape
        # in a CUDA graph and replay it; the Philox key lives in device memory so every replay draws fresh noise
        self.cuda_graph = True
        self.cuda_graph_max_rows = int(os.environ.get('RLSB_GRAPH_MAX_ROWS', 32768))
        self.reuse_actor_forward = os.environ.get('RLSB_ACTOR_REUSE', '1') != '0'   # K1's actor activations feed K4
        # below: launch-latency bound — where the persistent rollout kernel runs (ImaginationEngine.would_run_persistent: up to
        # persistent_max_rows = 2048 start states AND all row blocks resident in one wave; it has no actor slots) the update
        # recomputes the actor forward
        self.reuse_actor_min_rows = 2049
        self.max_rows_per_pass = 131072   # start states per pass of the fused update (HBM sizing, _fused_step_chunked)
        # world-model half of train(): forward + backward captured in a CUDA graph per input shape (the observe loop is
        # T sequential steps of small kernels — thousands of launches whose CPU dispatch cost exceeds their GPU time)
        self.cuda_graph_wm = True
        self.cuda_graph_act = True    # get_action(): the batch-1 acting step replays from a CUDA graph (flat world model)
        self._act_graph = None
        self._wm_graphs: dict = {}
        self._wm_sched_prev = None
        self._graphs: dict = {}
        self._ac_bucket = None       # GradBucket over actor + critic: their .grad are views of one flat fp32 buffer
        self._weights_version = 0    # bumped whenever parameters change
        self._packed_version = -1
        self._noise_seed = 0x5EED    # Philox key; the step counter below is mixed in per rollout
        self._rollouts = 0
        self.metrics_samples = 128   # draws per element for the actor statistics (ac.py:137)
        # torch's Bernoulli.mode yields NaN when p == 0.5 exactly (world_model.py:137); one such discount
        # turns every parameter into NaN through the losses.  False = ties resolve to 1.
        self.reference_exact_discount_nan = False
        self.last_rollout: t.Optional[dict] = None
        self.reset()

    # ------------------------------------------------------------------------------------------
    # K1
    # ------------------------------------------------------------------------------------------
    def _flat_wm(self) -> bool:
        return hasattr(self.world_model, 'recurrent_model') and not getattr(self.world_model, 'slots_num', 0)

    def _get_engine(self):
        from rl_sandbox_b200 import ops
        if self._engine is None:
            wm = self.world_model
            slots = int(getattr(wm, 'slots_num', 0) or 1)
            rm = wm.recurrent_model
            cfg = ops.ImagineConfig(D=wm.rssm_dim, A=self.actions_num, discrete=self.is_discrete,
                                    layer_norm=bool(wm.layer_norm), predict_discount=bool(wm.predict_discount),
                                    H=self.imagination_horizon, groups=wm.latent_dim, classes=wm.latent_classes,
                                    with_critic=True, discount_nan_on_tie=self.reference_exact_discount_nan,
                                    with_backward=(slots == 1 and not self.is_discrete and wm.rssm_dim <= 512
                                                   and wm.rssm_dim % 8 == 0),
                                    slots=slots, attention_blocks=int(getattr(rm, 'attention_block_num', 0)),
                                    symmetric_qk=bool(getattr(rm, 'symmetric_qk', False)))
            self._engine = ops.ImaginationEngine(cfg, device=self.device)
        if self._packed_version != self._weights_version:
            self._engine.pack(self.world_model.state_dict(), self.actor.state_dict(), self.critic.state_dict())
            self._packed_version = self._weights_version
        return self._engine

    def _get_ac_engine(self):
        from rl_sandbox_b200 import ops
        eng = self._get_engine()
        if self._ac_engine is None or self._ac_engine.ccfg.metrics_samples != self.metrics_samples:
            self._ac_engine = ops.ACUpdateEngine(eng.cfg, rho=float(self.actor.rho), eta=float(self.actor.eta),
                                                 metrics_samples=self.metrics_samples, device=self.device)
            self._ac_packed_version = -1
        if self._ac_packed_version != self._weights_version:
            self._ac_engine.pack(self.actor.state_dict(), self.critic.state_dict())
            self._ac_packed_version = self._weights_version
        return self._ac_engine

    def _can_fuse_ac(self) -> bool:
        if not (self.fused_ac_update and not self.is_f16 and self._flat_wm() and str(self.device).startswith('cuda')):
            return False
        if self.is_discrete:
            return self.actor.rho == 1.0   # nothing differentiates through the rollout (ac.py:121-125)
        # continuous actor: dynamics back-propagation runs through rlsb_imagine_bwd (needs the tape: D <= 512)
        return self.world_model.rssm_dim <= 512 and self.world_model.rssm_dim % 8 == 0

    def mark_weights_changed(self):
        self._weights_version += 1

    def imagine_trajectory(self, init_state: State, precomp_actions: t.Optional[list[Action]] = None,
                           horizon: t.Optional[int] = None, noise: t.Optional[dict] = None,
                           keep_packed: bool = False, tape: bool = False, actor_slots=None,
                           last_step_value_only: bool = False) -> tuple[State, torch.Tensor, torch.Tensor, torch.Tensor]:
        """H-step closed-loop rollout from (1, N, .) start states (dreamer_v2.py:68-96).

        ``noise`` (extension, optional): {'latent_uniforms': (H,N,1024), 'action_noise': (H,N,A)} to
        inject explicit noise, or {'seed': int, 'row_offset': int} for the Philox stream."""
        if horizon is None:
            horizon = self.imagination_horizon
        needs_grad = torch.is_grad_enabled() and self.actor.rho != 1.0 and precomp_actions is None and not tape
        if needs_grad:
            return self._imagine_autograd(init_state, horizon)
        if not init_state.determ.is_cuda:
            raise RuntimeError("DreamerV2.imagine_trajectory runs on the B200 kernels: tensors must be on CUDA "
                               "(rl_sandbox_b200 has no CPU fallback)")
        eng = self._get_engine()
        N = init_state.determ.shape[1]
        S = self.world_model.latent_dim * self.world_model.latent_classes
        slotted = eng.cfg.slots > 1
        if slotted:   # (1, N, slots, .) states; the mixer coefficient follows the world model's scheduler
            eng.cfg.mixer_coeff = float(self.world_model.recurrent_model.attention_scheduler.val)
        h0 = init_state.determ[0].detach().float()
        z0 = init_state.stoch[0].detach().float()
        logits0 = init_state.stoch_logits[0].detach().float().reshape(-1, S)
        noise = dict(noise or {})
        if 'seed' not in noise and 'latent_uniforms' not in noise:
            noise['seed'] = (self._noise_seed << 20) + self._rollouts
        self._rollouts += 1
        pre = None
        if precomp_actions is not None:
            pre = torch.stack([torch.as_tensor(a) for a in precomp_actions[:horizon]]).to(h0.device).float()
            pre = pre.reshape(horizon, -1, self.actions_num).expand(horizon, N, self.actions_num).contiguous()
        out = eng.rollout(h0, z0, logits0, latent_uniforms=noise.get('latent_uniforms'),
                          action_noise=noise.get('action_noise'), seed=noise.get('seed', 0),
                          row_offset=noise.get('row_offset', 0), precomp_actions=pre, horizon=horizon,
                          keep_packed=keep_packed, want_stoch=not keep_packed, tape=tape, actor_slots=actor_slots,
                          last_step_value_only=last_step_value_only)
        self.last_rollout = out
        wm = self.world_model
        if slotted:
            from rl_sandbox_b200.agents.dreamer.rssm_slots_attention import State as SlotState
            K = eng.cfg.slots
            states = SlotState(out['determ'], out['logits'].view(horizon + 1, N, K, wm.latent_dim, wm.latent_classes),
                               out['stoch'], init_state.pos_enc)
            return states, out['actions'], out['rewards'].unsqueeze(-1), out['discounts'].unsqueeze(-1)
        states = State(out['determ'], out['logits'].view(horizon + 1, N, wm.latent_dim, wm.latent_classes),
                       out['stoch'])
        return states, out['actions'], out['rewards'].unsqueeze(-1), out['discounts'].unsqueeze(-1)

    def _imagine_autograd(self, init_state: State, horizon: int):
        """Differentiable replay of the rollout with torch ops (continuous actor, rho != 1 only)."""
        self.last_rollout = None
        prev = init_state
        zero_a = torch.zeros(prev.determ.shape[:2] + (self.actions_num,), device=prev.determ.device)
        states, actions = [init_state], [zero_a]
        rewards = [self.world_model.reward_predictor(init_state.combined).mode]
        ts = [torch.ones(zero_a.shape[:-1] + (1,), device=zero_a.device)]
        for _ in range(horizon):
            a = self.actor(prev.combined.detach()).rsample()
            prior, reward, discount = self.world_model.predict_next(prev, a)
            prev = prior
            states.append(prior); rewards.append(reward); ts.append(discount); actions.append(a)
        return states[0].stack(states), torch.cat(actions), torch.cat(rewards), torch.cat(ts)

    # ------------------------------------------------------------------------------------------
    def reset(self):
        self._state = self.world_model.get_initial_state()
        self._last_action = torch.zeros((1, 1, self.actions_num), device=self.device)
        self._action_probs = torch.zeros((self.actions_num), device=self.device)

    def preprocess(self, rollout: Rollout):
        obs = self.preprocess_obs(rollout.obs)
        additional = self.world_model.precalc_data(obs.to(self.device))
        return Rollout(obs=obs, actions=rollout.actions, rewards=self.reward_clipper(rollout.rewards),
                       is_finished=rollout.is_finished, is_first=rollout.is_first,
                       additional_data=rollout.additional_data | additional)

    def preprocess_obs(self, obs: torch.Tensor):
        """(..., H, W, 3) uint8 -> (..., 3, H, W) float in [-0.5, 0.5] (dreamer_v2.py:113-122)."""
        dims = list(range(obs.dim()))
        order = dims[:-3] + [dims[-1]] + dims[-3:-1]
        return ((obs.type(torch.float32) / 255.0) - 0.5).permute(order)

    def unprocess_obs(self, obs: torch.Tensor):
        return ((obs + 0.5).clamp(0, 1) * 255).cpu().to(dtype=torch.uint8)

    def get_action(self, obs: Observation) -> Action:
        if self.cuda_graph_act and self._flat_wm() and str(self.device).startswith('cuda') and not self.is_f16:
            return self._get_action_graphed(obs)
        obs = self.preprocess_obs(torch.from_numpy(obs).to(self.device))
        self._state = self.world_model.get_latent(obs, self._last_action, self._state)
        dist = self.actor.get_action(self._state)
        self._last_action = dist.sample()
        if self.is_discrete:
            self._action_probs += dist.probs.squeeze()
            return self._last_action.argmax()
        return self._last_action.squeeze().detach().cpu()

    def _get_action_graphed(self, obs: Observation) -> Action:
        """The acting step (dreamer_v2.py:139-154: encoder on one frame, one RSSM observe step, actor, action draw)
        replayed from a CUDA graph: ~60 batch-1 kernels whose launch cost dominates an eager call.  The recurrent state,
        the previous action and the action statistics live in static buffers the graph updates in place."""
        import torch.distributions as td
        frame = torch.from_numpy(np.ascontiguousarray(obs))
        st = self._act_graph
        if st is None or st['obs'].shape != frame.shape or st['obs'].dtype != frame.dtype:
            dev = self.device
            init = self.world_model.get_initial_state()
            st = {'obs': torch.zeros(frame.shape, dtype=frame.dtype, device=dev),
                  'state': State(init.determ.clone(), init.stoch_logits.clone(), init.stoch.clone()),
                  'action': torch.zeros((1, 1, self.actions_num), device=dev),
                  'probs': torch.zeros((self.actions_num), device=dev)}

            def body():
                x = self.preprocess_obs(st['obs'])
                new = self.world_model.get_latent(x, st['action'], st['state'])
                dist = self.actor.get_action(new)
                a = dist.sample()
                st['state'].determ.copy_(new.determ)
                st['state'].stoch_logits.copy_(new.stoch_logits)
                st['state'].stoch_.copy_(new.stoch)
                st['action'].copy_(a.reshape(st['action'].shape))
                if self.is_discrete:
                    st['probs'].add_(dist.probs.reshape(-1))
                    return a.argmax()
                return a.reshape(-1).clone()

            validate = td.Distribution._validate_args
            td.Distribution.set_default_validate_args(False)
            try:
                with torch.no_grad():
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        for _ in range(2):
                            body()
                    torch.cuda.current_stream().wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        st['out'] = body()
            finally:
                td.Distribution.set_default_validate_args(validate)
            st['graph'] = graph
            self._act_graph = st   # self._state is not the static State yet: the block below loads it (the warm-up
            #                        passes advanced the static buffers)
        if self._state is not st['state']:   # reset() (or a state set from outside): load it into the static buffers
            src = self._state if self._state is not None else self.world_model.get_initial_state()
            with torch.no_grad():
                st['state'].determ.copy_(src.determ)
                st['state'].stoch_logits.copy_(src.stoch_logits)
                st['state'].stoch_.copy_(src.stoch)
                st['action'].copy_(self._last_action.reshape(st['action'].shape))
                st['probs'].copy_(self._action_probs.reshape(-1))
            self._state = st['state']
        st['obs'].copy_(frame, non_blocking=False)
        st['graph'].replay()
        self._last_action = st['action']
        self._action_probs = st['probs']
        if self.is_discrete:
            return st['out'].clone()
        return st['out'].detach().cpu()

    def from_np(self, arr: np.ndarray):
        arr = torch.from_numpy(arr) if isinstance(arr, np.ndarray) else arr
        return arr.to(self.device, non_blocking=True)

    # ------------------------------------------------------------------------------------------
    # the hot path: imagination + lambda-return + actor-critic update (dreamer_v2.py:179-211)
    # ------------------------------------------------------------------------------------------
    def behaviour_update(self, initial_states: State, noise: t.Optional[dict] = None):
        """Second half of ``train``; returns (losses, metrics) as tensors.  Split out so that the
        benchmark and the tests can drive it from synthetic start states."""
        from rl_sandbox_b200 import ops
        if self._can_fuse_ac():
            return self._behaviour_update_fused(initial_states, noise)
        if not getattr(self, '_warned_torch_behaviour', False):
            import warnings
            self._warned_torch_behaviour = True
            why = ("the slotted world model has no rollout backward / fused update kernels" if not self._flat_wm() else
                   "fused_ac_update / f16 / device settings" if (not self.fused_ac_update or self.is_f16) else
                   f"a continuous actor with rssm_dim = {self.world_model.rssm_dim} > 512 has no rollout backward kernel")
            warnings.warn(f"DreamerV2.behaviour_update: critic / actor losses and their backward pass run as torch autograd on "
                          f"CUDA tensors ({why}); the librlsb kernels cover the rollout forward only", stacklevel=2)
        # torch-replay rollouts only (rho != 1): their noise comes from torch's graph-safe generator
        only_offset = noise is None or set(noise) <= {'row_offset'}   # (the torch replay draws from torch's generator)
        graphable = (self.cuda_graph_wm and only_offset and not self.is_f16 and initial_states.determ.is_cuda
                     and torch.is_grad_enabled() and self.actor.rho != 1.0)
        if graphable:
            return self._behaviour_update_graphed_torch(initial_states)
        losses_a, losses_c, metrics_a, metrics_c = self._behaviour_losses(initial_states, noise)
        metrics_a |= self.actor_optimizer.step(losses_a['loss_actor'])
        metrics_c |= self.critic_optimizer.step(losses_c['loss_critic'])
        self.critic.update_target()
        self.mark_weights_changed()
        return losses_a | losses_c, metrics_a | metrics_c

    def _behaviour_losses(self, initial_states: State, noise: t.Optional[dict] = None):
        """imagination -> lambda-return -> critic / actor losses under torch autograd (dreamer_v2.py:179-207): the path of
        configurations the fused update does not cover (slotted world model, D > 512 with a continuous actor)."""
        from rl_sandbox_b200 import ops
        with torch.autocast(device_type='cuda', enabled=self.is_f16):
            no_grad_rollout = self.actor.rho == 1.0
            with (torch.no_grad() if no_grad_rollout else torch.enable_grad()):
                states, actions, rewards, discounts = self.imagine_trajectory(initial_states, noise=noise)
            rewards = self.world_model.reward_normalizer(rewards.float())
            discounts = discounts.float()
            zs = states.combined
            k1 = self.last_rollout
            if k1 is not None:
                # K1 already evaluated the target critic on every state; K2 returns the lambda-returns,
                # the shifted cumprod weights and the advantage in one launch
                values = k1['values'].unsqueeze(-1)
                vs, w, _ = ops.lambda_return(rewards, values, discounts, self.critic.lambda_)
                w = w.detach()
            else:
                values = self.critic.target_critic(zs).mode
                vs = self.critic.lambda_return(zs, rewards[:-1], discounts, vs=values)
                w = torch.cumprod(torch.cat([torch.ones_like(discounts[:1]), discounts[:-1]], dim=0), dim=0).detach()
            losses_c, metrics_c = self.critic.calculate_loss(zs[:-1], vs, w[:-1], target_values=values[:-1])
            losses_a, metrics_a = self.actor.calculate_loss(zs[:-2], vs[1:], values[:-2].detach(), w[:-2],
                                                            actions[1:-1], metrics_samples=self.metrics_samples)
        return losses_a, losses_c, metrics_a, metrics_c

    def _behaviour_update_graphed_torch(self, initial_states: State):
        """``_behaviour_losses`` + both backward passes replayed from a CUDA graph (static start states; the gradients of
        loss_actor w.r.t. the actor and of loss_critic w.r.t. the critic land in static buffers); all-reduce, clipping,
        AdamW and the target update run eagerly after the replay."""
        import dataclasses
        import torch.distributions as td
        from torch.nn.utils.stateless import _reparametrize_module
        fields = {f.name: getattr(initial_states, f.name) for f in dataclasses.fields(initial_states)}
        tens = {k: v for k, v in fields.items() if torch.is_tensor(v)}
        key = (type(initial_states).__name__,) + tuple((k, tuple(v.shape), v.dtype) for k, v in tens.items()) + (
            self.metrics_samples, float(getattr(getattr(self.world_model.recurrent_model, 'attention_scheduler', None), 'val', 0.0)))
        st = self._wm_graphs.get(key)
        named_a = [(n, p) for n, p in self.actor.named_parameters() if p.requires_grad]
        named_c = [(n, p) for n, p in self.critic.named_parameters() if p.requires_grad]

        def body(static):
            state = type(initial_states)(**{k: static.get(k, v) for k, v in fields.items()})
            la = {n: p.detach().requires_grad_(True) for n, p in named_a}   # see _world_model_update_graphed
            lc = {n: p.detach().requires_grad_(True) for n, p in named_c}
            with _reparametrize_module(self.actor, la), _reparametrize_module(self.critic, lc):
                losses_a, losses_c, metrics_a, metrics_c = self._behaviour_losses(state, None)
            ga = torch.autograd.grad(losses_a['loss_actor'], list(la.values()), allow_unused=True, retain_graph=True)
            gc = torch.autograd.grad(losses_c['loss_critic'], list(lc.values()), allow_unused=True)
            return losses_a | losses_c, metrics_a | metrics_c, ga, gc

        if st is None:
            st = {'in': {k: v.detach().clone() for k, v in tens.items()}}
            validate = td.Distribution._validate_args
            td.Distribution.set_default_validate_args(False)
            rollouts0 = self._rollouts
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        body(st['in'])
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                lib = _lib_handle()
                before = lib.rlsb_launch_count(0)
                with torch.cuda.graph(graph):
                    st['out'] = body(st['in'])
                st['launches'] = lib.rlsb_launch_count(0) - before
                lib.rlsb_launch_count_add(-st['launches'])
            finally:
                td.Distribution.set_default_validate_args(validate)
                self._rollouts = rollouts0
            st['graph'] = graph
            st['grads'] = [(p, g) for (_, p), g in zip(named_a, st['out'][2])] + \
                          [(p, g) for (_, p), g in zip(named_c, st['out'][3])]
            self._wm_graphs[key] = st
        for k, v in tens.items():
            st['in'][k].copy_(v)
        st['graph'].replay()
        _lib_handle().rlsb_launch_count_add(st['launches'])
        for p, gbuf in st['grads']:
            p.grad = gbuf
        losses, metrics, _, _ = st['out']
        metrics = dict(metrics) | self.actor_optimizer.step_with_grads() | self.critic_optimizer.step_with_grads()
        self.critic.update_target()
        self.mark_weights_changed()
        return dict(losses), metrics

    def _behaviour_update_fused(self, initial_states: State, noise: t.Optional[dict]):
        """Discrete actor (rho == 1): K1 rollout keeping the packed state images -> K2 -> K4 (critic / actor
        forward, losses, backward to parameter gradients in librlsb) -> all-reduce, clip, AdamW."""
        from rl_sandbox_b200 import _lib, ops
        noise = dict(noise or {})
        n_rows = initial_states.determ.shape[1]
        explicit = 'latent_uniforms' in noise or 'action_noise' in noise
        # K4 writes both networks' gradients into views of ONE flat buffer (stable addresses for the captured graph);
        # data-parallel: a single all-reduce of that buffer after both backward passes, before either clip
        if self._ac_bucket is None:
            from rl_sandbox_b200.utils.optimizer import GradBucket
            self._ac_bucket = GradBucket(list(self.actor.actor.parameters()) + list(self.critic.critic.parameters()))
        elif not self._ac_bucket.attached():
            self._ac_bucket.attach()
        if self.cuda_graph and not explicit and n_rows <= self.cuda_graph_max_rows:
            scal = self._fused_step_graphed(initial_states, noise)
        elif n_rows > self.max_rows_per_pass and not explicit:
            with torch.no_grad():
                scal = self._fused_step_chunked(initial_states, noise)
        else:
            with torch.no_grad():
                scal = self._fused_step(initial_states, noise)
        reduced = self._ac_bucket.all_reduce()
        metrics_a = self.actor_optimizer.step_with_grads(reduced=reduced)
        metrics_c = self.critic_optimizer.step_with_grads(reduced=reduced)
        self.critic.update_target()
        self.mark_weights_changed()
        idx = _lib.AC_SCALAR_NAMES
        scal = scal.clone()
        losses = {k: scal[idx[k]] for k in ('loss_actor_reinforce', 'loss_actor_dynamics_backprop',
                                            'loss_actor_entropy', 'loss_actor', 'loss_critic')}
        metrics = {k: scal[i] for k, i in idx.items() if '/' in k}
        return losses, metrics | metrics_a | metrics_c

    def _fused_step_chunked(self, initial_states: State, noise: dict):
        """More start states than one pass holds in HBM (the rollout outputs and the K4 activation images of 131 072 start
        states x H = 15 take ~100 GB): passes over contiguous chunks, parameter gradients and scalars combined with the
        chunks' weights.  The Philox counters are global start-state indices, so the result does not depend on the split."""
        from rl_sandbox_b200 import _lib
        n = initial_states.determ.shape[1]
        chunks = -(-n // self.max_rows_per_pass)
        per = -(-n // chunks)
        per = -(-per // 128) * 128
        params = list(self.actor.actor.parameters()) + list(self.critic.critic.parameters())
        seed = noise.get('seed', (self._noise_seed << 20) + self._rollouts)
        off0 = int(noise.get('row_offset', 0))
        idx = _lib.AC_SCALAR_NAMES
        acc, scal = None, None
        for a in range(0, n, per):
            b = min(n, a + per)
            w = (b - a) / n
            sub = type(initial_states)(initial_states.determ[:, a:b], initial_states.stoch_logits[:, a:b],
                                       initial_states.stoch[:, a:b])
            s = self._fused_step(sub, {'seed': seed, 'row_offset': off0 + a}).clone()
            grads = [p.grad for p in params]
            if acc is None:
                acc = [g * w for g in grads]
                scal = s * w
                lo, hi = s[idx['actor/min_val']].clone(), s[idx['actor/max_val']].clone()
            else:
                torch._foreach_add_(acc, grads, alpha=w)
                scal += s * w
                lo, hi = torch.minimum(lo, s[idx['actor/min_val']]), torch.maximum(hi, s[idx['actor/max_val']])
        for p, g in zip(params, acc):
            p.grad.copy_(g)
        scal[idx['actor/min_val']], scal[idx['actor/max_val']] = lo, hi
        return scal

    def _fused_step(self, initial_states: State, noise: dict, seed_device=None, static=None):
        """pack (if stale) -> K1 -> K2 [-> K2 bwd -> K1 bwd] -> K4; returns the K4 scalar vector (device)."""
        from rl_sandbox_b200 import ops
        dyn = self.actor.rho != 1.0   # dynamics back-propagation (ac.py:121-123): K2 bwd -> K1 bwd -> g_actions
        # the rollout evaluates the actor on every state with the weights the update differentiates: it leaves the
        # actor's activations in the update's workspace and K4 runs the critic's forward only
        n_rows = initial_states.determ.shape[1] if static is None else static['h0'].shape[0]
        reuse = (self.reuse_actor_forward and not dyn
                 and (n_rows >= self.reuse_actor_min_rows or not self._get_engine().would_run_persistent(n_rows)))
        if static is None:
            ac = self._get_ac_engine()
            slots = ac.actor_slots(initial_states.determ.shape[1], self.imagination_horizon) if reuse else None
            # (the update reads rewards / discounts of steps 0..H-1 only: step H evaluates the target critic alone)
            self.imagine_trajectory(initial_states, noise=noise, keep_packed=True, tape=dyn, actor_slots=slots,
                                    last_step_value_only=not dyn)
            k1 = self.last_rollout
        else:   # graph body: always re-pack (the parameters change between replays), static buffers, device-resident key
            eng, ac = static['eng'], static['ac']
            pin = static['pin']   # this graph's own workspaces (ops.ImaginationEngine.workspace)
            eng.pack(self.world_model.state_dict(), self.actor.state_dict(), self.critic.state_dict())
            ac.pack(self.actor.state_dict(), self.critic.state_dict())
            slots = ac.actor_slots(static['h0'].shape[0], self.imagination_horizon, pin=pin) if reuse else None
            k1 = eng.rollout(static['h0'], static['z0'], static['logits0'], seed_device=seed_device,
                             row_offset=noise.get('row_offset', 0), keep_packed=True, tape=dyn, want_stoch=False,
                             out=static.get('out'), actor_slots=slots, pin=pin, last_step_value_only=not dyn)
            static['out'] = k1
            self.last_rollout = k1
        H, n = k1['determ'].shape[0] - 1, k1['determ'].shape[1]
        # reward_normalizer is the identity (momentum 1.0, world_model.py:111; SURVEY hard part 7)
        vs, w, _ = ops.lambda_return(k1['rewards'], k1['values'], k1['discounts'], self.critic.lambda_)
        g_actions = None
        if dyn:
            # d/d vs of -mean((1 - rho) * vs[1:] * w[:-2]) (vs has H rows: V_0 .. V_{H-1})
            g_vs = torch.zeros_like(vs)
            g_vs[1:] = w[:H - 1] * (-(1.0 - float(self.actor.rho)) / ((H - 1) * n))
            g_r, g_v, _ = ops.lambda_return_bwd(g_vs, k1['values'], k1['discounts'], vs, self.critic.lambda_)
            g_actions = self._engine.backward(k1, g_r, g_v, pin=None if static is None else static['pin'])
        return ac.update(k1, vs, w, self.actor.actor, self.critic.critic, seed=self._noise_seed + self._rollouts,
                         horizon=H, g_actions=g_actions, seed_device=seed_device, actor_forward_done=reuse,
                         pin=None if static is None else static['pin'])

    def _fused_step_graphed(self, initial_states: State, noise: dict):
        """CUDA-graph replay of ``_fused_step`` for one (rows, horizon, shard offset) shape."""
        n = initial_states.determ.shape[1]
        S = self.world_model.latent_dim * self.world_model.latent_classes
        h0 = initial_states.determ[0].detach().float()
        z0 = initial_states.stoch[0].detach().float()
        logits0 = initial_states.stoch_logits[0].detach().float().reshape(n, S)
        seed = noise.get('seed', (self._noise_seed << 20) + self._rollouts)
        self._rollouts += 1
        key = (n, self.imagination_horizon, int(noise.get('row_offset', 0)), self.metrics_samples)
        st = self._graphs.get(key)
        with torch.no_grad():
            if st is None:
                st = {'h0': h0.clone(), 'z0': z0.clone(), 'logits0': logits0.clone(),
                      'seed': torch.zeros(1, device=h0.device, dtype=torch.int64),
                      'eng': self._get_engine(), 'ac': self._get_ac_engine(), 'pin': key}
                st['seed'].fill_(seed)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):   # warm-up outside capture: workspaces, .grad buffers, kernel attributes
                    for _ in range(2):
                        self._fused_step(None, noise, seed_device=st['seed'], static=st)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                lib = st['eng'].lib
                before = lib.rlsb_launch_count(0)
                with torch.cuda.graph(graph):
                    st['scal'] = self._fused_step(None, noise, seed_device=st['seed'], static=st)
                st['launches'] = lib.rlsb_launch_count(0) - before   # kernels recorded in the graph
                lib.rlsb_launch_count_add(-st['launches'])           # capture itself launched nothing
                st['graph'] = graph
                self._graphs[key] = st
            st['h0'].copy_(h0)
            st['z0'].copy_(z0)
            st['logits0'].copy_(logits0)
            st['seed'].fill_(seed)
            st['graph'].replay()
            st['eng'].lib.rlsb_launch_count_add(st['launches'])
            self.last_rollout = st['out']
        return st['scal']

    # ------------------------------------------------------------------------------------------
    # world-model half of train() (dreamer_v2.py:172-177)
    # ------------------------------------------------------------------------------------------
    def _world_model_update(self, obs, a, r, discount_factors, first_flags, additional):
        sched = getattr(self.world_model.recurrent_model, 'attention_scheduler', None)
        sched_val = float(sched.val) if sched is not None else None
        # a Python-side schedule that is still moving would be frozen into the graph: stay eager until it settles
        stable = sched is None or sched_val == self._wm_sched_prev
        self._wm_sched_prev = sched_val
        if self.cuda_graph_wm and obs.is_cuda and not self.is_f16 and stable and torch.is_grad_enabled():
            return self._world_model_update_graphed(obs, a, r, discount_factors, first_flags, additional, sched_val)
        with torch.autocast(device_type='cuda', enabled=self.is_f16):
            losses_wm, discovered_states, metrics_wm = self.world_model.calculate_loss(
                obs, a, r, discount_factors, first_flags, additional)
        metrics_wm |= self.world_model_optimizer.step(losses_wm['loss_wm'])
        return losses_wm, discovered_states, metrics_wm

    def _world_model_update_graphed(self, obs, a, r, discount_factors, first_flags, additional, sched_val):
        """calculate_loss + loss_wm.backward() replayed from a CUDA graph (static inputs, gradients in static .grad
        buffers, device-resident noise key for the observe scan); all-reduce, clipping and AdamW run eagerly after."""
        import torch.distributions as td
        wm, opt = self.world_model, self.world_model_optimizer
        ins = {'obs': obs, 'a': a.float(), 'r': r, 'disc': discount_factors, 'first': first_flags}
        ins |= {f'add.{k}': v for k, v in additional.items() if torch.is_tensor(v)}
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in ins.items()) + (sched_val,)
        st = self._wm_graphs.get(key)

        named = [(n, p) for n, p in wm.named_parameters() if p.requires_grad]
        params = [p for _, p in named]

        def body(static):
            from torch.nn.utils.stateless import _reparametrize_module
            add = {k[4:]: v for k, v in static.items() if k.startswith('add.')}
            # Fresh leaves aliasing the parameters: their gradient edges are created on the current (capturing) stream.
            # The parameters' own AccumulateGrad nodes live on whatever stream first used them — the default stream after
            # an eager step or get_action() — and a capture cannot synchronise with that stream.
            leaves = {n: p.detach().requires_grad_(True) for n, p in named}
            for m in wm.modules():   # engines that re-pack only when a parameter version changed: always inside the graph
                if hasattr(m, 'mark_weights_changed'):
                    m.mark_weights_changed()
            with _reparametrize_module(wm, leaves):
                losses, states, metrics = wm.calculate_loss(static['obs'], static['a'], static['r'], static['disc'],
                                                            static['first'], add)
            grads = torch.autograd.grad(losses['loss_wm'], list(leaves.values()), allow_unused=True)
            return losses, states, metrics, grads

        if st is None:
            st = {'in': {k: v.detach().clone() for k, v in ins.items()},
                  'seed': torch.zeros(1, device=obs.device, dtype=torch.int64)}
            validate = td.Distribution._validate_args
            td.Distribution.set_default_validate_args(False)   # argument checks read tensors back on the host
            wm._observe_seed_device = st['seed']
            sched = getattr(wm.recurrent_model, 'attention_scheduler', None)
            t0 = sched._curr_t if sched is not None else None
            calls0 = getattr(wm, '_observe_calls', 0)
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):   # warm-up outside capture: cuDNN plans, workspaces, kernel attributes
                    for _ in range(2):
                        body(st['in'])
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                lib = _lib_handle()
                before = lib.rlsb_launch_count(0)
                with torch.cuda.graph(graph):
                    st['out'] = body(st['in'])
                st['launches'] = lib.rlsb_launch_count(0) - before
                lib.rlsb_launch_count_add(-st['launches'])
            finally:
                td.Distribution.set_default_validate_args(validate)
                wm._observe_seed_device = None
                if sched is not None:
                    sched._curr_t = t0 + 1   # the warm-up passes do not count as training steps
                if hasattr(wm, '_observe_calls'):
                    wm._observe_calls = calls0
            st['graph'] = graph
            st['grads'] = [(p, g) for p, g in zip(params, st['out'][3])]
            self._wm_graphs[key] = st
        else:
            wm.recurrent_model.on_train_step()   # the Python-side bookkeeping calculate_loss does per call
        for k, v in ins.items():
            st['in'][k].copy_(v)
        # the observe scan's Philox key: the same sequence the eager path uses (world_model.py::_observe_scan)
        st['seed'].fill_(0x0B5E0000 + getattr(wm, '_observe_calls', 0))
        if hasattr(wm, '_observe_calls'):
            wm._observe_calls += 1
        st['graph'].replay()
        _lib_handle().rlsb_launch_count_add(st['launches'])
        for p, gbuf in st['grads']:   # the graph's static gradient buffers (None: parameter not on the loss path)
            p.grad = gbuf
        losses, states, metrics, _ = st['out']
        metrics = dict(metrics) | opt.step_with_grads()
        return dict(losses), states, metrics

    def train(self, rollout_chunks: RolloutChunks):
        obs, a, r, is_finished, is_first, additional = unpack(rollout_chunks)
        if self.is_discrete:
            a = F.one_hot(a.to(torch.int64), num_classes=self.actions_num).squeeze()
        discount_factors = self.critic.gamma * (1 - is_finished).float()
        first_flags = is_first.float()

        losses_wm, discovered_states, metrics_wm = self._world_model_update(obs, a, r, discount_factors, first_flags,
                                                                          additional)
        self.mark_weights_changed()

        initial_states = discovered_states.flatten().detach()
        # data-parallel ranks hold different start states: their Philox counters are global start-state indices
        # (rank * rows per rank), so no two ranks draw the same noise and the result does not depend on the split
        noise = None
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            noise = {'row_offset': torch.distributed.get_rank() * initial_states.determ.shape[1]}
        losses_ac, metrics_ac = self.behaviour_update(initial_states, noise)

        losses = losses_wm | losses_ac
        metrics = metrics_wm | metrics_ac
        losses = {k: v.detach().cpu().numpy() for k, v in losses.items()}
        metrics = {k: v.detach().cpu().numpy() for k, v in metrics.items()}
        losses['total'] = sum(losses.values())
        return losses | metrics

    # ------------------------------------------------------------------------------------------
    def save_ckpt(self, epoch_num: int, losses: dict[str, float]):
        """Same file layout as the reference (dreamer_v2.py:222-233), including the `_orig_mod.`
        prefix its torch.compile wrappers put on world-model / critic keys."""
        pref = lambda sd: {f'_orig_mod.{k}': v for k, v in sd.items()}
        torch.save({
            'epoch': epoch_num,
            'world_model_state_dict': pref(self.world_model.state_dict()),
            'world_model_optimizer_state_dict': self.world_model_optimizer.optimizer.state_dict(),
            'actor_state_dict': self.actor.state_dict(),
            'critic_state_dict': pref(self.critic.state_dict()),
            'actor_optimizer_state_dict': self.actor_optimizer.optimizer.state_dict(),
            'critic_optimizer_state_dict': self.critic_optimizer.optimizer.state_dict(),
            'losses': losses,
        }, f'dreamerV2-{epoch_num}-{losses["total"]}.ckpt')

    def load_ckpt(self, ckpt_path: Path):
        ckpt = torch.load(ckpt_path, map_location=self.device, weights_only=False)
        self.world_model.load_state_dict(_strip_compile_prefix(ckpt['world_model_state_dict']))
        self.actor.load_state_dict(ckpt['actor_state_dict'])
        self.critic.load_state_dict(_strip_compile_prefix(ckpt['critic_state_dict']))
        # the reference calls load_state_dict on its Optimizer wrapper, which has none (its own FIXME,
        # dreamer_v2.py:238); restoring the wrapped torch optimizers is what was meant
        self.world_model_optimizer.optimizer.load_state_dict(ckpt['world_model_optimizer_state_dict'])
        self.actor_optimizer.optimizer.load_state_dict(ckpt['actor_optimizer_state_dict'])
        self.critic_optimizer.optimizer.load_state_dict(ckpt['critic_optimizer_state_dict'])
        self.mark_weights_changed()
        return ckpt['epoch']
