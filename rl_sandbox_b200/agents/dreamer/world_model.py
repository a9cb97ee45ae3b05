"""Vanilla DreamerV2 world model (reference: rl_sandbox/agents/dreamer/world_model.py:17-245).

Hot-path surface: ``predict_next`` (one imagination step + reward / discount read-out) and the
parameters K1 consumes (recurrent_model.*, reward_predictor.*, discount_predictor.*).  The observe
loop / losses (``calculate_loss``) are the "next" tier and run as torch ops.
"""
import typing as t

import torch
import torch.distributions as td
from torch import nn

from rl_sandbox_b200.agents.dreamer.common import Dist, Normalizer
from rl_sandbox_b200.agents.dreamer.rssm import RSSM, State
from rl_sandbox_b200.agents.dreamer.vision import Decoder, Encoder, SpatialBroadcastDecoder
from rl_sandbox_b200.utils.dists import DistLayer
from rl_sandbox_b200.utils.fc_nn import fc_nn_generator
from rl_sandbox_b200.vision.dino import ViTFeat


class WorldModel(nn.Module):
    def __init__(self, batch_cluster_size, latent_dim, latent_classes, rssm_dim, actions_num,
                 discount_loss_scale, kl_loss_scale, kl_loss_balancing, kl_free_nats, discrete_rssm,
                 predict_discount, layer_norm: bool, encode_vit: bool, decode_vit: bool,
                 vit_l2_ratio: float, vit_img_size: int):
        super().__init__()
        self.register_buffer('kl_free_nats', kl_free_nats * torch.ones(1))
        self.discount_scale = discount_loss_scale
        self.kl_beta = kl_loss_scale
        self.alpha = kl_loss_balancing
        self.rssm_dim, self.latent_dim, self.latent_classes = rssm_dim, latent_dim, latent_classes
        self.state_size = rssm_dim + latent_dim * latent_classes
        self.cluster_size = batch_cluster_size
        self.actions_num = actions_num
        self.predict_discount = predict_discount
        self.encode_vit, self.decode_vit = encode_vit, decode_vit
        self.vit_l2_ratio, self.vit_img_size = vit_l2_ratio, vit_img_size
        self.layer_norm = layer_norm
        if encode_vit:
            # non-functional in the reference as well (hard part 6)
            raise NotImplementedError("encode_vit=true is not supported (it is broken in the reference too)")
        norm2d = nn.GroupNorm if layer_norm else nn.Identity
        self.recurrent_model = RSSM(latent_dim, rssm_dim, actions_num, latent_classes, discrete_rssm,
                                    norm_layer=nn.LayerNorm if layer_norm else nn.Identity)
        if decode_vit:
            # frozen DINO ViT-S supplying the `d_features` loss targets (world_model.py:44-59)
            if vit_img_size == 224:
                self.dino_vit = ViTFeat("/dino/dino_deitsmall16_pretrain/dino_deitsmall16_pretrain.pth",
                                        feat_dim=384, vit_arch='small', patch_size=16)
                self.decoder_kernels, self.vit_size = [3, 3, 2], 14
            elif vit_img_size == 64:
                self.dino_vit = ViTFeat("/dino/dino_deitsmall8_pretrain/dino_deitsmall8_pretrain.pth",
                                        feat_dim=384, vit_arch='small', patch_size=8)
                self.decoder_kernels, self.vit_size = [3, 4], 8
            else:
                raise RuntimeError("Unknown vit img size")
            self.vit_feat_dim = self.dino_vit.feat_dim
            self.dino_vit.requires_grad_(False)
        # submodules are registered in the reference's order (recurrent_model, dino_vit, encoder, dino_predictor,
        # image_predictor, heads): optimizer state dicts are keyed by parameter index (checkpoint round trip)
        self.encoder = Encoder(norm_layer=norm2d, kernel_sizes=[4, 4, 4, 4], channel_step=48)
        if decode_vit:
            self.dino_predictor = SpatialBroadcastDecoder(self.state_size, norm_layer=norm2d, out_image=(14, 14),
                                                          kernel_sizes=[5, 5, 5, 5], channel_step=self.vit_feat_dim,
                                                          output_channels=self.vit_feat_dim, return_dist=True)
        self.image_predictor = Decoder(self.state_size, norm_layer=norm2d)
        head = lambda kind: fc_nn_generator(self.state_size, 1, hidden_size=400, num_layers=5,
                                            intermediate_activation=nn.ELU, layer_norm=layer_norm,
                                            final_activation=DistLayer(kind))
        self.reward_predictor = head('mse')
        self.discount_predictor = head('binary')
        self.reward_normalizer = Normalizer(momentum=1.00, scale=1.0, eps=1e-8)

    # ------------------------------------------------------------------------------------------
    def precalc_data(self, obs: torch.Tensor) -> dict[str, torch.Tensor]:
        if not self.decode_vit:
            return {}
        import torchvision as tv
        prep = tv.transforms.Compose([tv.transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
                                      tv.transforms.Resize(self.vit_img_size, antialias=True)])
        with torch.no_grad():
            return {'d_features': self.dino_vit(prep(obs + 0.5)).cpu()}

    def get_initial_state(self, batch_size: int = 1, seq_size: int = 1):
        dev = next(self.parameters()).device
        z = lambda *s: torch.zeros(seq_size, batch_size, *s, device=dev)
        return State(z(self.rssm_dim), z(self.latent_classes, self.latent_dim), z(self.latent_classes * self.latent_dim))

    def predict_next(self, prev_state: State, action):
        prior, _ = self.recurrent_model.predict_next(prev_state, action)
        reward = self.reward_predictor(prior.combined).mode
        if self.predict_discount:
            discount = self.discount_predictor(prior.combined).mode
        else:
            discount = torch.ones_like(reward)
        return prior, reward, discount

    def get_latent(self, obs: torch.Tensor, action, state: t.Optional[State]) -> State:
        if state is None:
            state = self.get_initial_state()
        embed = self.encoder(obs.unsqueeze(0))
        _, posterior, _ = self.recurrent_model.forward(state, embed.unsqueeze(0), action)
        return posterior

    # ------------------------------------------------------------------------------------------
    def _kl(self, prior_logits, post_logits):
        """KL balancing with free nats applied to the batch mean (world_model.py:169-179)."""
        kl = td.kl_divergence
        floor = self.kl_free_nats
        lhs = torch.maximum(kl(Dist(post_logits.detach()), Dist(prior_logits)).mean(), floor)
        rhs = torch.maximum(kl(Dist(post_logits), Dist(prior_logits.detach())).mean(), floor)
        return self.alpha * lhs + (1 - self.alpha) * rhs

    kernel_observe = True   # False: the reference's op sequence (T torch RSSM.forward calls under autograd)
    _observe_engine = None    # the engine of the last call
    _observe_engines = None   # {(T, E, B): ObserveEngine}
    _observe_calls = 0
    _observe_seed_device = None   # int64 device tensor holding the Philox key (CUDA-graph replays change it in place)

    def _observe_scan(self, embed, actions):
        """The T-step observe loop in librlsb (K5, rlsb_observe_fwd / _bwd under torch autograd).
        embed (B, T, E), actions (B, T, A) already masked by is_first -> posterior, prior States (T, B, .)."""
        from rl_sandbox_b200 import ops
        B, T, E = embed.shape
        rm = self.recurrent_model
        # one engine per input shape, kept alive: a captured CUDA graph of the world-model update replays the raw
        # pointers of its engine's packed weights / backward workspace, which therefore must never be freed or resized
        if self._observe_engines is None:
            self._observe_engines = {}
        eng = self._observe_engines.get((T, E, B))
        if eng is None:
            eng = ops.ObserveEngine(self.rssm_dim, self.actions_num, E, bool(self.layer_norm), T, groups=self.latent_dim,
                                    classes=self.latent_classes, device=embed.device)
            self._observe_engines[(T, E, B)] = eng
        self._observe_engine = eng
        sd = dict(rm.named_parameters())
        eng.pack({k: v.detach() for k, v in sd.items()})   # the parameters change every optimizer step
        names = eng.names()
        noise = {"seed": 0x0B5E0000 + self._observe_calls}
        if self._observe_seed_device is not None:
            noise = {"seed_device": self._observe_seed_device}
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            noise["row_offset"] = torch.distributed.get_rank() * B   # global sequence index: ranks draw distinct noise
        self._observe_calls += 1
        prior_l, post_l, determ, stoch, _idx = ops.ObserveScanFn.apply(
            eng, names, noise, embed.transpose(0, 1).float(), actions.transpose(0, 1).float(), *[sd[n] for n in names])
        shape = (T, B, self.latent_dim, self.latent_classes)
        return State(determ, post_l.view(shape), stoch), State(determ, prior_l.view(shape))

    def calculate_loss(self, obs, a, r, discount, first, additional):
        self.recurrent_model.on_train_step()
        b = obs.shape[0]
        T = self.cluster_size
        B = b // T
        embed = self.encoder(obs).reshape(B, T, -1)
        a_c = a.reshape(B, T, self.actions_num)
        r_c, d_c, first_c = r.reshape(B, T, 1), discount.reshape(B, T, 1), first.reshape(B, T, 1)

        if self.kernel_observe and embed.is_cuda and not self.recurrent_model.discrete_rssm:
            posterior, prior = self._observe_scan(embed, a_c * (1 - first_c))
        else:
            priors, posts = [], []
            state = self.get_initial_state(B)
            for step in range(T):
                a_t = (a_c[:, step] * (1 - first_c[:, step])).unsqueeze(0)
                prior, post, _ = self.recurrent_model.forward(state, embed[:, step].unsqueeze(0), a_t)
                priors.append(prior)
                posts.append(post)
                state = post
            posterior, prior = State.stack(posts), State.stack(priors)
        feat = posterior.combined.transpose(0, 1)  # (B, T, Z)

        losses, metrics = {}, {}
        r_pred, f_pred = self.reward_predictor(feat), self.discount_predictor(feat)
        losses['loss_reconstruction_img'] = torch.zeros(1, device=obs.device)
        flat = feat.flatten(0, 1)
        if not self.decode_vit:
            losses['loss_reconstruction'] = -self.image_predictor(flat).log_prob(obs).float().mean()
        else:
            if self.vit_l2_ratio != 1.0:
                img_rec = -self.image_predictor(flat).log_prob(obs).float().mean()
            else:
                img_rec = torch.zeros((), device=obs.device, dtype=torch.long)
                losses['loss_reconstruction_img'] = -self.image_predictor(flat.detach()).log_prob(obs).float().mean()
            d_obs = additional['d_features'].reshape(b, self.vit_feat_dim, self.vit_size, self.vit_size)
            d_rec = -self.dino_predictor(flat).log_prob(d_obs).float().mean()
            d_rec = d_rec / d_obs[0].numel() * obs[0].numel()
            losses['loss_reconstruction'] = self.vit_l2_ratio * d_rec + (1 - self.vit_l2_ratio) * img_rec
            metrics['loss_l2_rec'], metrics['loss_dino_rec'] = img_rec, d_rec
        losses['loss_reward_pred'] = -r_pred.log_prob(r_c).float().mean()
        losses['loss_discount_pred'] = -f_pred.log_prob(d_c).float().mean()
        losses['loss_kl_reg'] = self._kl(prior.stoch_logits, posterior.stoch_logits)
        metrics['reward_mean'], metrics['reward_std'] = r.mean(), r.std()
        metrics['reward_sae'] = torch.abs(r_pred.mode - r_c).mean()
        metrics['prior_entropy'] = Dist(prior.stoch_logits).entropy().mean()
        metrics['posterior_entropy'] = Dist(posterior.stoch_logits).entropy().mean()
        losses['loss_wm'] = (losses['loss_reconstruction'] + losses['loss_reward_pred'] +
                             self.kl_beta * losses['loss_kl_reg'] + self.discount_scale * losses['loss_discount_pred'])
        return losses, posterior, metrics
