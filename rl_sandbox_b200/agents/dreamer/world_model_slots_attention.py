"""Slotted world model (reference: rl_sandbox/agents/dreamer/world_model_slots_attention.py:18-393).

Same constructor kwargs (incl. the two the shipped YAML forgets, ``discount_loss_scale`` and
``vit_img_size``: defaults are supplied here, SURVEY hard part 6), same methods and return conventions
(``get_initial_state`` / ``get_latent`` return ``(State, slots)`` tuples).  Hot-path pieces:
``predict_next`` chains run in librlsb (K1, slots > 1) when driven by DreamerV2.imagine_trajectory;
``SlotAttention.forward`` is K3.  The observe loop, the conv encoder / decoders and the losses are torch
ops (out of scope as kernel targets).  DINO targets ``additional['d_features']`` come from
the frozen ``dino_vit`` (``precalc_data``, world_model_slots_attention.py:160-173).
"""
import typing as t

import torch
import torch.distributions as td
from torch import nn
from torch.nn import functional as F

from rl_sandbox_b200.agents.dreamer.common import Dist, Normalizer, get_position_encoding
from rl_sandbox_b200.agents.dreamer.rssm_slots_attention import RSSM, State
from rl_sandbox_b200.agents.dreamer.vision import Decoder, Encoder, SpatialBroadcastDecoder
from rl_sandbox_b200.utils.dists import DistLayer
from rl_sandbox_b200.utils.fc_nn import fc_nn_generator
from rl_sandbox_b200.vision.dino import ViTFeat
from rl_sandbox_b200.vision.slot_attention import PositionalEmbedding, SlotAttention


class WorldModel(nn.Module):
    def __init__(self, batch_cluster_size, latent_dim, latent_classes, rssm_dim, actions_num,
                 discount_loss_scale=1.0, kl_loss_scale=1.0, kl_loss_balancing=0.8, kl_free_nats=0.0,
                 discrete_rssm=False, predict_discount=False, layer_norm: bool = True, encode_vit: bool = False,
                 decode_vit: bool = False, vit_l2_ratio: float = 0.5, vit_img_size: int = 224, slots_num: int = 4,
                 slots_iter_num: int = 2, use_prev_slots: bool = True, full_qk_from: int = 1,
                 symmetric_qk: bool = False, attention_block_num: int = 3, mask_combination: str = 'soft',
                 per_slot_rec_loss: bool = False, spatial_decoder: bool = False):
        super().__init__()
        if encode_vit:
            # the reference's encode_vit branch feeds raw ViT features of the wrong rank into slot attention
            # (hard part 6): not a working configuration there either
            raise NotImplementedError("encode_vit=true is not supported (it is broken in the reference too)")
        self.use_prev_slots = use_prev_slots
        self.register_buffer('kl_free_nats', kl_free_nats * torch.ones(1))
        self.discount_scale, self.kl_beta, self.alpha = discount_loss_scale, kl_loss_scale, kl_loss_balancing
        self.rssm_dim, self.latent_dim, self.latent_classes = rssm_dim, latent_dim, latent_classes
        self.slots_num = slots_num
        self.mask_combination = mask_combination
        self.state_size = slots_num * (rssm_dim + latent_dim * latent_classes)
        self.cluster_size, self.actions_num = batch_cluster_size, actions_num
        self.predict_discount, self.layer_norm = predict_discount, layer_norm
        self.encode_vit, self.decode_vit = encode_vit, decode_vit
        self.vit_l2_ratio, self.vit_img_size = vit_l2_ratio, vit_img_size
        self.per_slot_rec_loss = per_slot_rec_loss
        self.n_dim = 384
        norm2d = nn.GroupNorm if layer_norm else (lambda *a, **k: nn.Identity())
        self.recurrent_model = RSSM(latent_dim, rssm_dim, actions_num, latent_classes, discrete_rssm,
                                    norm_layer=nn.LayerNorm if layer_norm else nn.Identity, embed_size=self.n_dim,
                                    full_qk_from=full_qk_from, symmetric_qk=symmetric_qk,
                                    attention_block_num=attention_block_num)
        if decode_vit:
            # frozen DINO ViT-S (world_model_slots_attention.py:66-82); registered right after recurrent_model as in
            # the reference so that parameter indices (optimizer state) and checkpoint keys `dino_vit.*` line up
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
        self.encoder = Encoder(norm_layer=norm2d, kernel_sizes=[4, 4], channel_step=48 * (self.n_dim // 192) * 2,
                               post_conv_num=2, flatten_output=False)
        self.slot_attention = SlotAttention(slots_num, self.n_dim, slots_iter_num, use_prev_slots)
        self.register_buffer('pos_enc', torch.from_numpy(
            get_position_encoding(slots_num, self.state_size // slots_num)).to(dtype=torch.float32))
        self.positional_augmenter_inp = PositionalEmbedding(self.n_dim, (14, 14))
        self.slot_mlp = nn.Sequential(nn.Linear(self.n_dim, self.n_dim), nn.ReLU(inplace=True),
                                      nn.Linear(self.n_dim, self.n_dim))
        z = rssm_dim + latent_dim * latent_classes
        if decode_vit and spatial_decoder:
            self.dino_predictor = SpatialBroadcastDecoder(z, norm_layer=norm2d, out_image=(14, 14), kernel_sizes=[5, 5, 5],
                                                          channel_step=self.vit_feat_dim,
                                                          output_channels=self.vit_feat_dim + 1, return_dist=False)
        elif decode_vit:
            self.dino_predictor = Decoder(z, norm_layer=norm2d, conv_kernel_sizes=[3], channel_step=self.vit_feat_dim,
                                          kernel_sizes=self.decoder_kernels, output_channels=self.vit_feat_dim + 1,
                                          return_dist=False)
        self.image_predictor = Decoder(z, norm_layer=norm2d, output_channels=3 + 1, return_dist=False)
        head = lambda kind: fc_nn_generator(self.state_size, 1, hidden_size=400, num_layers=5,
                                            intermediate_activation=nn.ELU, layer_norm=layer_norm,
                                            final_activation=DistLayer(kind))
        self.reward_predictor = head('mse')
        self.discount_predictor = head('binary')
        self.reward_normalizer = Normalizer(momentum=1.00, scale=1.0, eps=1e-8)
        self.last_attn = None

    # --------------------------------------------------------------------------------------------
    def slot_mask(self, masks: torch.Tensor) -> torch.Tensor:
        if self.mask_combination == 'soft':
            return F.softmax(masks, dim=1)
        if self.mask_combination == 'hard':
            probs = F.softmax(masks - masks.logsumexp(dim=1, keepdim=True), dim=1)
            return F.one_hot(masks.argmax(dim=1), num_classes=masks.shape[1]).permute(0, 4, 1, 2, 3) + (probs - probs.detach())
        raise NotImplementedError

    def precalc_data(self, obs: torch.Tensor) -> dict[str, torch.Tensor]:
        if not self.decode_vit:
            return {}
        import torchvision as tv
        prep = tv.transforms.Compose([tv.transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
                                      tv.transforms.Resize(self.vit_img_size, antialias=True)])
        return {'d_features': self.dino_vit(prep(obs + 0.5)).squeeze()}

    def get_initial_state(self, batch_size: int = 1, seq_size: int = 1):
        dev = next(self.parameters()).device
        z = lambda *s: torch.zeros(seq_size, batch_size, self.slots_num, *s, device=dev)
        return State(z(self.rssm_dim), z(self.latent_classes, self.latent_dim),
                     z(self.latent_classes * self.latent_dim), self.pos_enc.unsqueeze(0).unsqueeze(0)), None

    def predict_next(self, prev_state: State, action):
        prior, _ = self.recurrent_model.predict_next(prev_state, action)
        reward = self.reward_predictor(prior.combined).mode
        if self.predict_discount:
            discount_factors = self.discount_predictor(prior.combined).mode
        else:
            discount_factors = torch.ones_like(reward)
        return prior, reward, discount_factors

    def _slot_features(self, obs):
        embed = self.positional_augmenter_inp(self.encoder(obs))
        return self.slot_mlp(embed.permute(0, 2, 3, 1).reshape(obs.shape[0], -1, self.n_dim))

    def get_latent(self, obs: torch.Tensor, action, state):
        if state is None or state[0] is None:
            state, prev_slots = self.get_initial_state()
        elif self.use_prev_slots:
            state, prev_slots = state
        else:
            state, prev_slots = state[0], None
        feats = self._slot_features(obs.unsqueeze(0))
        slots_t = self.slot_attention(feats, prev_slots)
        _, posterior, _ = self.recurrent_model.forward(state, slots_t.unsqueeze(0), action)
        return posterior, slots_t

    def _decode(self, predictor, posterior, b, channels, h, w, flatten_to):
        out = predictor(posterior.combined_slots.transpose(0, 1).flatten(0, flatten_to))
        return out.reshape(b, -1, channels + 1, h, w).split([channels, 1], dim=2)

    def _rec_loss(self, decoded, mask, target):
        decoded = decoded * mask
        if self.per_slot_rec_loss:
            l2 = (mask * ((decoded - target.unsqueeze(1)) ** 2)).sum(dim=[2, 3, 4])
            norm = torch.prod(torch.tensor(target.shape)[-3:]) / mask.sum(dim=[2, 3, 4]).clamp(min=1)
            return l2, norm
        mean = torch.sum(decoded, dim=1)
        dist = td.Independent(td.Normal(mean, torch.ones((), device=mean.device, dtype=mean.dtype)), 3)
        return -dist.log_prob(target).float().mean(), None

    def calculate_loss(self, obs, a, r, discount, first, additional):
        self.recurrent_model.on_train_step()
        b, _, h, w = obs.shape
        feats = self._slot_features(obs)
        a_c = a.reshape(-1, self.cluster_size, self.actions_num)
        r_c = r.reshape(-1, self.cluster_size, 1)
        d_c = discount.reshape(-1, self.cluster_size, 1)
        first_c = first.reshape(-1, self.cluster_size, 1)
        losses, metrics = {}, {}
        nb = b // self.cluster_size

        def KL(dist1, dist2):
            kl = torch.distributions.kl_divergence
            one = lambda lg: td.Independent(td.OneHotCategoricalStraightThrough(logits=lg), 1)
            lhs = torch.maximum(kl(one(dist2.detach()), one(dist1)).mean(), self.kl_free_nats)
            rhs = torch.maximum(kl(one(dist2), one(dist1.detach())).mean(), self.kl_free_nats)
            return self.alpha * lhs + (1 - self.alpha) * rhs

        prev_state, _ = self.get_initial_state(nb)
        self.last_attn = torch.zeros((self.slots_num, self.slots_num), device=a_c.device)
        prev_slots = self.slot_attention.generate_initial(nb).repeat(self.cluster_size, 1, 1, 1).transpose(0, 1)
        slots_c = self.slot_attention(feats, prev_slots.flatten(0, 1)).reshape(nb, self.cluster_size, self.slots_num, -1)
        priors, posteriors = [], []
        for t_ in range(self.cluster_size):
            a_t = a_c[:, t_].unsqueeze(0) * (1 - first_c[:, t_].unsqueeze(0))
            prior, posterior, _ = self.recurrent_model.forward(prev_state, slots_c[:, t_].unsqueeze(0), a_t)
            prev_state = posterior
            self.last_attn += self.recurrent_model.last_attention
            priors.append(prior)
            posteriors.append(posterior)
        self.last_attn /= self.cluster_size
        posterior, prior = State.stack(posteriors), State.stack(priors)
        r_pred = self.reward_predictor(posterior.combined.transpose(0, 1))
        f_pred = self.discount_predictor(posterior.combined.transpose(0, 1))
        losses['loss_reconstruction_img'] = torch.zeros((), device=obs.device, dtype=torch.long)

        def image_rec(post):
            imgs, masks = self._decode(self.image_predictor, post, b, 3, h, w, 2)
            val, norm = self._rec_loss(imgs, self.slot_mask(masks), obs)
            return val.mean() * norm * 8 if norm is not None else val

        if not self.decode_vit:
            losses['loss_reconstruction'] = image_rec(posterior)
        else:
            if self.vit_l2_ratio != 1.0:
                img_rec = image_rec(posterior)
            else:
                img_rec = torch.zeros((), device=obs.device, dtype=torch.long)
                losses['loss_reconstruction_img'] = image_rec(posterior.detach())
            d_features = additional['d_features']
            feats_dec, masks = self._decode(self.dino_predictor, posterior, b, self.vit_feat_dim, self.vit_size,
                                            self.vit_size, 1)
            d_obs = d_features.reshape(b, self.vit_feat_dim, self.vit_size, self.vit_size)
            val, norm = self._rec_loss(feats_dec, self.slot_mask(masks), d_obs)
            d_rec = val.mean() * norm * 4 if norm is not None else val
            d_rec = d_rec / torch.prod(torch.tensor(d_obs.shape[-3:])) * torch.prod(torch.tensor(obs.shape[-3:]))
            losses['loss_reconstruction'] = self.vit_l2_ratio * d_rec + (1 - self.vit_l2_ratio) * img_rec
            metrics['loss_l2_rec'] = img_rec
            metrics['loss_dino_rec'] = d_rec
        losses['loss_reward_pred'] = -r_pred.log_prob(r_c).float().mean()
        losses['loss_discount_pred'] = -f_pred.log_prob(d_c).float().mean()
        losses['loss_kl_reg'] = KL(prior.stoch_logits, posterior.stoch_logits)
        metrics['attention_coeff'] = torch.tensor(self.recurrent_model.attention_scheduler.val)
        metrics['reward_mean'] = r.mean()
        metrics['reward_std'] = r.std()
        metrics['reward_sae'] = (torch.abs(r_pred.mode - r_c)).mean()
        metrics['prior_entropy'] = Dist(prior.stoch_logits).entropy().mean()
        metrics['posterior_entropy'] = Dist(posterior.stoch_logits).entropy().mean()
        losses['loss_wm'] = (losses['loss_reconstruction'] + losses['loss_reward_pred'] +
                             self.kl_beta * losses['loss_kl_reg'] + self.discount_scale * losses['loss_discount_pred'])
        return losses, posterior, metrics
