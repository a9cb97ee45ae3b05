"""Frozen DINO ViT feature extractor (reference: rl_sandbox/vision/dino.py:178-352).

``ViTFeat`` supplies the world-model loss targets ``d_features`` (`WorldModel.precalc_data`,
world_model.py:113-129): the *key* projections of the last transformer block for every patch token.
It is frozen, runs once per replay-buffer insert (``DreamerV2.preprocess``) and is outside the kernel
scope: plain torch, attention through ``scaled_dot_product_attention``.

State-dict keys are the reference's (``model.cls_token``, ``model.pos_embed``,
``model.patch_embed.proj.*``, ``model.blocks.<i>.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*``,
``model.norm.*``), so DINO checkpoints and reference agent checkpoints load unchanged.

Only the last block's q/k/v projection is needed for the features, so ``forward`` stops after
``qkv(norm1(x))`` of that block instead of evaluating its attention map as the reference does
(dino.py:327-331 reads ``feat_qkv`` and discards the attention) — same values, 1/12 less work.
"""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

_ARCH = {
    # name: (embed_dim, depth, heads)          vit_small / vit_base of dino.py:282-295
    'small': (384, 12, 6),
    'base': (768, 12, 12),
}
DINO_URL_ROOT = "https://dl.fbaipublicfiles.com"


class _SelfAttention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, T, C = x.shape
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, T, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        y = F.scaled_dot_product_attention(q, k, v)            # scale = head_dim ** -0.5 (dino.py:115)
        return self.proj(y.transpose(1, 2).reshape(B, T, C))


class _FeedForward(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float = 4.0, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = _SelfAttention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _FeedForward(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size: int, patch_size: int, in_chans: int, dim: int):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    """ViT backbone with the DINO parameter layout (dino.py:178-278); no classifier head."""

    def __init__(self, img_size=(224,), patch_size=16, in_chans=3, embed_dim=384, depth=12, num_heads=6,
                 mlp_ratio=4.0):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.patch_embed = _PatchEmbed(img_size[0], patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _pos_embed_for(self, n_patches: int, w: int, h: int):
        """Bicubic resampling of the patch position table for other input sizes (dino.py:213-234)."""
        table = self.pos_embed.shape[1] - 1
        if n_patches == table and w == h:
            return self.pos_embed
        side = int(math.sqrt(table))
        ps = self.patch_embed.patch_size
        w0, h0 = w // ps + 0.1, h // ps + 0.1          # the +0.1 of facebookresearch/dino#8
        grid = self.pos_embed[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, scale_factor=(w0 / side, h0 / side), mode='bicubic')
        if (int(w0), int(h0)) != tuple(grid.shape[-2:]):
            raise RuntimeError(f"position table {side}x{side} cannot be resampled to {int(w0)}x{int(h0)}")
        grid = grid.permute(0, 2, 3, 1).reshape(1, -1, self.pos_embed.shape[-1])
        return torch.cat([self.pos_embed[:, :1], grid], dim=1)

    def prepare_tokens(self, x):
        B, _, w, h = x.shape
        tok = self.patch_embed(x)
        tok = torch.cat([self.cls_token.expand(B, -1, -1), tok], dim=1)
        return tok + self._pos_embed_for(tok.shape[1] - 1, w, h)

    def forward(self, x):
        x = self.prepare_tokens(x)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 0]

    def last_block_qkv(self, x):
        """``qkv(norm1(.))`` of the last block on the output of blocks[:-1]: (B, tokens, 3 * dim)."""
        x = self.prepare_tokens(x)
        for blk in self.blocks[:-1]:
            x = blk(x)
        last = self.blocks[-1]
        return last.attn.qkv(last.norm1(x))


def vit_small(patch_size=16, img_size=(224,), **kw):
    d, depth, heads = _ARCH['small']
    return VisionTransformer(img_size=img_size, patch_size=patch_size, embed_dim=d, depth=depth, num_heads=heads, **kw)


def vit_base(patch_size=16, img_size=(224,), **kw):
    d, depth, heads = _ARCH['base']
    return VisionTransformer(img_size=img_size, patch_size=patch_size, embed_dim=d, depth=depth, num_heads=heads, **kw)


class ViTFeat(nn.Module):
    """``ViTFeat(pretrained_pth, feat_dim, vit_arch, vit_feat, patch_size)`` as in dino.py:298-352.

    ``pretrained_pth`` is the path below https://dl.fbaipublicfiles.com the reference downloads
    (dino.py:313); the same ``torch.hub.load_state_dict_from_url`` call is made here, so a populated hub
    cache (or a patched loader) yields identical weights.  Without network access and without a cached
    file the backbone keeps its random initialisation and a warning is issued: the features are only
    regression *targets* of the world-model loss, so the hot path's arithmetic is unaffected.
    ``pretrained_pth=None`` skips the lookup.
    """

    def __init__(self, pretrained_pth, feat_dim, vit_arch='base', vit_feat='k', patch_size=16, img_size=(224,)):
        super().__init__()
        make = vit_base if vit_arch == 'base' else vit_small
        self.model = make(patch_size=patch_size, img_size=img_size)
        self.feat_dim, self.vit_feat, self.patch_size = feat_dim, vit_feat, patch_size
        self.pretrained = False
        if pretrained_pth is not None:
            try:
                sd = torch.hub.load_state_dict_from_url(DINO_URL_ROOT + pretrained_pth, map_location='cpu')
            except Exception as exc:   # offline box, no cached checkpoint
                warnings.warn(f"DINO weights {pretrained_pth} unavailable ({type(exc).__name__}: {exc}); "
                              "ViTFeat keeps its random initialisation", stacklevel=2)
            else:
                self.model.load_state_dict(sd, strict=True)
                self.pretrained = True

    @torch.no_grad()
    def forward(self, img):
        B, _, h, w = img.shape
        fh, fw = h // self.patch_size, w // self.patch_size
        qkv = self.model.last_block_qkv(img)[:, 1:]                      # drop [CLS]: (B, fh*fw, 3*C)
        q, k, v = (t.transpose(1, 2).reshape(B, self.feat_dim, fh * fw) for t in qkv.chunk(3, dim=-1))
        if self.vit_feat == 'k':
            return k
        if self.vit_feat == 'q':
            return q
        if self.vit_feat == 'v':
            return v
        if self.vit_feat == 'kqv':
            return torch.cat([k, q, v], dim=1)
        raise ValueError(f"unknown vit_feat {self.vit_feat!r}")
