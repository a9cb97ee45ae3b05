"""Conv encoder / decoders of the world model (reference: rl_sandbox/agents/dreamer/vision.py:7-145).
Out of scope as kernel targets (cuDNN through torch); present so the world-model half of
DreamerV2.train() runs and checkpoints load.  Sequential indices match the reference."""
import torch
import torch.distributions as td
from torch import nn


class Encoder(nn.Module):
    def __init__(self, norm_layer, channel_step=96, kernel_sizes=[4, 4, 4, 4], post_conv_num: int = 0,
                 flatten_output=True, in_channels=3):
        super().__init__()
        mods, c_in = [], in_channels
        for i, k in enumerate(kernel_sizes):
            c_out = channel_step << i
            mods += [nn.Conv2d(c_in, c_out, kernel_size=k, stride=2), norm_layer(1, c_out), nn.ELU(inplace=True)]
            c_in = c_out
        for _ in range(post_conv_num):
            mods += [nn.Conv2d(c_in, c_in, kernel_size=5, padding='same'), norm_layer(1, c_in), nn.ELU(inplace=True)]
        if flatten_output:
            mods.append(nn.Flatten())
        self.net = nn.Sequential(*mods)

    def forward(self, X):
        return self.net(X)


class Decoder(nn.Module):
    def __init__(self, input_size, norm_layer, kernel_sizes=[5, 5, 6, 6], channel_step=48, output_channels=3,
                 conv_kernel_sizes=[], return_dist=True):
        super().__init__()
        n = len(kernel_sizes)
        self.channel_step = channel_step
        self.in_channels = (2 ** (n + 1)) * channel_step
        self.convin = nn.Linear(input_size, self.in_channels)
        self.return_dist = return_dist
        mods, c_in = [], self.in_channels
        for i, k in enumerate(kernel_sizes):
            last = i == n - 1
            c_out = output_channels if last else (2 ** (n - i - 2)) * channel_step
            mods.append(nn.ConvTranspose2d(c_in, c_out, kernel_size=k, stride=2, output_padding=0))
            if not last:
                mods += [norm_layer(1, c_out), nn.ELU(inplace=True)]
            c_in = c_out
        for k in conv_kernel_sizes:
            mods += [norm_layer(1, c_in), nn.ELU(inplace=True),
                     nn.Conv2d(output_channels, output_channels, kernel_size=k, padding='same')]
        self.net = nn.Sequential(*mods)

    def forward(self, X):
        y = self.net(self.convin(X).view(-1, self.in_channels, 1, 1))
        return td.Independent(td.Normal(y, torch.ones((), device=y.device, dtype=y.dtype)), 3) if self.return_dist else y


class SpatialBroadcastDecoder(nn.Module):
    """Latent vector tiled over the output grid + learned position embedding + 'same' convolutions
    (reference: vision.py:40-89).  Decodes the DINO feature targets of config_dino (14 x 14 x 384)."""

    def __init__(self, input_size, norm_layer, kernel_sizes=[3, 3, 3], out_image=(64, 64), channel_step=64,
                 output_channels=3, return_dist=True):
        super().__init__()
        from rl_sandbox_b200.vision.slot_attention import PositionalEmbedding
        self.channel_step = channel_step
        self.in_channels = 2 * channel_step
        self.out_shape = out_image
        self.positional_augmenter = PositionalEmbedding(self.in_channels, out_image)
        self.convin = nn.Linear(input_size, self.in_channels)
        self.return_dist = return_dist
        mods, c_in = [], self.in_channels
        for i, k in enumerate(kernel_sizes):
            last = i == len(kernel_sizes) - 1
            c_out = output_channels if last else channel_step
            mods.append(nn.Conv2d(c_in, c_out, kernel_size=k, padding='same'))
            if not last:
                mods += [norm_layer(1, c_out), nn.ELU(inplace=True)]
            c_in = c_out
        self.net = nn.Sequential(*mods)

    def forward(self, X):
        x = self.convin(X).view(-1, self.in_channels, 1, 1)
        x = self.positional_augmenter(x.expand(-1, -1, *self.out_shape))   # broadcast instead of torch.tile's copy
        y = self.net(x)
        return td.Independent(td.Normal(y, torch.ones((), device=y.device, dtype=y.dtype)), 3) if self.return_dist else y


class ViTDecoder(nn.Module):
    """Transposed-conv decoder to a 384-channel feature map (reference: vision.py:147-187)."""

    def __init__(self, input_size, norm_layer, kernel_sizes=[5, 5, 5, 3, 3]):
        super().__init__()
        self.channel_step = 12
        width = 32 * self.channel_step
        self.convin = nn.Linear(input_size, width)
        n = len(kernel_sizes)
        mods, c_in = [], width
        for i, k in enumerate(kernel_sizes):
            if i == n - 1:
                mods.append(nn.ConvTranspose2d(c_in, 384, kernel_size=k, stride=1, padding=1))
                break
            c_out = (2 ** (n - i - 2)) * self.channel_step
            mods += [norm_layer(1, c_in),
                     nn.ConvTranspose2d(c_in, c_out, kernel_size=k, stride=2, padding=2, output_padding=1),
                     nn.ELU(inplace=True)]
            c_in = c_out
        self.net = nn.Sequential(*mods)

    def forward(self, X):
        y = self.net(self.convin(X).view(-1, 32 * self.channel_step, 1, 1))
        return td.Independent(td.Normal(y, torch.ones((), device=y.device, dtype=y.dtype)), 3)
