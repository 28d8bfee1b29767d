"""ORACLE — test infrastructure only. Plain-PyTorch restatement of `AutoencoderKLTemporalDecoder.encode` as
`train_svd.py` uses it (tensor_to_vae_latent, /root/reference/train_svd.py:283-291; called twice per step at :948 and :959):

    latents = vae.encode(frames).latent_dist.sample() * vae.config.scaling_factor

[D] = diffusers models/autoencoders/autoencoder_kl_temporal_decoder.py (encode path: Encoder + quant_conv +
DiagonalGaussianDistribution) and models/autoencoders/vae.py Encoder / models/unets/unet_2d_blocks.py DownEncoderBlock2D,
UNetMidBlock2D / models/resnet.py ResnetBlock2D / models/downsampling.py Downsample2D / models/attention_processor.py
Attention — absent from /root/reference, restated from the published algorithm: PARITY UNPINNED against diffusers (no tests or
golden vectors in the reference; tests/test_oracle_vs_diffusers.py A/Bs this file wherever diffusers is importable). The
temporal DECODER is not restated (outside SURVEY.md §8f-1). Module / parameter names follow the diffusers state dict
(`encoder.conv_in`, `encoder.down_blocks.i.resnets.j.*`, `encoder.mid_block.attentions.0.*`, `quant_conv`, ...).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

VAE_CONFIG = dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2, scaling_factor=0.18215)
TINY_VAE_CONFIG = dict(in_channels=3, latent_channels=4, block_out_channels=(64, 128), layers_per_block=1, scaling_factor=0.18215)


class ResnetBlock2D(nn.Module):
    """[D] resnet.py ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6, swish, output_scale_factor=1)."""

    def __init__(self, in_channels, out_channels, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb=None):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    """[D] downsampling.py Downsample2D(use_conv=True, padding=0): pad (0,1,0,1) then a stride-2 3x3 conv without padding."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = d(x)
        return x


class VaeAttention(nn.Module):
    """[D] attention_processor.py Attention as UNetMidBlock2D builds it: one head of dim C, GroupNorm(32, eps 1e-6) on the
    input, biased q/k/v/out projections, residual connection, rescale_output_factor 1."""

    def __init__(self, channels):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, channels, eps=1e-6, affine=True)
        self.to_q = nn.Linear(channels, channels, bias=True)
        self.to_k = nn.Linear(channels, channels, bias=True)
        self.to_v = nn.Linear(channels, channels, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels, bias=True), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        residual = x
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)            # [b, hw, c]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        attn = torch.softmax((q @ k.transpose(1, 2)) * (c ** -0.5), dim=-1)
        o = self.to_out[0](attn @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + residual


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels), ResnetBlock2D(channels, channels)])
        self.attentions = nn.ModuleList([VaeAttention(channels)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    """[D] vae.py Encoder(double_z=True, act_fn='silu', norm_num_groups=32, mid_block_add_attention=True)."""

    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([])
        oc = block_out_channels[0]
        for i, boc in enumerate(block_out_channels):
            ic, oc = oc, boc
            self.down_blocks.append(DownEncoderBlock2D(ic, oc, layers_per_block, add_downsample=i != len(block_out_channels) - 1))
        self.mid_block = UNetMidBlock2D(block_out_channels[-1])
        self.conv_norm_out = nn.GroupNorm(32, block_out_channels[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[-1], 2 * out_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    """[D] vae.py DiagonalGaussianDistribution."""

    def __init__(self, parameters):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class AutoencoderKLTemporalDecoder(nn.Module):
    """encode path only: encoder + quant_conv -> latent_dist."""

    def __init__(self, in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2, scaling_factor=0.18215):
        super().__init__()
        self.config = SimpleNamespace(in_channels=in_channels, latent_channels=latent_channels, block_out_channels=tuple(block_out_channels),
                                      layers_per_block=layers_per_block, scaling_factor=scaling_factor)
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))


def tensor_to_vae_latent(t, vae, noise=None):
    """train_svd.py:283-291."""
    b, f = t.shape[:2]
    latents = vae.encode(t.flatten(0, 1)).latent_dist.sample(noise=noise)
    return latents.reshape(b, f, *latents.shape[1:]) * vae.config.scaling_factor
