"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of diffusers 0.24.0 `AutoencoderKL` as configured for
stable-diffusion-2-1-base (`vae/config.json`: block_out_channels (128, 256, 512, 512), layers_per_block 2,
latent_channels 4, norm_num_groups 32, act silu, scaling_factor 0.18215), the module the reference's pipelines call at
/root/reference/src/pipelines/stage2_inpaint_pipeline.py:443 (`vae.encode(...).latent_dist.sample`) and :528
(`vae.decode(latents / scaling_factor)`), and stage3_refined_pipeline.py:479,563.

diffusers is not vendored in /root/reference: the blocks follow the published implementation (`models/vae.py`
Encoder / Decoder / DiagonalGaussianDistribution, `models/unet_2d_blocks.py` DownEncoderBlock2D / UpDecoderBlock2D /
UNetMidBlock2D, `models/attention_processor.py` Attention with group_norm + residual_connection, one head of dim 512),
with the same state-dict key names so a real `vae/diffusion_pytorch_model.*` loads.  PARITY UNPINNED against diffusers
itself; pinned by the published parameter count of the SD VAE (83 653 863) and closed-form block checks
(tests/test_vae.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import Downsample2D, ResnetBlock2D, Upsample2D


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_num_groups: int = 32
    act_fn: str = "silu"
    sample_size: int = 512
    scaling_factor: float = 0.18215

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(64, 128, 128, 128), layers_per_block=1)


class VAEAttention(nn.Module):
    """diffusers Attention(channels, heads=channels // attention_head_dim (= 1), dim_head=channels, eps=1e-6,
    norm_num_groups=32, residual_connection=True, bias=True, upcast_softmax=True) run by AttnProcessor on a 4-D input."""

    def __init__(self, channels: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(num_channels=channels, num_groups=groups, eps=1e-6, affine=True)
        self.to_q = nn.Linear(channels, channels, bias=True)
        self.to_k = nn.Linear(channels, channels, bias=True)
        self.to_v = nn.Linear(channels, channels, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels, bias=True), nn.Dropout(0.0)])
        self.scale = channels ** -0.5

    def forward(self, x):
        b, c, h, w = x.shape
        residual = x
        hs = x.view(b, c, h * w).transpose(1, 2)
        hs = self.group_norm(hs.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(hs), self.to_k(hs), self.to_v(hs)
        probs = (torch.bmm(q, k.transpose(-1, -2)) * self.scale).float().softmax(dim=-1).to(q.dtype)
        hs = self.to_out[0](torch.bmm(probs, v))
        hs = hs.transpose(-1, -2).reshape(b, c, h, w)
        return hs + residual


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels: int, groups: int):
        super().__init__()
        mk = lambda: ResnetBlock2D(in_channels=channels, out_channels=channels, temb_channels=None, eps=1e-6,
                                   groups=groups)
        self.attentions = nn.ModuleList([VAEAttention(channels, groups)])
        self.resnets = nn.ModuleList([mk(), mk()])

    def forward(self, x):
        x = self.resnets[0](x, None)
        x = self.attentions[0](x)
        return self.resnets[1](x, None)


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=cin if i == 0 else cout, out_channels=cout,
                                                    temb_channels=None, eps=1e-6, groups=groups)
                                      for i in range(layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(cout, use_conv=True, out_channels=cout, padding=0)])
                             if add_downsample else None)

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=cin if i == 0 else cout, out_channels=cout,
                                                    temb_channels=None, eps=1e-6, groups=groups)
                                      for i in range(layers)])
        self.upsamplers = (nn.ModuleList([Upsample2D(cout, use_conv=True, out_channels=cout)])
                           if add_upsample else None)

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_c = ch[0]
        for i, c in enumerate(ch):
            in_c, out_c = out_c, c
            self.down_blocks.append(DownEncoderBlock2D(in_c, out_c, cfg.layers_per_block, g, i < len(ch) - 1))
        self.mid_block = UNetMidBlock2D(ch[-1], g)
        self.conv_norm_out = nn.GroupNorm(num_channels=ch[-1], num_groups=g, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[-1], 3, padding=1)
        self.mid_block = UNetMidBlock2D(ch[-1], g)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        out_c = rev[0]
        for i, c in enumerate(rev):
            prev, out_c = out_c, c
            self.up_blocks.append(UpDecoderBlock2D(prev, out_c, cfg.layers_per_block + 1, g, i < len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(num_channels=ch[0], num_groups=g, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class OracleAutoencoderKL(nn.Module):
    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.config = SimpleNamespace(**cfg.__dict__)
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z, return_dict=True):
        dec = self.decoder(self.post_quant_conv(z))
        return (dec,) if not return_dict else SimpleNamespace(sample=dec)


def make_vae(cfg: VAEConfig = VAEConfig(), seed: int = 0) -> OracleAutoencoderKL:
    """Seeded random weights with O(1) activations (norm affines perturbed so that they are exercised)."""
    torch.manual_seed(seed)
    m = OracleAutoencoderKL(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "norm" in name:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return m
