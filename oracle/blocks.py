"""ORACLE (test infrastructure, never shipped or measured as the product) — CPU/PyTorch fp32 restatement of the
diffusers-0.24.0 building blocks that the reference's stage-2 UNet instantiates.

The reference does not vendor these; it imports them from the pinned third-party package `diffusers==0.24.0`
(/root/reference/README.md:37) at src/models/stage2_inpaint_unet_2d_condition.py:21-44 and builds them at :321-343
(down blocks), :348-361 (mid block), :407-429 (up blocks), :184-197,247 (time / class embeddings).  diffusers is not
installed here and there is no network, so the published algorithm of that release is restated below, class by class,
with the same constructor arguments (as far as the reference passes them), module names and state-dict keys, so a
real SD-2.1 checkpoint would load.  PARITY UNPINNED against diffusers itself: the reference ships no tests or golden
vectors for this path; what IS pinned is (a) the reference's own UNet class running unmodified on top
of these blocks (tests/test_oracle_reference_shim.py), (b) closed-form checks of every block (tests/test_oracle.py)
and (c) the published SD-2.x parameter count, 865 910 724, which fixes the topology.

Only the configuration the reference uses is implemented (SD-2.1-base unet/config.json + the overrides at
stage2_batchtest_inpaint_model.py:125-128); anything else raises NotImplementedError rather than guessing.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def get_activation(name: str) -> nn.Module:
    name = name.lower()
    if name in ("silu", "swish"):
        return nn.SiLU()
    if name == "gelu":
        return nn.GELU()
    if name == "mish":
        return nn.Mish()
    if name == "relu":
        return nn.ReLU()
    raise ValueError(f"unsupported activation {name}")


# ------------------------------------------------------------------------------------------------------------
# embeddings (diffusers.models.embeddings) — SURVEY.md App. A.1
# ------------------------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool = False,
                           downscale_freq_shift: float = 1, scale: float = 1, max_period: int = 10000):
    assert timesteps.dim() == 1
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu", out_dim: Optional[int] = None,
                 post_act_fn: Optional[str] = None, cond_proj_dim: Optional[int] = None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.cond_proj = nn.Linear(cond_proj_dim, in_channels, bias=False) if cond_proj_dim is not None else None
        self.act = get_activation(act_fn)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)
        self.post_act = get_activation(post_act_fn) if post_act_fn is not None else None

    def forward(self, sample, condition=None):
        if condition is not None:
            sample = sample + self.cond_proj(condition)
        sample = self.linear_1(sample)
        sample = self.act(sample)
        sample = self.linear_2(sample)
        if self.post_act is not None:
            sample = self.post_act(sample)
        return sample


# ------------------------------------------------------------------------------------------------------------
# resnet / resampling (diffusers.models.resnet) — App. A.2, A.5
# ------------------------------------------------------------------------------------------------------------
class Upsample2D(nn.Module):
    def __init__(self, channels: int, use_conv: bool = False, use_conv_transpose: bool = False,
                 out_channels: Optional[int] = None, name: str = "conv"):
        super().__init__()
        if use_conv_transpose or not use_conv:
            raise NotImplementedError("oracle: only Upsample2D(use_conv=True) is on the reference path")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None, scale: float = 1.0):
        assert hidden_states.shape[1] == self.channels
        if output_size is None:
            hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        else:
            hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        return self.conv(hidden_states)


class Downsample2D(nn.Module):
    def __init__(self, channels: int, use_conv: bool = False, out_channels: Optional[int] = None, padding: int = 1,
                 name: str = "conv"):
        super().__init__()
        if not use_conv:
            raise NotImplementedError("oracle: only Downsample2D(use_conv=True) is on the reference path")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.padding = padding
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, hidden_states, scale: float = 1.0):
        assert hidden_states.shape[1] == self.channels
        if self.padding == 0:
            hidden_states = F.pad(hidden_states, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(hidden_states)


class ResnetBlock2D(nn.Module):
    def __init__(self, *, in_channels: int, out_channels: Optional[int] = None, conv_shortcut: bool = False,
                 dropout: float = 0.0, temb_channels: int = 512, groups: int = 32, groups_out: Optional[int] = None,
                 pre_norm: bool = True, eps: float = 1e-6, non_linearity: str = "swish", skip_time_act: bool = False,
                 time_embedding_norm: str = "default", output_scale_factor: float = 1.0,
                 use_in_shortcut: Optional[bool] = None, up: bool = False, down: bool = False):
        super().__init__()
        if time_embedding_norm != "default" or up or down or skip_time_act:
            raise NotImplementedError("oracle: only the 'default' ResnetBlock2D variant is on the reference path")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.output_scale_factor = output_scale_factor
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(num_groups=groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = get_activation(non_linearity)
        use_in_shortcut = in_channels != out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = (nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
                              if use_in_shortcut else None)

    def forward(self, input_tensor, temb, scale: float = 1.0):
        h = self.norm1(input_tensor)
        h = self.nonlinearity(h)
        h = self.conv1(h)
        if self.time_emb_proj is not None:
            t = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
            h = h + t
        h = self.norm2(h)
        h = self.nonlinearity(h)
        h = self.dropout(h)
        h = self.conv2(h)
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


# ------------------------------------------------------------------------------------------------------------
# attention (diffusers.models.attention_processor / attention / transformer_2d) — App. A.3, A.4
# ------------------------------------------------------------------------------------------------------------
class AttnProcessor:
    """Default processor: softmax(Q K^T * scale) V with explicit score matrix (attention_processor.AttnProcessor)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0):
        residual = hidden_states
        if attn.spatial_norm is not None or attn.group_norm is not None:
            raise NotImplementedError
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b, c, h, w = hidden_states.shape
            hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        batch_size, sequence_length, _ = (hidden_states.shape if encoder_hidden_states is None
                                          else encoder_hidden_states.shape)
        attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        elif attn.norm_cross:
            encoder_hidden_states = attn.norm_encoder_hidden_states(encoder_hidden_states)
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        query = attn.head_to_batch_dim(query)
        key = attn.head_to_batch_dim(key)
        value = attn.head_to_batch_dim(value)
        attention_probs = attn.get_attention_scores(query, key, attention_mask)
        hidden_states = torch.bmm(attention_probs, value)
        hidden_states = attn.batch_to_head_dim(hidden_states)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(b, c, h, w)
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        return hidden_states / attn.rescale_output_factor


AttentionProcessor = AttnProcessor


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias: bool = False, upcast_attention: bool = False,
                 upcast_softmax: bool = False, cross_attention_norm: Optional[str] = None,
                 only_cross_attention: bool = False, processor=None, out_bias: bool = True):
        super().__init__()
        if cross_attention_norm is not None or only_cross_attention:
            raise NotImplementedError
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention, self.upcast_softmax = upcast_attention, upcast_softmax
        self.rescale_output_factor = 1.0
        self.residual_connection = False
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else AttnProcessor())

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self):
        return self.processor

    def set_use_memory_efficient_attention_xformers(self, *args, **kwargs):  # harmless no-op (xformers absent)
        return None

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)

    def head_to_batch_dim(self, tensor):
        b, s, d = tensor.shape
        h = self.heads
        return tensor.reshape(b, s, h, d // h).permute(0, 2, 1, 3).reshape(b * h, s, d // h)

    def batch_to_head_dim(self, tensor):
        bh, s, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        scores = torch.bmm(query, key.transpose(-1, -2)) * self.scale
        if attention_mask is not None:
            scores = scores + attention_mask
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return None
        raise NotImplementedError("oracle: attention masks are never passed on the reference path")


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)  # exact (erf) GELU


class GELU(nn.Module):
    """diffusers.models.activations.GELU (approximate="none"): Linear then exact (erf) GELU — the `activation_fn="gelu"`
    feed-forward of the stage-1 prior's blocks (/root/reference/src/models/stage1_prior_transformer.py:112-120)."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, hidden_states, scale: float = 1.0):
        return F.gelu(self.proj(hidden_states))


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, dropout: float = 0.0,
                 activation_fn: str = "geglu", final_dropout: bool = False):
        super().__init__()
        if activation_fn not in ("geglu", "gelu"):
            raise NotImplementedError
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        act = GEGLU(dim, inner_dim) if activation_fn == "geglu" else GELU(dim, inner_dim)
        self.net = nn.ModuleList([act, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])

    def forward(self, hidden_states, scale: float = 1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, dropout=0.0,
                 cross_attention_dim: Optional[int] = None, activation_fn: str = "geglu",
                 attention_bias: bool = False, only_cross_attention: bool = False, double_self_attention: bool = False,
                 upcast_attention: bool = False, norm_elementwise_affine: bool = True, norm_type: str = "layer_norm",
                 **unused):
        super().__init__()
        if only_cross_attention or double_self_attention or norm_type != "layer_norm":
            raise NotImplementedError
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                               bias=attention_bias, upcast_attention=upcast_attention)
        if cross_attention_dim is not None:
            self.norm2 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
            self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                                   dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                                   upcast_attention=upcast_attention)
        else:   # diffusers 0.24.0: no second attention without a cross_attention_dim (the stage-1 prior's blocks)
            self.norm2 = None
            self.attn2 = None
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        kw = cross_attention_kwargs if cross_attention_kwargs is not None else {}
        n = self.norm1(hidden_states)
        hidden_states = self.attn1(n, encoder_hidden_states=None, attention_mask=attention_mask, **kw) + hidden_states
        if self.attn2 is not None:
            n = self.norm2(hidden_states)
            hidden_states = self.attn2(n, encoder_hidden_states=encoder_hidden_states,
                                       attention_mask=encoder_attention_mask, **kw) + hidden_states
        n = self.norm3(hidden_states)
        hidden_states = self.ff(n) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels: Optional[int] = None,
                 out_channels: Optional[int] = None, num_layers: int = 1, dropout: float = 0.0,
                 norm_num_groups: int = 32, cross_attention_dim: Optional[int] = None, attention_bias: bool = False,
                 use_linear_projection: bool = False, only_cross_attention: bool = False,
                 upcast_attention: bool = False, **unused):
        super().__init__()
        if not use_linear_projection:
            raise NotImplementedError("oracle: SD-2.x uses use_linear_projection=True")
        inner_dim = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, attention_bias=attention_bias,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner_dim, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None,
                return_dict: bool = True):
        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        inner_dim = hidden_states.shape[1]
        hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, attention_mask=attention_mask,
                                  encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=encoder_attention_mask, timestep=timestep,
                                  cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        hidden_states = self.proj_out(hidden_states)
        hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        output = hidden_states + residual
        if not return_dict:
            return (output,)
        return _Out(sample=output)


class _Out:
    def __init__(self, sample):
        self.sample = sample


# ------------------------------------------------------------------------------------------------------------
# UNet blocks (diffusers.models.unet_2d_blocks)
# ------------------------------------------------------------------------------------------------------------
class DownBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, dropout: float = 0.0,
                 num_layers: int = 1, resnet_eps: float = 1e-6, resnet_time_scale_shift: str = "default",
                 resnet_act_fn: str = "swish", resnet_groups: int = 32, resnet_pre_norm: bool = True,
                 output_scale_factor=1.0, add_downsample=True, downsample_padding=1):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                          temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups, dropout=dropout,
                          time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                          output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm)
            for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=downsample_padding, name="op")])
                             if add_downsample else None)
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, scale: float = 1.0):
        output_states = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, dropout: float = 0.0,
                 num_layers: int = 1, transformer_layers_per_block: int = 1, resnet_eps: float = 1e-6,
                 resnet_time_scale_shift: str = "default", resnet_act_fn: str = "swish", resnet_groups: int = 32,
                 resnet_pre_norm: bool = True, num_attention_heads=1, cross_attention_dim=1280,
                 output_scale_factor=1.0, downsample_padding=1, add_downsample=True, dual_cross_attention=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False, **unused):
        super().__init__()
        if dual_cross_attention:
            raise NotImplementedError
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnets, attentions = [], []
        for i in range(num_layers):
            resnets.append(ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels,
                                         out_channels=out_channels, temb_channels=temb_channels, eps=resnet_eps,
                                         groups=resnet_groups, dropout=dropout,
                                         time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                         output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm))
            attentions.append(Transformer2DModel(num_attention_heads, out_channels // num_attention_heads,
                                                 in_channels=out_channels, num_layers=transformer_layers_per_block,
                                                 cross_attention_dim=cross_attention_dim,
                                                 norm_num_groups=resnet_groups,
                                                 use_linear_projection=use_linear_projection,
                                                 only_cross_attention=only_cross_attention,
                                                 upcast_attention=upcast_attention))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=downsample_padding, name="op")])
                             if add_downsample else None)
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None, additional_residuals=None):
        output_states = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, in_channels: int, temb_channels: int, dropout: float = 0.0, num_layers: int = 1,
                 transformer_layers_per_block: int = 1, resnet_eps: float = 1e-6,
                 resnet_time_scale_shift: str = "default", resnet_act_fn: str = "swish", resnet_groups: int = 32,
                 resnet_pre_norm: bool = True, num_attention_heads=1, output_scale_factor=1.0,
                 cross_attention_dim=1280, dual_cross_attention=False, use_linear_projection=False,
                 upcast_attention=False, **unused):
        super().__init__()
        if dual_cross_attention:
            raise NotImplementedError
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnet_groups = resnet_groups if resnet_groups is not None else min(in_channels // 4, 32)

        def _res():
            return ResnetBlock2D(in_channels=in_channels, out_channels=in_channels, temb_channels=temb_channels,
                                 eps=resnet_eps, groups=resnet_groups, dropout=dropout,
                                 time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                 output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm)

        resnets, attentions = [_res()], []
        for _ in range(num_layers):
            attentions.append(Transformer2DModel(num_attention_heads, in_channels // num_attention_heads,
                                                 in_channels=in_channels, num_layers=transformer_layers_per_block,
                                                 cross_attention_dim=cross_attention_dim,
                                                 norm_num_groups=resnet_groups,
                                                 use_linear_projection=use_linear_projection,
                                                 upcast_attention=upcast_attention))
            resnets.append(_res())
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class UNetMidBlock2DSimpleCrossAttn(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("oracle: UNetMidBlock2DSimpleCrossAttn is not on the reference path")


class UpBlock2D(nn.Module):
    def __init__(self, in_channels: int, prev_output_channel: int, out_channels: int, temb_channels: int,
                 resolution_idx: Optional[int] = None, dropout: float = 0.0, num_layers: int = 1,
                 resnet_eps: float = 1e-6, resnet_time_scale_shift: str = "default", resnet_act_fn: str = "swish",
                 resnet_groups: int = 32, resnet_pre_norm: bool = True, output_scale_factor=1.0, add_upsample=True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(in_channels=resnet_in_channels + res_skip_channels,
                                         out_channels=out_channels, temb_channels=temb_channels, eps=resnet_eps,
                                         groups=resnet_groups, dropout=dropout,
                                         time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                         output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.gradient_checkpointing = False

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, scale: float = 1.0):
        for resnet in self.resnets:
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class CrossAttnUpBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, prev_output_channel: int, temb_channels: int,
                 resolution_idx: Optional[int] = None, dropout: float = 0.0, num_layers: int = 1,
                 transformer_layers_per_block: int = 1, resnet_eps: float = 1e-6,
                 resnet_time_scale_shift: str = "default", resnet_act_fn: str = "swish", resnet_groups: int = 32,
                 resnet_pre_norm: bool = True, num_attention_heads=1, cross_attention_dim=1280,
                 output_scale_factor=1.0, add_upsample=True, dual_cross_attention=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False, **unused):
        super().__init__()
        if dual_cross_attention:
            raise NotImplementedError
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnets, attentions = [], []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(in_channels=resnet_in_channels + res_skip_channels,
                                         out_channels=out_channels, temb_channels=temb_channels, eps=resnet_eps,
                                         groups=resnet_groups, dropout=dropout,
                                         time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                         output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm))
            attentions.append(Transformer2DModel(num_attention_heads, out_channels // num_attention_heads,
                                                 in_channels=out_channels, num_layers=transformer_layers_per_block,
                                                 cross_attention_dim=cross_attention_dim,
                                                 norm_num_groups=resnet_groups,
                                                 use_linear_projection=use_linear_projection,
                                                 only_cross_attention=only_cross_attention,
                                                 upcast_attention=upcast_attention))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.gradient_checkpointing = False

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None, attention_mask=None, encoder_attention_mask=None):
        for resnet, attn in zip(self.resnets, self.attentions):
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                   resnet_act_fn, transformer_layers_per_block=1, num_attention_heads=None, resnet_groups=None,
                   cross_attention_dim=None, downsample_padding=None, dual_cross_attention=False,
                   use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                   resnet_time_scale_shift="default", attention_type="default", resnet_skip_time_act=False,
                   resnet_out_scale_factor=1.0, cross_attention_norm=None, attention_head_dim=None,
                   downsample_type=None, dropout=0.0):
    if down_block_type.startswith("UNetRes"):
        down_block_type = down_block_type[7:]
    if down_block_type == "DownBlock2D":
        return DownBlock2D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                           temb_channels=temb_channels, dropout=dropout, add_downsample=add_downsample,
                           resnet_eps=resnet_eps, resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                           downsample_padding=downsample_padding, resnet_time_scale_shift=resnet_time_scale_shift)
    if down_block_type == "CrossAttnDownBlock2D":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnDownBlock2D")
        return CrossAttnDownBlock2D(num_layers=num_layers, transformer_layers_per_block=transformer_layers_per_block,
                                    in_channels=in_channels, out_channels=out_channels, temb_channels=temb_channels,
                                    dropout=dropout, add_downsample=add_downsample, resnet_eps=resnet_eps,
                                    resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                                    downsample_padding=downsample_padding, cross_attention_dim=cross_attention_dim,
                                    num_attention_heads=num_attention_heads,
                                    dual_cross_attention=dual_cross_attention,
                                    use_linear_projection=use_linear_projection,
                                    only_cross_attention=only_cross_attention, upcast_attention=upcast_attention,
                                    resnet_time_scale_shift=resnet_time_scale_shift)
    raise NotImplementedError(f"oracle: down block {down_block_type} is not on the reference path")


def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                 add_upsample, resnet_eps, resnet_act_fn, resolution_idx=None, transformer_layers_per_block=1,
                 num_attention_heads=None, resnet_groups=None, cross_attention_dim=None, dual_cross_attention=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 resnet_time_scale_shift="default", attention_type="default", resnet_skip_time_act=False,
                 resnet_out_scale_factor=1.0, cross_attention_norm=None, attention_head_dim=None, upsample_type=None,
                 dropout=0.0):
    if up_block_type.startswith("UNetRes"):
        up_block_type = up_block_type[7:]
    if up_block_type == "UpBlock2D":
        return UpBlock2D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                         prev_output_channel=prev_output_channel, temb_channels=temb_channels, dropout=dropout,
                         add_upsample=add_upsample, resnet_eps=resnet_eps, resnet_act_fn=resnet_act_fn,
                         resnet_groups=resnet_groups, resnet_time_scale_shift=resnet_time_scale_shift)
    if up_block_type == "CrossAttnUpBlock2D":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnUpBlock2D")
        return CrossAttnUpBlock2D(num_layers=num_layers, transformer_layers_per_block=transformer_layers_per_block,
                                  in_channels=in_channels, out_channels=out_channels,
                                  prev_output_channel=prev_output_channel, temb_channels=temb_channels,
                                  dropout=dropout, add_upsample=add_upsample, resnet_eps=resnet_eps,
                                  resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                                  cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads,
                                  dual_cross_attention=dual_cross_attention,
                                  use_linear_projection=use_linear_projection,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention,
                                  resnet_time_scale_shift=resnet_time_scale_shift)
    raise NotImplementedError(f"oracle: up block {up_block_type} is not on the reference path")
