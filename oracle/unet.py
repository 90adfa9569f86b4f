"""ORACLE (test infrastructure only) — CPU fp32 restatement of the reference's stage-2 UNet.

Follows /root/reference/src/models/stage2_inpaint_unet_2d_condition.py:
  * topology  : __init__ :166-448 (conv_in :168-170, time/class embedding :184-247, down :314-343, mid :348-361,
                up :381-429, out :432-448) for the SD-2.1-base config with the overrides the batch-test driver
                passes (stage2_batchtest_inpaint_model.py:125-128: in_channels=9, class_embed_type="projection",
                projection_class_embeddings_input_dim=1024);
  * forward   : :579-825 — timestep broadcast :661-675, sinusoid + cast :677-682, time MLP :684, class embedding
                :687-708, `conv_in(sample) + my_pose_cond` :742 (the one functional change vs stock diffusers),
                skip bookkeeping :747-761, mid :775-783, up blocks popping skips :789-814, GroupNorm/SiLU/conv_out
                :817-820.
The building blocks live in oracle/blocks.py (restated from diffusers 0.24.0).  State-dict keys equal the diffusers
keys, so the product's weight loader and this oracle consume the same checkpoint dict.
Only `tests/`, `__graft_entry__.smoke()` and bench.py's CPU-baseline leg may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import blocks as B


@dataclass
class UNetConfig:
    """The subset of diffusers' UNet2DConditionModel config that the reference path exercises."""
    sample_size: Optional[int] = 64
    in_channels: int = 9
    out_channels: int = 4
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                         "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    downsample_padding: int = 1
    act_fn: str = "silu"
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    cross_attention_dim: int = 1024
    attention_head_dim: Tuple[int, ...] = (5, 10, 20, 20)  # = number of heads (diffusers naming quirk, ref :122-128)
    use_linear_projection: bool = True
    class_embed_type: Optional[str] = "projection"
    projection_class_embeddings_input_dim: Optional[int] = 1024
    use_pose_cond: bool = True  # stage-2: forward() requires my_pose_cond; stage-3 (stock UNet) has none
    time_cond_proj_dim: Optional[int] = None
    _diffusers_version: str = "0.24.0"

    @staticmethod
    def stage2() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def stage3() -> "UNetConfig":  # stock UNet2DConditionModel(in_channels=8), stage3_batchtest_refined_model.py:121-122
        return UNetConfig(in_channels=8, class_embed_type=None, projection_class_embeddings_input_dim=None,
                          use_pose_cond=False)

    @staticmethod
    def sd21_stock() -> "UNetConfig":  # published SD-2.1-base UNet: 865 910 724 parameters
        return UNetConfig(in_channels=4, class_embed_type=None, projection_class_embeddings_input_dim=None,
                          use_pose_cond=False)

    @staticmethod
    def tiny(in_channels: int = 9, stage2: bool = True) -> "UNetConfig":
        """Same topology, 64/128/256/256 channels, 1/2/4/4 heads of dim 64: for fast tests."""
        return UNetConfig(in_channels=in_channels, block_out_channels=(64, 128, 256, 256),
                          attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128,
                          class_embed_type="projection" if stage2 else None,
                          projection_class_embeddings_input_dim=128 if stage2 else None, use_pose_cond=stage2,
                          sample_size=32)


class OracleUNet(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        heads = cfg.attention_head_dim
        time_embed_dim = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], kernel_size=3, padding=1)
        self.time_proj = B.Timesteps(ch[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        self.time_embedding = B.TimestepEmbedding(ch[0], time_embed_dim, act_fn=cfg.act_fn)
        if cfg.class_embed_type == "projection":
            self.class_embedding = B.TimestepEmbedding(cfg.projection_class_embeddings_input_dim, time_embed_dim)
        elif cfg.class_embed_type is None:
            self.class_embedding = None
        else:
            raise NotImplementedError(cfg.class_embed_type)

        self.down_blocks = nn.ModuleList()
        out_c = ch[0]
        for i, kind in enumerate(cfg.down_block_types):
            in_c, out_c = out_c, ch[i]
            last = i == len(ch) - 1
            self.down_blocks.append(B.get_down_block(
                kind, num_layers=cfg.layers_per_block, in_channels=in_c, out_channels=out_c,
                temb_channels=time_embed_dim, add_downsample=not last, resnet_eps=cfg.norm_eps,
                resnet_act_fn=cfg.act_fn, resnet_groups=cfg.norm_num_groups,
                cross_attention_dim=cfg.cross_attention_dim, num_attention_heads=heads[i],
                downsample_padding=cfg.downsample_padding, use_linear_projection=cfg.use_linear_projection))
        self.mid_block = B.UNetMidBlock2DCrossAttn(
            in_channels=ch[-1], temb_channels=time_embed_dim, resnet_eps=cfg.norm_eps, resnet_act_fn=cfg.act_fn,
            output_scale_factor=1, cross_attention_dim=cfg.cross_attention_dim, num_attention_heads=heads[-1],
            resnet_groups=cfg.norm_num_groups, use_linear_projection=cfg.use_linear_projection)
        self.up_blocks = nn.ModuleList()
        rch, rheads = list(reversed(ch)), list(reversed(heads))
        out_c = rch[0]
        self.num_upsamplers = 0
        for i, kind in enumerate(cfg.up_block_types):
            last = i == len(ch) - 1
            prev_c, out_c = out_c, rch[i]
            in_c = rch[min(i + 1, len(ch) - 1)]
            if not last:
                self.num_upsamplers += 1
            self.up_blocks.append(B.get_up_block(
                kind, num_layers=cfg.layers_per_block + 1, in_channels=in_c, out_channels=out_c,
                prev_output_channel=prev_c, temb_channels=time_embed_dim, add_upsample=not last,
                resnet_eps=cfg.norm_eps, resnet_act_fn=cfg.act_fn, resnet_groups=cfg.norm_num_groups,
                cross_attention_dim=cfg.cross_attention_dim, num_attention_heads=rheads[i],
                use_linear_projection=cfg.use_linear_projection))
        self.conv_norm_out = nn.GroupNorm(num_channels=ch[0], num_groups=cfg.norm_num_groups, eps=cfg.norm_eps)
        self.conv_act = B.get_activation(cfg.act_fn)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, kernel_size=3, padding=1)

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, my_pose_cond=None,
                return_dict: bool = False):
        cfg = self.cfg
        # 1. time (ref :661-684)
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float64 if isinstance(t, float) else torch.int64, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t.expand(sample.shape[0])
        t_emb = self.time_proj(t).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        # class embedding (ref :687-708)
        if self.class_embedding is not None:
            if class_labels is None:
                raise ValueError("class_labels should be provided when num_class_embeds > 0")
            class_emb = self.class_embedding(class_labels.squeeze(1)).to(dtype=sample.dtype)
            emb = emb + class_emb
        # 2. pre-process (ref :742)
        sample = self.conv_in(sample)
        if cfg.use_pose_cond:
            sample = sample + my_pose_cond
        # 3. down (ref :747-761)
        skips = (sample,)
        for blk in self.down_blocks:
            if getattr(blk, "has_cross_attention", False):
                sample, res = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
            else:
                sample, res = blk(hidden_states=sample, temb=emb)
            skips += res
        # 4. mid (ref :775-783)
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        # 5. up (ref :789-814)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            res, skips = skips[-n:], skips[:-n]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             encoder_hidden_states=encoder_hidden_states)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res)
        # 6. post-process (ref :817-820)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        return (sample,)


def param_count(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
