"""B200CLIPVisionModelWithProjection — drop-in for the `transformers.CLIPVisionModelWithProjection` the reference's
drivers load from OpenCLIP-ViT-H-14:
  * stage 1: /root/reference/stage1_batchtest_prior_model.py:61 (`from_pretrained(args.image_encoder_path)`), applied
    at :100-101 (`image_encoder(pixels).image_embeds` of the source and target image), and
    src/pipelines/stage1_prior_pipeline.py:282-289 (`get_zero_embed`);
  * stage 2: stage2_batchtest_inpaint_model.py:97 (`image_encoder_g`), applied at :181-183 for the "train" split.
SURVEY.md §8f-3.

Same config keys, same state-dict key names (`vision_model.embeddings.*`, `vision_model.pre_layrnorm.*`,
`vision_model.encoder.layers.N.*`, `vision_model.post_layernorm.*`, `visual_projection.weight`), same call surface
(`model(pixel_values).image_embeds / .last_hidden_state / .pooler_output`, item access as at
stage1_prior_pipeline.py:287).  Everything after the patch unfold runs on the sm_100a kernels of libpcdm_b200.so:
patch projection, fused q/k/v, attention output, fc1 (GELU in the GEMM epilogue), fc2 (residual in the epilogue) and
the visual projection as tcgen05 GEMMs; LayerNorm; flash attention.  ViT-H/14 has 16 heads of **80** channels: at load
the q/k/v projection rows (and the out_proj columns) of every head are zero-padded to 128, the width the attention
kernel implements (`pcdm_attention_hd`) — zero q/k columns add nothing to a score, zero v columns produce zero outputs
that meet zero out_proj columns — and the softmax scale stays 80^-1/2.  No PyTorch / CPU compute fallback.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import torch

from . import ops
from .arena import WeightArenaMixin
from .unet import _Config

# OpenCLIP ViT-H/14 (laion2B) vision tower as shipped in the HF checkpoint the reference points at
_DEFAULT_CONFIG = dict(hidden_size=1280, intermediate_size=5120, projection_dim=1024, num_hidden_layers=32,
                       num_attention_heads=16, num_channels=3, image_size=224, patch_size=14, hidden_act="gelu",
                       layer_norm_eps=1e-5, attention_dropout=0.0)


def _pad(n, m):
    return (n + m - 1) // m * m


class CLIPVisionOutput(dict):
    """Attribute + item access, like transformers' ModelOutput (`out.image_embeds`, `out["image_embeds"]`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class B200CLIPVisionModelWithProjection(WeightArenaMixin):
    def __init__(self, config=None, dtype: torch.dtype = torch.float16, device="cuda", **kw):
        cfg = dict(_DEFAULT_CONFIG)
        src = dict(config.to_dict() if hasattr(config, "to_dict") else (config or {}))
        if "vision_config" in src and isinstance(src["vision_config"], dict):   # a full CLIPConfig json
            proj = src.get("projection_dim")
            src = dict(src["vision_config"])
            if proj is not None:
                src.setdefault("projection_dim", proj)
        src.update(kw)
        cfg.update({k: v for k, v in src.items() if k in cfg})
        c = _Config(cfg)

        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"pcdm_b200 CLIP vision: unsupported config ({what})")
        need(c.hidden_act == "gelu", "hidden_act must be 'gelu' (ViT-H/14; quick_gelu is not implemented)")
        need(c.hidden_size % c.num_attention_heads == 0, "hidden_size % heads")
        self.head_dim = c.hidden_size // c.num_attention_heads
        need(self.head_dim <= 128 and self.head_dim % 8 == 0, "head_dim <= 128")
        need(c.hidden_size % 64 == 0 and c.hidden_size <= 2048, "hidden_size % 64 == 0 and <= 2048")
        need(c.intermediate_size % 64 == 0 and c.projection_dim % 32 == 0, "intermediate % 64, projection % 32")
        self.head_dim_padded = 64 if self.head_dim <= 64 else 128
        self.config = c
        self._dtype, self._device = dtype, torch.device(device)
        self._w: Dict[str, torch.Tensor] = {}
        self._pos_cache = {}
        self._loaded = False
        self.patch_k = c.num_channels * c.patch_size * c.patch_size
        self.patch_k_padded = _pad(self.patch_k, 64)
        ops.ensure_workspace(self._device)

    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.float16, device="cuda", **kw):
        cfg = {}
        if os.path.exists(os.path.join(path, "config.json")):
            with open(os.path.join(path, "config.json")) as f:
                cfg = json.load(f)
        m = cls(cfg, dtype=torch_dtype, device=device, **kw)
        st, pt = os.path.join(path, "model.safetensors"), os.path.join(path, "pytorch_model.bin")
        sd = None
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        elif os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu")
        if sd is not None:   # a full CLIPModel checkpoint also carries the text tower: keep the vision side only
            keep = set(m.state_dict_shapes())
            m.load_state_dict({k: v for k, v in sd.items() if k in keep or k.endswith("position_ids")})
        return m

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a != self._dtype:
                raise NotImplementedError("pcdm_b200 CLIP vision: choose the dtype at construction (weights are pre-packed)")
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise RuntimeError("pcdm_b200 CLIP vision runs on CUDA only (no CPU fallback)")
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    # -- weights -----------------------------------------------------------------------------------------------------
    def state_dict_shapes(self) -> Dict[str, tuple]:
        c = self.config
        C, P, I = c.hidden_size, c.patch_size, c.intermediate_size
        n_pos = (c.image_size // P) ** 2 + 1
        v = "vision_model"
        sh = {f"{v}.embeddings.class_embedding": (C,),
              f"{v}.embeddings.patch_embedding.weight": (C, c.num_channels, P, P),
              f"{v}.embeddings.position_embedding.weight": (n_pos, C),
              f"{v}.pre_layrnorm.weight": (C,), f"{v}.pre_layrnorm.bias": (C,),
              f"{v}.post_layernorm.weight": (C,), f"{v}.post_layernorm.bias": (C,),
              "visual_projection.weight": (c.projection_dim, C)}
        for i in range(c.num_hidden_layers):
            p = f"{v}.encoder.layers.{i}"
            for n in ("layer_norm1", "layer_norm2"):
                sh[f"{p}.{n}.weight"] = (C,)
                sh[f"{p}.{n}.bias"] = (C,)
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                sh[f"{p}.self_attn.{n}.weight"] = (C, C)
                sh[f"{p}.self_attn.{n}.bias"] = (C,)
            sh[f"{p}.mlp.fc1.weight"], sh[f"{p}.mlp.fc1.bias"] = (I, C), (I,)
            sh[f"{p}.mlp.fc2.weight"], sh[f"{p}.mlp.fc2.bias"] = (C, I), (C,)
        return sh

    def load_state_dict(self, state_dict, strict: bool = True):
        shapes = self.state_dict_shapes()
        state_dict = {k: v for k, v in state_dict.items() if not k.endswith("position_ids")}   # old non-persistent buffer
        missing = [k for k in shapes if k not in state_dict]
        unexpected = [k for k in state_dict if k not in shapes]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for B200CLIPVisionModelWithProjection: missing "
                               f"{missing[:5]} unexpected {unexpected[:5]}")
        for k, shp in shapes.items():
            if k in state_dict and tuple(state_dict[k].shape) != shp:
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(state_dict[k].shape)} vs model {shp}")
        sd, w, dev, dt, c = state_dict, self._w, self._device, self._dtype, self.config
        C, heads, hd, hp = c.hidden_size, c.num_attention_heads, self.head_dim, self.head_dim_padded
        v = "vision_model"

        def f(k):
            return sd[k].detach().float()

        def mat(t):
            return t.to(device=dev, dtype=dt).contiguous()

        def vec(t):
            return t.to(device=dev, dtype=torch.float32).contiguous()

        def pad_heads_rows(t):   # [heads*hd, ...] -> [heads*hp, ...], zero rows after each head's hd real ones
            out = t.new_zeros((heads, hp) + tuple(t.shape[1:]))
            out[:, :hd] = t.reshape((heads, hd) + tuple(t.shape[1:]))
            return out.reshape((heads * hp,) + tuple(t.shape[1:]))

        pw = f(f"{v}.embeddings.patch_embedding.weight").reshape(C, self.patch_k)
        w["patch.weight"] = mat(torch.cat([pw, pw.new_zeros(C, self.patch_k_padded - self.patch_k)], dim=1))
        w["_pos_raw"] = vec(f(f"{v}.embeddings.position_embedding.weight"))     # [1 + n, C] fp32 (interpolated on demand)
        w["_cls_raw"] = vec(f(f"{v}.embeddings.class_embedding").reshape(1, C))
        self._pos_cache = {}
        self._arena = None
        for n in ("pre_layrnorm", "post_layernorm"):
            w[f"{n}.weight"], w[f"{n}.bias"] = vec(f(f"{v}.{n}.weight")), vec(f(f"{v}.{n}.bias"))
        w["proj.weight"] = mat(f("visual_projection.weight"))
        for i in range(c.num_hidden_layers):
            p = f"{v}.encoder.layers.{i}"
            for n in ("layer_norm1", "layer_norm2"):
                w[f"{i}.{n}.weight"], w[f"{i}.{n}.bias"] = vec(f(f"{p}.{n}.weight")), vec(f(f"{p}.{n}.bias"))
            a = f"{p}.self_attn"
            w[f"{i}.qkv.weight"] = mat(torch.cat([pad_heads_rows(f(f"{a}.{n}.weight")) for n in ("q_proj", "k_proj", "v_proj")]))
            w[f"{i}.qkv.bias"] = vec(torch.cat([pad_heads_rows(f(f"{a}.{n}.bias")) for n in ("q_proj", "k_proj", "v_proj")]))
            w[f"{i}.out.weight"] = mat(pad_heads_rows(f(f"{a}.out_proj.weight").t().contiguous()).t())
            w[f"{i}.out.bias"] = vec(f(f"{a}.out_proj.bias"))
            w[f"{i}.fc1.weight"], w[f"{i}.fc1.bias"] = mat(f(f"{p}.mlp.fc1.weight")), vec(f(f"{p}.mlp.fc1.bias"))
            w[f"{i}.fc2.weight"], w[f"{i}.fc2.bias"] = mat(f(f"{p}.mlp.fc2.weight")), vec(f(f"{p}.mlp.fc2.bias"))
        self._loaded = True
        self._weights_version += 1
        from types import SimpleNamespace
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def _after_adopt(self):
        self._pos_cache = {}

    def synthetic_state_dict(self, seed: int = 0, device=None) -> Dict[str, torch.Tensor]:
        dev = torch.device(device) if device is not None else self._device
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}
        for k, shp in self.state_dict_shapes().items():
            if ("norm" in k) and k.endswith(".weight"):
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g, device=dev)
            elif k.endswith("position_embedding.weight") or k.endswith("class_embedding"):
                sd[k] = 0.5 * torch.randn(shp, generator=g, device=dev)
            elif k.endswith(".weight"):
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
                sd[k] = torch.randn(shp, generator=g, device=dev) * fan_in ** -0.5
            else:
                sd[k] = 0.05 * torch.randn(shp, generator=g, device=dev)
        return sd

    def _positions(self, gh, gw, interpolate):
        """[1 + gh*gw, C] 16-bit: row 0 = class embedding + its position, rows 1.. = patch position embeddings
        (bicubically interpolated to the gh x gw grid under `interpolate_pos_encoding=True`, as transformers does)."""
        key = (gh, gw)
        if key not in self._pos_cache:
            pos = self._w["_pos_raw"].cpu()          # host-side, once per grid (cached below)
            n, C = pos.shape[0] - 1, pos.shape[1]
            patch = pos[1:]
            if not (gh * gw == n and gh == gw):
                if not interpolate:
                    s = self.config.image_size
                    raise ValueError(f"Input image size ({gh * self.config.patch_size}*{gw * self.config.patch_size}) "
                                     f"doesn't match model ({s}*{s}).")
                s = int(n ** 0.5)
                patch = torch.nn.functional.interpolate(patch.reshape(1, s, s, C).permute(0, 3, 1, 2), size=(gh, gw),
                                                        mode="bicubic", align_corners=False)
                patch = patch.permute(0, 2, 3, 1).reshape(gh * gw, C)
            full = torch.cat([self._w["_cls_raw"].cpu() + pos[:1], patch], dim=0)
            self._pos_cache[key] = full.to(device=self._device, dtype=self._dtype).contiguous()
        return self._pos_cache[key]

    # -- forward -------------------------------------------------------------------------------------------------------
    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def _guard(self, x):
        if not self._loaded:
            raise RuntimeError("B200CLIPVisionModelWithProjection: load_state_dict() first")
        if not x.is_cuda:
            raise RuntimeError("pcdm_b200 CLIP vision runs on CUDA tensors only (no CPU fallback)")

    @torch.no_grad()
    def forward(self, pixel_values, interpolate_pos_encoding: bool = False, **unused):
        self._guard(pixel_values)
        c, w, dt = self.config, self._w, self._dtype
        B, Cin, Hh, Ww = pixel_values.shape
        P, C, heads, hp = c.patch_size, c.hidden_size, c.num_attention_heads, self.head_dim_padded
        if Cin != c.num_channels or Hh % P or Ww % P:
            raise ValueError(f"pixel_values must be [B, {c.num_channels}, k*{P}, k*{P}], got {tuple(pixel_values.shape)}")
        gh, gw = Hh // P, Ww // P
        S = 1 + gh * gw
        pos = self._positions(gh, gw, interpolate_pos_encoding)
        # patch unfold (layout only): [B, Cin, gh, P, gw, P] -> [B*gh*gw, Cin*P*P], zero-padded to the GEMM's K
        cols = pixel_values.to(dt).reshape(B, Cin, gh, P, gw, P).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, -1)
        a = torch.zeros((B * gh * gw, self.patch_k_padded), device=pixel_values.device, dtype=dt)
        a[:, : self.patch_k] = cols
        e = torch.empty((B, S, C), device=pixel_values.device, dtype=dt)
        e[:, 0] = pos[0]
        for b in range(B):   # patch projection (no bias) + position embedding straight into rows 1.. of image b
            ops.gemm(a[b * gh * gw:(b + 1) * gh * gw], w["patch.weight"], out=e[b, 1:], residual=pos[1:])
        x = ops.layernorm(e.view(B * S, C), w["pre_layrnorm.weight"], w["pre_layrnorm.bias"], c.layer_norm_eps)
        Cp = heads * hp
        scale = float(self.head_dim) ** -0.5
        for i in range(c.num_hidden_layers):
            n = ops.layernorm(x, w[f"{i}.layer_norm1.weight"], w[f"{i}.layer_norm1.bias"], c.layer_norm_eps)
            qkv = ops.gemm(n, w[f"{i}.qkv.weight"], bias=w[f"{i}.qkv.bias"])
            att = ops.attention(qkv[:, :Cp], qkv[:, Cp:2 * Cp], qkv[:, 2 * Cp:], B, heads, scale=scale, head_dim=hp)
            x = ops.gemm(att, w[f"{i}.out.weight"], bias=w[f"{i}.out.bias"], residual=x)
            n = ops.layernorm(x, w[f"{i}.layer_norm2.weight"], w[f"{i}.layer_norm2.bias"], c.layer_norm_eps)
            h = ops.gemm(n, w[f"{i}.fc1.weight"], bias=w[f"{i}.fc1.bias"], gelu=True)
            x = ops.gemm(h, w[f"{i}.fc2.weight"], bias=w[f"{i}.fc2.bias"], residual=x)
        last = x.view(B, S, C)
        pooled = ops.layernorm(last[:, 0, :], w["post_layernorm.weight"], w["post_layernorm.bias"], c.layer_norm_eps)
        embeds = ops.gemm(pooled, w["proj.weight"])
        return CLIPVisionOutput(image_embeds=embeds, last_hidden_state=last, pooler_output=pooled)
