"""B200Dinov2Model — drop-in for the `transformers.Dinov2Model` the reference's drivers use as `image_encoder_p`
(/root/reference/stage2_batchtest_inpaint_model.py:98 `Dinov2Model.from_pretrained(...)`, applied at :168
`image_encoder_p(pixel_values).last_hidden_state`; pcdms_demo.ipynb; stage3_batchtest_refined_model.py) — the
DINOv2-giant ViT (40 layers, width 1536, 24 heads of 64, SwiGLU FFN, LayerScale) that turns the source image into the
257 conditioning tokens of the stage-2 UNet.  SURVEY.md §8f-3.

Same config keys, same state-dict key names (`embeddings.*`, `encoder.layer.N.*`, `layernorm.*`), same call surface
(`model(pixel_values).last_hidden_state / .pooler_output`).  Everything after the patch unfold runs on the sm_100a
kernels of libpcdm_b200.so: patch projection, fused-QKV, attention output, SwiGLU (gate activation in the GEMM
epilogue) and FFN output as tcgen05 GEMMs with bias / residual epilogues, the d = 64 flash-attention kernel over the
257 tokens, LayerNorm.  Weight-only transformations are done once at load: q/k/v concatenation, LayerScale folded
into the projection that precedes it (lambda * (W x + b) = (lambda W) x + lambda b), SwiGLU rows interleaved for the
gated epilogue and zero-padded to a multiple of 64, bicubic interpolation of the position embeddings to the input grid
(as transformers' `interpolate_pos_encoding`), cls token + its position.  No PyTorch / CPU compute fallback.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict

import torch

from . import ops
from .arena import WeightArenaMixin
from .unet import _Config

_DEFAULT_CONFIG = dict(hidden_size=1536, num_hidden_layers=40, num_attention_heads=24, mlp_ratio=4, hidden_act="gelu",
                       layer_norm_eps=1e-6, image_size=518, patch_size=14, num_channels=3, qkv_bias=True,
                       layerscale_value=1.0, use_swiglu_ffn=True, hidden_dropout_prob=0.0,
                       attention_probs_dropout_prob=0.0, drop_path_rate=0.0, use_mask_token=True)


def _pad(n, m):
    return (n + m - 1) // m * m


class B200Dinov2Model(WeightArenaMixin):
    def __init__(self, config=None, dtype: torch.dtype = torch.float16, device="cuda", **kw):
        cfg = dict(_DEFAULT_CONFIG)
        src = dict(config.to_dict() if hasattr(config, "to_dict") else (config or {}))
        src.update(kw)
        cfg.update({k: v for k, v in src.items() if k in cfg})
        c = _Config(cfg)

        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"pcdm_b200 DINOv2: unsupported config ({what})")
        need(c.use_swiglu_ffn, "only the SwiGLU FFN of dinov2-giant is implemented")
        need(c.hidden_size % 64 == 0 and c.hidden_size == 64 * c.num_attention_heads, "head_dim must be 64")
        need(c.hidden_size <= 2048, "hidden_size <= 2048 (LayerNorm kernel)")
        need(c.qkv_bias, "qkv_bias")
        self.config = c
        self._dtype, self._device = dtype, torch.device(device)
        self._w: Dict[str, torch.Tensor] = {}
        self._pos_cache = {}
        self._loaded = False
        hf = int(c.hidden_size * c.mlp_ratio)
        self.ffn_hidden = (int(hf * 2 / 3) + 7) // 8 * 8
        self.ffn_hidden_padded = _pad(self.ffn_hidden, 64)
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
        if os.path.exists(st):
            from safetensors.torch import load_file
            m.load_state_dict(load_file(st))
        elif os.path.exists(pt):
            m.load_state_dict(torch.load(pt, map_location="cpu"))
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
                raise NotImplementedError("pcdm_b200 DINOv2: choose the dtype at construction (weights are pre-packed)")
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise RuntimeError("pcdm_b200 DINOv2 runs on CUDA only (no CPU fallback)")
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    # -- weights -----------------------------------------------------------------------------------------------------
    def state_dict_shapes(self) -> Dict[str, tuple]:
        c = self.config
        C, P = c.hidden_size, c.patch_size
        n_pos = (c.image_size // P) ** 2 + 1
        sh = {"embeddings.cls_token": (1, 1, C), "embeddings.position_embeddings": (1, n_pos, C),
              "embeddings.patch_embeddings.projection.weight": (C, c.num_channels, P, P),
              "embeddings.patch_embeddings.projection.bias": (C,), "layernorm.weight": (C,), "layernorm.bias": (C,)}
        if c.use_mask_token:
            sh["embeddings.mask_token"] = (1, C)
        for i in range(c.num_hidden_layers):
            p = f"encoder.layer.{i}"
            for n in ("norm1", "norm2"):
                sh[f"{p}.{n}.weight"] = (C,)
                sh[f"{p}.{n}.bias"] = (C,)
            for n in ("query", "key", "value"):
                sh[f"{p}.attention.attention.{n}.weight"] = (C, C)
                sh[f"{p}.attention.attention.{n}.bias"] = (C,)
            sh[f"{p}.attention.output.dense.weight"] = (C, C)
            sh[f"{p}.attention.output.dense.bias"] = (C,)
            sh[f"{p}.layer_scale1.lambda1"] = (C,)
            sh[f"{p}.layer_scale2.lambda1"] = (C,)
            sh[f"{p}.mlp.weights_in.weight"] = (2 * self.ffn_hidden, C)
            sh[f"{p}.mlp.weights_in.bias"] = (2 * self.ffn_hidden,)
            sh[f"{p}.mlp.weights_out.weight"] = (C, self.ffn_hidden)
            sh[f"{p}.mlp.weights_out.bias"] = (C,)
        return sh

    def load_state_dict(self, state_dict, strict: bool = True):
        shapes = self.state_dict_shapes()
        missing = [k for k in shapes if k not in state_dict and k != "embeddings.mask_token"]
        unexpected = [k for k in state_dict if k not in shapes]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for B200Dinov2Model: missing {missing[:5]} "
                               f"unexpected {unexpected[:5]}")
        for k, shp in shapes.items():
            if k in state_dict and tuple(state_dict[k].shape) != shp:
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(state_dict[k].shape)} vs model {shp}")
        sd, w, dev, dt, c = state_dict, self._w, self._device, self._dtype, self.config
        C, H, Hp = c.hidden_size, self.ffn_hidden, self.ffn_hidden_padded

        def f(k):
            return sd[k].detach().float()

        def mat(t):
            return t.to(device=dev, dtype=dt).contiguous()

        def vec(t):
            return t.to(device=dev, dtype=torch.float32).contiguous()

        pw = f("embeddings.patch_embeddings.projection.weight").reshape(C, self.patch_k)
        w["patch.weight"] = mat(torch.cat([pw, pw.new_zeros(C, self.patch_k_padded - self.patch_k)], dim=1))
        w["patch.bias"] = vec(f("embeddings.patch_embeddings.projection.bias"))
        w["_pos_raw"] = vec(f("embeddings.position_embeddings"))               # [1, 1 + n, C] fp32 (interpolated on demand)
        w["_cls_raw"] = vec(f("embeddings.cls_token").reshape(1, C))
        self._pos_cache = {}
        self._arena = None
        w["layernorm.weight"], w["layernorm.bias"] = vec(f("layernorm.weight")), vec(f("layernorm.bias"))
        # SwiGLU: hidden = silu(x1) * x2 with (x1 | x2) = chunk(weights_in(x)): x2 is the value, x1 the gate.  Rows are
        # interleaved in groups of [32 value | 32 gate] for the gated GEMM epilogue; the hidden width is zero-padded.
        idx = torch.arange(Hp).view(-1, 32)
        perm = torch.cat([idx + Hp, idx], dim=1).reshape(-1)
        for i in range(c.num_hidden_layers):
            p = f"encoder.layer.{i}"
            for n in ("norm1", "norm2"):
                w[f"{p}.{n}.weight"], w[f"{p}.{n}.bias"] = vec(f(f"{p}.{n}.weight")), vec(f(f"{p}.{n}.bias"))
            a = f"{p}.attention.attention"
            w[f"{p}.qkv.weight"] = mat(torch.cat([f(f"{a}.query.weight"), f(f"{a}.key.weight"), f(f"{a}.value.weight")]))
            w[f"{p}.qkv.bias"] = vec(torch.cat([f(f"{a}.query.bias"), f(f"{a}.key.bias"), f(f"{a}.value.bias")]))
            l1, l2 = f(f"{p}.layer_scale1.lambda1"), f(f"{p}.layer_scale2.lambda1")
            w[f"{p}.out.weight"] = mat(l1[:, None] * f(f"{p}.attention.output.dense.weight"))
            w[f"{p}.out.bias"] = vec(l1 * f(f"{p}.attention.output.dense.bias"))
            wi, bi = f(f"{p}.mlp.weights_in.weight"), f(f"{p}.mlp.weights_in.bias")
            wi_p = wi.new_zeros(2 * Hp, C)
            bi_p = bi.new_zeros(2 * Hp)
            wi_p[:H], wi_p[Hp:Hp + H] = wi[:H], wi[H:]
            bi_p[:H], bi_p[Hp:Hp + H] = bi[:H], bi[H:]
            w[f"{p}.ffn_in.weight"], w[f"{p}.ffn_in.bias"] = mat(wi_p[perm]), vec(bi_p[perm])
            wo = f(f"{p}.mlp.weights_out.weight")
            wo_p = torch.cat([wo, wo.new_zeros(C, Hp - H)], dim=1)
            w[f"{p}.ffn_out.weight"] = mat(l2[:, None] * wo_p)
            w[f"{p}.ffn_out.bias"] = vec(l2 * f(f"{p}.mlp.weights_out.bias"))
        self._loaded = True
        self._weights_version += 1
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def _after_adopt(self):
        self._pos_cache = {}

    def synthetic_state_dict(self, seed: int = 0, device=None) -> Dict[str, torch.Tensor]:
        dev = torch.device(device) if device is not None else self._device
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}
        for k, shp in self.state_dict_shapes().items():
            if k.endswith("lambda1"):
                sd[k] = 0.5 + 0.1 * torch.randn(shp, generator=g, device=dev)
            elif "norm" in k and k.endswith(".weight"):
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g, device=dev)
            elif k.endswith(".weight"):
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
                sd[k] = torch.randn(shp, generator=g, device=dev) * fan_in ** -0.5
            elif k.endswith(".bias"):
                sd[k] = 0.05 * torch.randn(shp, generator=g, device=dev)
            else:
                sd[k] = 0.5 * torch.randn(shp, generator=g, device=dev)
        return sd

    def _positions(self, gh, gw):
        """[1 + gh*gw, C] 16-bit: row 0 = cls token + its position, rows 1.. = patch position embeddings, bicubically
        interpolated to the gh x gw grid when it differs from the trained one (transformers' interpolate_pos_encoding)."""
        key = (gh, gw)
        if key not in self._pos_cache:
            pos = self._w["_pos_raw"].cpu()                       # [1, 1 + n, C] fp32, host-side once per grid
            n = pos.shape[1] - 1
            C = pos.shape[-1]
            patch = pos[:, 1:]
            if not (gh * gw == n and gh == gw):
                s = int(n ** 0.5)
                patch = torch.nn.functional.interpolate(patch.reshape(1, s, s, C).permute(0, 3, 1, 2), size=(gh, gw),
                                                        mode="bicubic", align_corners=False)
                patch = patch.permute(0, 2, 3, 1).reshape(1, gh * gw, C)
            full = torch.cat([self._w["_cls_raw"].cpu() + pos[0, :1], patch[0]], dim=0)
            self._pos_cache[key] = full.to(device=self._device, dtype=self._dtype).contiguous()
        return self._pos_cache[key]

    # -- forward -------------------------------------------------------------------------------------------------------
    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def _guard(self, x):
        if not self._loaded:
            raise RuntimeError("B200Dinov2Model: load_state_dict() first")
        if not x.is_cuda:
            raise RuntimeError("pcdm_b200 DINOv2 runs on CUDA tensors only (no CPU fallback)")

    @torch.no_grad()
    def forward(self, pixel_values, bool_masked_pos=None, **unused):
        if bool_masked_pos is not None:
            raise NotImplementedError("bool_masked_pos (pre-training only)")
        self._guard(pixel_values)
        c, w, dt = self.config, self._w, self._dtype
        B, Cin, Hh, Ww = pixel_values.shape
        P, C, heads = c.patch_size, c.hidden_size, c.num_attention_heads
        if Cin != c.num_channels or Hh % P or Ww % P:
            raise ValueError(f"pixel_values must be [B, {c.num_channels}, k*{P}, k*{P}], got {tuple(pixel_values.shape)}")
        gh, gw = Hh // P, Ww // P
        S = 1 + gh * gw
        # patch unfold (layout only): [B, Cin, gh, P, gw, P] -> [B*gh*gw, Cin*P*P], zero-padded to the GEMM's K
        cols = pixel_values.to(dt).reshape(B, Cin, gh, P, gw, P).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, -1)
        a = torch.zeros((B * gh * gw, self.patch_k_padded), device=pixel_values.device, dtype=dt)
        a[:, : self.patch_k] = cols
        pos = self._positions(gh, gw)
        x = torch.empty((B, S, C), device=pixel_values.device, dtype=dt)
        x[:, 0] = pos[0]
        for b in range(B):   # patch projection + bias + position embedding straight into rows 1.. of image b
            ops.gemm(a[b * gh * gw:(b + 1) * gh * gw], w["patch.weight"], out=x[b, 1:], bias=w["patch.bias"],
                     residual=pos[1:])
        x = x.view(B * S, C)
        for i in range(c.num_hidden_layers):
            p = f"encoder.layer.{i}"
            n = ops.layernorm(x, w[f"{p}.norm1.weight"], w[f"{p}.norm1.bias"], c.layer_norm_eps)
            qkv = ops.gemm(n, w[f"{p}.qkv.weight"], bias=w[f"{p}.qkv.bias"])
            att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
            x = ops.gemm(att, w[f"{p}.out.weight"], bias=w[f"{p}.out.bias"], residual=x)        # + LayerScale folded
            n = ops.layernorm(x, w[f"{p}.norm2.weight"], w[f"{p}.norm2.bias"], c.layer_norm_eps)
            h = ops.gemm(n, w[f"{p}.ffn_in.weight"], bias=w[f"{p}.ffn_in.bias"], geglu=True, silu=True)   # SwiGLU
            x = ops.gemm(h, w[f"{p}.ffn_out.weight"], bias=w[f"{p}.ffn_out.bias"], residual=x)
        y = ops.layernorm(x, w["layernorm.weight"], w["layernorm.bias"], c.layer_norm_eps).view(B, S, C)
        return SimpleNamespace(last_hidden_state=y, pooler_output=y[:, 0, :])
