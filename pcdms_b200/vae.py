"""B200AutoencoderKL — drop-in for the diffusers `AutoencoderKL` the reference's pipelines hold as `pipe.vae`
(/root/reference/src/pipelines/stage2_inpaint_pipeline.py:133-137 registration, :443 `encode(...).latent_dist.sample`,
:528 `decode(latents / scaling_factor, return_dict=False)[0]`; stage3_refined_pipeline.py:479,563).  SURVEY.md §8f-1.

Same constructor config (stable-diffusion-2-1-base `vae/config.json`), same state-dict keys (both the current
`to_q/to_k/to_v/to_out.0` and the deprecated `query/key/value/proj_attn` attention names load), same `encode` /
`decode` / `.config.scaling_factor` surface.  Everything between the NCHW boundary tensors runs on the sm_100a kernels
of libpcdm_b200.so (NHWC 16-bit activations, fp32 accumulation): the implicit-GEMM conv3x3 (rows up to any multiple of
128 pixels; the encoder's bottom/right-padded stride-2 convs), GroupNorm(+SiLU), the 1x1 shortcut / quant GEMMs, and
the single-head dim-512 mid-block attention as GEMM (Q K^T, fp32 scores) -> row softmax -> GEMM (P V).  No PyTorch /
CPU compute fallback.
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

_DEFAULT_CONFIG = dict(
    in_channels=3, out_channels=3,
    down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
    block_out_channels=(128, 256, 512, 512), layers_per_block=2, act_fn="silu", latent_channels=4,
    norm_num_groups=32, sample_size=512, scaling_factor=0.18215, force_upcast=True, _diffusers_version="0.24.0",
)


class B200DiagonalGaussianDistribution:
    """`latent_dist` of `encode()`: holds the fp32 moments rows; `sample` / `mode` run pcdm_gaussian_sample."""

    def __init__(self, moments_rows, B, C, h, w, dtype):
        self._m, self._shape, self._dtype = moments_rows, (B, C, h, w), dtype

    def sample(self, generator=None):
        B, C, h, w = self._shape
        dev = self._m.device
        noise = torch.randn((B, C, h, w), generator=generator, dtype=torch.float32,
                            device=generator.device if generator is not None else dev).to(dev)
        return ops.gaussian_sample(self._m, B, C, h * w, noise=noise.contiguous()).view(B, C, h, w).to(self._dtype)

    def mode(self):
        B, C, h, w = self._shape
        return ops.gaussian_sample(self._m, B, C, h * w).view(B, C, h, w).to(self._dtype)

    @property
    def mean(self):
        return self.mode()


class B200AutoencoderKL(WeightArenaMixin):
    def __init__(self, dtype: torch.dtype = torch.float16, device="cuda", **config):
        cfg = dict(_DEFAULT_CONFIG)
        unknown = set(config) - set(cfg)
        if unknown:
            raise TypeError(f"unknown AutoencoderKL config keys: {sorted(unknown)}")
        cfg.update(config)
        ch = tuple(cfg["block_out_channels"])

        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"pcdm_b200 VAE: unsupported config ({what})")
        need(all(t == "DownEncoderBlock2D" for t in cfg["down_block_types"]), "down_block_types")
        need(all(t == "UpDecoderBlock2D" for t in cfg["up_block_types"]), "up_block_types")
        need(len(cfg["down_block_types"]) == len(ch) == len(cfg["up_block_types"]), "block counts")
        need(all(c % 64 == 0 for c in ch), "block_out_channels % 64")
        need(cfg["act_fn"] in ("silu", "swish"), "act_fn")
        need(cfg["in_channels"] <= 64 and cfg["out_channels"] <= 32 and 2 * cfg["latent_channels"] <= 32, "channels")
        cfg["block_out_channels"] = ch
        self.config = _Config(cfg)
        self._dtype, self._device = dtype, torch.device(device)
        self._w: Dict[str, torch.Tensor] = {}
        self._loaded = False
        ops.ensure_workspace(self._device)

    # -- diffusers-style surface ----------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=torch.float16, device="cuda",
                        **overrides):
        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        cfg = {}
        if os.path.exists(os.path.join(root, "config.json")):
            with open(os.path.join(root, "config.json")) as f:
                cfg = {k: v for k, v in json.load(f).items() if k in _DEFAULT_CONFIG}
        cfg.update({k: v for k, v in overrides.items() if k in _DEFAULT_CONFIG})
        model = cls(dtype=torch_dtype, device=device, **cfg)
        st, pt = (os.path.join(root, f"diffusion_pytorch_model.{e}") for e in ("safetensors", "bin"))
        if os.path.exists(st):
            from safetensors.torch import load_file
            model.load_state_dict(load_file(st))
        elif os.path.exists(pt):
            model.load_state_dict(torch.load(pt, map_location="cpu"))
        return model

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a != self._dtype:
                raise NotImplementedError("pcdm_b200 VAE: choose the dtype at construction (weights are pre-packed)")
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise RuntimeError("pcdm_b200 VAE runs on CUDA only (no CPU fallback)")
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    def enable_slicing(self):
        return None

    def enable_tiling(self):
        return None

    def parameters(self):
        return iter(self._w.values())

    # -- topology ---------------------------------------------------------------------------------------------------
    def _resnet_list(self):
        """(prefix, cin, cout) of every ResnetBlock2D, plus samplers, in diffusers key order."""
        cfg = self.config
        ch, L = list(cfg.block_out_channels), cfg.layers_per_block
        enc, dec = [], []
        out_c = ch[0]
        for i, c in enumerate(ch):
            in_c, out_c = out_c, c
            for j in range(L):
                enc.append(("res", f"encoder.down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c))
            if i < len(ch) - 1:
                enc.append(("down", f"encoder.down_blocks.{i}.downsamplers.0.conv", out_c, out_c))
        rev = ch[::-1]
        out_c = rev[0]
        for i, c in enumerate(rev):
            prev, out_c = out_c, c
            for j in range(L + 1):
                dec.append(("res", f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else out_c, out_c))
            if i < len(ch) - 1:
                dec.append(("up", f"decoder.up_blocks.{i}.upsamplers.0.conv", out_c, out_c))
        return enc, dec

    def state_dict_shapes(self) -> Dict[str, tuple]:
        cfg = self.config
        ch, lc = cfg.block_out_channels, cfg.latent_channels
        sh: Dict[str, tuple] = {}

        def wb(name, wshape):
            sh[f"{name}.weight"] = tuple(wshape)
            sh[f"{name}.bias"] = (wshape[0],)

        def res(p, cin, cout):
            wb(f"{p}.norm1", (cin,))
            wb(f"{p}.conv1", (cout, cin, 3, 3))
            wb(f"{p}.norm2", (cout,))
            wb(f"{p}.conv2", (cout, cout, 3, 3))
            if cin != cout:
                wb(f"{p}.conv_shortcut", (cout, cin, 1, 1))

        def mid(p, c):
            res(f"{p}.resnets.0", c, c)
            res(f"{p}.resnets.1", c, c)
            wb(f"{p}.attentions.0.group_norm", (c,))
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                wb(f"{p}.attentions.0.{n}", (c, c))

        enc, dec = self._resnet_list()
        wb("encoder.conv_in", (ch[0], cfg.in_channels, 3, 3))
        for kind, p, cin, cout in enc:
            res(p, cin, cout) if kind == "res" else wb(p, (cout, cin, 3, 3))
        mid("encoder.mid_block", ch[-1])
        wb("encoder.conv_norm_out", (ch[-1],))
        wb("encoder.conv_out", (2 * lc, ch[-1], 3, 3))
        wb("quant_conv", (2 * lc, 2 * lc, 1, 1))
        wb("post_quant_conv", (lc, lc, 1, 1))
        wb("decoder.conv_in", (ch[-1], lc, 3, 3))
        mid("decoder.mid_block", ch[-1])
        for kind, p, cin, cout in dec:
            res(p, cin, cout) if kind == "res" else wb(p, (cout, cin, 3, 3))
        wb("decoder.conv_norm_out", (ch[0],))
        wb("decoder.conv_out", (cfg.out_channels, ch[0], 3, 3))
        return sh

    def synthetic_state_dict(self, seed: int = 0, device=None) -> Dict[str, torch.Tensor]:
        """Random weights of the real shapes (fan-in scaled), for bench / smoke runs without a checkpoint."""
        dev = torch.device(device) if device is not None else self._device
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}
        for k, shp in self.state_dict_shapes().items():
            if k.endswith(".weight") and "norm" not in k:
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
                sd[k] = torch.randn(shp, generator=g, device=dev) * (fan_in ** -0.5)
            elif k.endswith(".weight"):
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g, device=dev)
            else:
                sd[k] = 0.05 * torch.randn(shp, generator=g, device=dev)
        return sd

    _DEPRECATED_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}

    def load_state_dict(self, state_dict, strict: bool = True):
        dev, dt = self._device, self._dtype
        sd = {}
        for k, v in state_dict.items():   # deprecated AttentionBlock names (diffusers converts them on load too)
            parts = k.split(".")
            if len(parts) >= 2 and parts[-2] in self._DEPRECATED_ATTN and ".attentions." in k:
                k = ".".join(parts[:-2] + [self._DEPRECATED_ATTN[parts[-2]], parts[-1]])
            sd[k] = v
        shapes = self.state_dict_shapes()
        missing = [k for k in shapes if k not in sd]
        unexpected = [k for k in sd if k not in shapes]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for B200AutoencoderKL: missing {missing[:5]} "
                               f"unexpected {unexpected[:5]}")
        for k, shp in shapes.items():
            if k in sd:
                got = tuple(sd[k].shape)
                if got != shp and not (".attentions." in k and len(got) == 4 and got[:2] == shp):
                    raise RuntimeError(f"size mismatch for {k}: checkpoint {got} vs model {shp}")
        w = self._w

        def f32(k, pad=None):
            t = sd[k].detach().float().reshape(-1)
            if pad is not None and t.numel() < pad:
                t = torch.cat([t, t.new_zeros(pad - t.numel())])
            return t.to(dev).contiguous()

        def conv(k, pad_in=None, pad_out=None):
            t = sd[k].detach().float()
            if pad_in is not None and t.shape[1] < pad_in:
                t = torch.cat([t, t.new_zeros(t.shape[0], pad_in - t.shape[1], 3, 3)], dim=1)
            if pad_out is not None and t.shape[0] < pad_out:
                t = torch.cat([t, t.new_zeros(pad_out - t.shape[0], *t.shape[1:])], dim=0)
            return ops.pack_conv3x3_weight(t, dt).to(dev)

        def lin(k, n_pad=None, k_pad=None):
            t = sd[k].detach().float()
            t = t.reshape(t.shape[0], t.shape[1])
            if k_pad is not None and t.shape[1] < k_pad:
                t = torch.cat([t, t.new_zeros(t.shape[0], k_pad - t.shape[1])], dim=1)
            if n_pad is not None and t.shape[0] < n_pad:
                t = torch.cat([t, t.new_zeros(n_pad - t.shape[0], t.shape[1])], dim=0)
            return t.to(device=dev, dtype=dt).contiguous()

        def res(p, cin, cout):
            for n in ("norm1", "norm2"):
                w[f"{p}.{n}.weight"], w[f"{p}.{n}.bias"] = f32(f"{p}.{n}.weight"), f32(f"{p}.{n}.bias")
            for n in ("conv1", "conv2"):
                w[f"{p}.{n}.weight"], w[f"{p}.{n}.bias"] = conv(f"{p}.{n}.weight"), f32(f"{p}.{n}.bias")
            if cin != cout:
                w[f"{p}.conv_shortcut.weight"] = lin(f"{p}.conv_shortcut.weight")
                w[f"{p}.conv_shortcut.bias"] = f32(f"{p}.conv_shortcut.bias")

        def mid(p, c):
            res(f"{p}.resnets.0", c, c)
            res(f"{p}.resnets.1", c, c)
            a = f"{p}.attentions.0"
            w[f"{a}.group_norm.weight"], w[f"{a}.group_norm.bias"] = f32(f"{a}.group_norm.weight"), f32(f"{a}.group_norm.bias")
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                w[f"{a}.{n}.weight"], w[f"{a}.{n}.bias"] = lin(f"{a}.{n}.weight"), f32(f"{a}.{n}.bias")

        cfg = self.config
        enc, dec = self._resnet_list()
        w["encoder.conv_in.weight"], w["encoder.conv_in.bias"] = conv("encoder.conv_in.weight", pad_in=64), f32("encoder.conv_in.bias")
        for kind, p, cin, cout in enc + dec:
            if kind == "res":
                res(p, cin, cout)
            else:
                w[f"{p}.weight"], w[f"{p}.bias"] = conv(f"{p}.weight"), f32(f"{p}.bias")
        mid("encoder.mid_block", cfg.block_out_channels[-1])
        mid("decoder.mid_block", cfg.block_out_channels[-1])
        for n in ("encoder.conv_norm_out", "decoder.conv_norm_out"):
            w[f"{n}.weight"], w[f"{n}.bias"] = f32(f"{n}.weight"), f32(f"{n}.bias")
        # encoder tail: conv_out (2*latent channels, padded to 64) -> quant_conv as a [32 x 64] GEMM with fp32 output
        w["encoder.conv_out.weight"], w["encoder.conv_out.bias"] = conv("encoder.conv_out.weight", pad_out=64), f32("encoder.conv_out.bias", pad=64)
        w["quant_conv.weight"], w["quant_conv.bias"] = lin("quant_conv.weight", n_pad=32, k_pad=64), f32("quant_conv.bias", pad=32)
        # decoder head: post_quant_conv as a [64 x 64] GEMM on the channel-padded latents -> conv_in (Cin padded to 64)
        w["post_quant_conv.weight"], w["post_quant_conv.bias"] = lin("post_quant_conv.weight", n_pad=64, k_pad=64), f32("post_quant_conv.bias", pad=64)
        w["decoder.conv_in.weight"], w["decoder.conv_in.bias"] = conv("decoder.conv_in.weight", pad_in=64), f32("decoder.conv_in.bias")
        w["decoder.conv_out.weight"], w["decoder.conv_out.bias"] = conv("decoder.conv_out.weight", pad_out=32), f32("decoder.conv_out.bias", pad=32)
        self._arena = None
        self._loaded = True
        self._weights_version += 1
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def weight_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._w.values())

    # -- blocks -------------------------------------------------------------------------------------------------------
    def _resnet(self, p, x, cin, cout):
        w = self._w
        B, H, W, _ = x.shape
        h = ops.groupnorm(x, w[f"{p}.norm1.weight"], w[f"{p}.norm1.bias"], 1e-6, silu=True)
        h = ops.conv3x3(h, w[f"{p}.conv1.weight"], bias=w[f"{p}.conv1.bias"])
        h = ops.groupnorm(h, w[f"{p}.norm2.weight"], w[f"{p}.norm2.bias"], 1e-6, silu=True)
        if cin != cout:
            x = ops.gemm(x.view(B * H * W, cin), w[f"{p}.conv_shortcut.weight"],
                         bias=w[f"{p}.conv_shortcut.bias"]).view(B, H, W, cout)
        return ops.conv3x3(h, w[f"{p}.conv2.weight"], bias=w[f"{p}.conv2.bias"], residual=x)

    def _attention(self, p, x):
        """Single-head attention over the H*W tokens of each image (head dim = C): Q K^T and P V on the GEMM kernel
        (K^T and V^T are never formed by a transpose: K rows are the [N, K] operand of the score GEMM as they are, and
        V^T = W_v X^T comes straight out of a GEMM with the roles of weight and activation swapped; V's bias is added
        after P V, exact because every row of P sums to one)."""
        w, dt = self._w, self._dtype
        B, H, W, C = x.shape
        S = H * W
        if S % 64:
            raise NotImplementedError("pcdm_b200 VAE attention: H*W of the latent must be a multiple of 64")
        hn = ops.groupnorm(x, w[f"{p}.group_norm.weight"], w[f"{p}.group_norm.bias"], 1e-6).view(B * S, C)
        q = ops.gemm(hn, w[f"{p}.to_q.weight"], bias=w[f"{p}.to_q.bias"])
        k = ops.gemm(hn, w[f"{p}.to_k.weight"], bias=w[f"{p}.to_k.bias"])
        o = torch.empty((B * S, C), device=x.device, dtype=dt)
        scores = torch.empty((S, S), device=x.device, dtype=torch.float32)
        probs = torch.empty((S, S), device=x.device, dtype=dt)
        vt = torch.empty((C, S), device=x.device, dtype=dt)
        for b in range(B):
            rows = slice(b * S, (b + 1) * S)
            ops.gemm(q[rows], k[rows], out=scores, out_f32=True, w_static=False)
            ops.softmax_rows(scores, C ** -0.5, dt, out=probs)
            ops.gemm(w[f"{p}.to_v.weight"], hn[rows], out=vt, w_static=False)
            ops.gemm(probs, vt, out=o[rows], bias=w[f"{p}.to_v.bias"], w_static=False)
        out = ops.gemm(o, w[f"{p}.to_out.0.weight"], bias=w[f"{p}.to_out.0.bias"], residual=x.view(B * S, C))
        return out.view(B, H, W, C)

    def _mid(self, p, x):
        c = x.shape[-1]
        x = self._resnet(f"{p}.resnets.0", x, c, c)
        x = self._attention(f"{p}.attentions.0", x)
        return self._resnet(f"{p}.resnets.1", x, c, c)

    def _check_input(self, x, channels, what):
        if not self._loaded:
            raise RuntimeError("B200AutoencoderKL: load_state_dict() first")
        if not x.is_cuda:
            raise RuntimeError("pcdm_b200 VAE runs on CUDA tensors only (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != channels:
            raise ValueError(f"{what}: expected [B, {channels}, H, W], got {tuple(x.shape)}")

    # -- encode / decode ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x, return_dict: bool = True):
        cfg, w, dt = self.config, self._w, self._dtype
        self._check_input(x, cfg.in_channels, "encode")
        h = ops.nchw_to_nhwc_pad(x.contiguous(), 64, dt)
        h = ops.conv3x3(h, w["encoder.conv_in.weight"], bias=w["encoder.conv_in.bias"])
        for kind, p, cin, cout in self._resnet_list()[0]:
            if kind == "res":
                h = self._resnet(p, h, cin, cout)
            else:
                h = ops.conv3x3(h, w[f"{p}.weight"], bias=w[f"{p}.bias"], stride=2, pad_br=True)
        h = self._mid("encoder.mid_block", h)
        h = ops.groupnorm(h, w["encoder.conv_norm_out.weight"], w["encoder.conv_norm_out.bias"], 1e-6, silu=True)
        m = ops.conv3x3(h, w["encoder.conv_out.weight"], bias=w["encoder.conv_out.bias"])
        B, hh, ww, _ = m.shape
        moments = ops.gemm(m.view(B * hh * ww, 64), w["quant_conv.weight"], bias=w["quant_conv.bias"], out_f32=True)
        dist = B200DiagonalGaussianDistribution(moments, B, cfg.latent_channels, hh, ww, dt)
        if not return_dict:
            return (dist,)
        return SimpleNamespace(latent_dist=dist)

    @torch.no_grad()
    def decode(self, z, return_dict: bool = True, generator=None):
        cfg, w, dt = self.config, self._w, self._dtype
        self._check_input(z, cfg.latent_channels, "decode")
        B, _, hh, ww = z.shape
        zi = ops.nchw_to_nhwc_pad(z.contiguous(), 64, dt)
        zq = ops.gemm(zi.view(B * hh * ww, 64), w["post_quant_conv.weight"], bias=w["post_quant_conv.bias"])
        h = ops.conv3x3(zq.view(B, hh, ww, 64), w["decoder.conv_in.weight"], bias=w["decoder.conv_in.bias"])
        h = self._mid("decoder.mid_block", h)
        for kind, p, cin, cout in self._resnet_list()[1]:
            if kind == "res":
                h = self._resnet(p, h, cin, cout)
            else:
                h = ops.conv3x3(ops.upsample_nearest2x(h), w[f"{p}.weight"], bias=w[f"{p}.bias"])
        h = ops.groupnorm(h, w["decoder.conv_norm_out.weight"], w["decoder.conv_norm_out.bias"], 1e-6, silu=True)
        rows = ops.conv3x3(h, w["decoder.conv_out.weight"], bias=w["decoder.conv_out.bias"], out_f32=True)
        img = ops.nhwc_to_nchw(rows, cfg.out_channels, z.dtype if z.dtype.is_floating_point else dt)
        if not return_dict:
            return (img,)
        return SimpleNamespace(sample=img)

    def __call__(self, sample, sample_posterior: bool = False, return_dict: bool = True, generator=None):
        dist = self.encode(sample).latent_dist
        z = dist.sample(generator=generator) if sample_posterior else dist.mode()
        return self.decode(z, return_dict=return_dict)
