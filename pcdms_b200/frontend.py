"""Conditioning front-end of the stage-2 driver on the B200 kernels: the two small modules the
reference runs once per image pair before the denoising loop.

* `B200ImageProjModel_p` — the reference's `ImageProjModel_p` (/root/reference/stage2_batchtest_inpaint_model.py:48-66,
  applied at :169; also stage3_batchtest_refined_model.py:51): Linear -> GELU -> LayerNorm -> Linear over the 257 DINOv2
  tokens.  Two GEMM launches (GELU in the first epilogue) and one LayerNorm launch.
* `B200ControlNetConditioningEmbedding` — diffusers' module the driver instantiates as `pose_proj`
  (stage2_batchtest_inpaint_model.py:101, applied at :179): 8 conv3x3 (+SiLU) on the pose canvas, three of them
  stride 2.  Channel counts 16 / 32 / 96 are zero-padded to 64 / 64 / 128 at load time (exact: padded weights and
  biases are zero and SiLU(0) = 0), so every layer is the tcgen05 implicit-GEMM conv with SiLU fused in its epilogue.

Both keep the reference's state-dict keys (`net.0/3/4.*`; `conv_in, blocks.0-5, conv_out`), take/return the tensors the
driver passes ([1, 257, 1536] -> [1, 257, 1024]; NCHW [1, 3, H, 2W] -> NCHW [1, 320, H/8, W/4]) and have no CPU
fallback.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .arena import WeightArenaMixin


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class _Module(WeightArenaMixin):
    def __init__(self, dtype, device):
        self._dtype, self._device = dtype, torch.device(device)
        self._w: Dict[str, torch.Tensor] = {}
        self._loaded = False

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def synthetic_state_dict(self, seed: int = 0, device=None) -> Dict[str, torch.Tensor]:
        """Seeded random weights with the reference's key names and shapes (benches, tests)."""
        g = torch.Generator().manual_seed(seed)
        sd = {}
        for k, shp in self.state_dict_shapes().items():
            if len(shp) >= 2:
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
                sd[k] = torch.randn(shp, generator=g) * fan_in ** -0.5
            elif k.endswith(".weight"):      # LayerNorm scale
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
            else:
                sd[k] = 0.05 * torch.randn(shp, generator=g)
        return sd

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a != self._dtype:
                raise NotImplementedError("pcdm_b200: choose the dtype at construction (weights are pre-packed)")
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise RuntimeError("pcdm_b200 modules run on CUDA only (no CPU fallback)")
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def _check(self, x):
        if not self._loaded:
            raise RuntimeError(f"{type(self).__name__}: load_state_dict() first")
        if not x.is_cuda:
            raise RuntimeError("pcdm_b200 modules run on CUDA tensors only (no CPU fallback)")

    def _load_check(self, state_dict, shapes, strict):
        missing = [k for k in shapes if k not in state_dict]
        unexpected = [k for k in state_dict if k not in shapes]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for {type(self).__name__}: missing {missing[:5]} "
                               f"unexpected {unexpected[:5]}")
        for k, shp in shapes.items():
            if k in state_dict and tuple(state_dict[k].shape) != shp:
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(state_dict[k].shape)} vs model {shp}")


class B200ImageProjModel_p(_Module):
    def __init__(self, in_dim=1536, hidden_dim=768, out_dim=1024, dropout=0.0, dtype=torch.float16, device="cuda"):
        super().__init__(dtype, device)
        if in_dim % 64 or hidden_dim % 64 or out_dim % 32:
            raise NotImplementedError("B200ImageProjModel_p: in/hidden dims must be multiples of 64, out_dim of 32")
        self.in_dim, self.hidden_dim, self.out_dim = in_dim, hidden_dim, out_dim

    def state_dict_shapes(self):
        i, h, o = self.in_dim, self.hidden_dim, self.out_dim
        return {"net.0.weight": (h, i), "net.0.bias": (h,), "net.3.weight": (h,), "net.3.bias": (h,),
                "net.4.weight": (o, h), "net.4.bias": (o,)}

    def load_state_dict(self, state_dict, strict: bool = True):
        self._load_check(state_dict, self.state_dict_shapes(), strict)
        dev, dt = self._device, self._dtype
        for k, v in state_dict.items():
            is_mat = k in ("net.0.weight", "net.4.weight")
            self._w[k] = v.detach().to(device=dev, dtype=dt if is_mat else torch.float32).contiguous()
        self._arena = None
        self._loaded = True
        self._weights_version += 1

    @torch.no_grad()
    def forward(self, x):
        self._check(x)
        w, dt = self._w, self._dtype
        shp = x.shape
        a = x.to(dt).reshape(-1, shp[-1]).contiguous()
        h = ops.gemm(a, w["net.0.weight"], bias=w["net.0.bias"], gelu=True)
        h = ops.layernorm(h, w["net.3.weight"], w["net.3.bias"], 1e-5)
        y = ops.gemm(h, w["net.4.weight"], bias=w["net.4.bias"])
        return y.view(*shp[:-1], self.out_dim)


class B200ControlNetConditioningEmbedding(_Module):
    def __init__(self, conditioning_embedding_channels=320, conditioning_channels=3,
                 block_out_channels=(16, 32, 96, 256), dtype=torch.float16, device="cuda"):
        super().__init__(dtype, device)
        if conditioning_channels > 64 or conditioning_embedding_channels % 32:
            raise NotImplementedError("B200ControlNetConditioningEmbedding: <= 64 input channels, out % 32 == 0")
        self.out_channels, self.in_channels = conditioning_embedding_channels, conditioning_channels
        self.block_out_channels = tuple(block_out_channels)
        # (key, cin, cout, stride) in execution order
        self._layers = [("conv_in", conditioning_channels, block_out_channels[0], 1)]
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self._layers.append((f"blocks.{2 * i}", cin, cin, 1))
            self._layers.append((f"blocks.{2 * i + 1}", cin, cout, 2))
        self._layers.append(("conv_out", block_out_channels[-1], conditioning_embedding_channels, 1))
        ops.ensure_workspace(self._device)

    def state_dict_shapes(self):
        sh = {}
        for k, cin, cout, _ in self._layers:
            sh[f"{k}.weight"] = (cout, cin, 3, 3)
            sh[f"{k}.bias"] = (cout,)
        return sh

    def load_state_dict(self, state_dict, strict: bool = True):
        self._load_check(state_dict, self.state_dict_shapes(), strict)
        dev, dt = self._device, self._dtype
        last = self._layers[-1][0]
        for k, cin, cout, _ in self._layers:
            cin_p = _pad64(cin)
            cout_p = cout if k == last else _pad64(cout)
            wt = torch.zeros(cout_p, cin_p, 3, 3)
            wt[:cout, :cin] = state_dict[f"{k}.weight"].detach().float()
            b = torch.zeros(cout_p)
            b[:cout] = state_dict[f"{k}.bias"].detach().float()
            self._w[f"{k}.weight"] = ops.pack_conv3x3_weight(wt, dt).to(dev)
            self._w[f"{k}.bias"] = b.to(dev)
        self._arena = None
        self._loaded = True
        self._weights_version += 1

    @torch.no_grad()
    def forward(self, conditioning):
        self._check(conditioning)
        w = self._w
        B, C, H, W = conditioning.shape
        if C != self.in_channels:
            raise ValueError(f"expected {self.in_channels} conditioning channels, got {C}")
        x = ops.nchw_to_nhwc_pad(conditioning.contiguous(), 64, self._dtype)
        last = self._layers[-1][0]
        for k, _, _, stride in self._layers:
            x = ops.conv3x3(x, w[f"{k}.weight"], bias=w[f"{k}.bias"], stride=stride, silu=(k != last))
        return ops.nhwc_to_nchw(x, self.out_channels, self._dtype)
