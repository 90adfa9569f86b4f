"""ORACLE (test infrastructure only) — deterministic synthetic weights and inputs for the stage-2 path.

There is no network, hence no SD-2.1 checkpoint and no DeepFashion data: weights are random-initialised with the
real architecture and inputs are seeded tensors of the real shapes.  Everything is generated on the
CPU with torch.Generator so that the build container and the GPU box produce bit-identical tensors.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .unet import OracleUNet, UNetConfig


def make_unet(cfg: UNetConfig, seed: int = 0) -> OracleUNet:
    """PyTorch default initialisation under a fixed seed, plus perturbed norm affines so that gamma/beta paths are
    actually exercised (the default gamma=1, beta=0 would hide a swapped or dropped affine)."""
    prev = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = OracleUNet(cfg)
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, (nn.GroupNorm, nn.LayerNorm)):
                    mod.weight.add_(0.1 * torch.randn(mod.weight.shape, generator=g))
                    mod.bias.add_(0.1 * torch.randn(mod.bias.shape, generator=g))
    finally:
        torch.random.set_rng_state(prev)
    m.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def _randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def make_inputs(cfg: UNetConfig, *, n: int = 1, h: int = 32, w: int = 64, s_kv: int = 258, seed: int = 42):
    """Pipeline-level inputs for `n` images per call (reference: one source/target pair, num_images_per_prompt=n).

    latents          [n, 4, h, w]      randn, seed            (ref prepare_latents, stage2_inpaint_pipeline.py:477-487)
    masked_latents   [1, 4, h, w]      randn seed+1, right half = 0 (stand-in for vae.encode(canvas)*0.18215, :443)
    st_pose_f        [1, C0, h, w]     0.1*randn seed+2       (stand-in for pose_proj(pose canvas), batchtest :173-174)
    s_img_proj_f     [1, s_kv-1, D]    randn seed+3           (ImageProjModel_p(DINOv2), batchtest :165-167)
    pred_t_img_embed [1, 1, D]         randn seed+4           (stage-1 prior embedding, batchtest :176-185)
    """
    d = cfg.cross_attention_dim
    c0 = cfg.block_out_channels[0]
    masked = _randn((1, 4, h, w), seed + 1)
    masked[..., w // 2:] = 0.0
    return dict(
        latents=_randn((n, 4, h, w), seed),
        masked_latents=masked,
        st_pose_f=0.1 * _randn((1, c0, h, w), seed + 2),
        s_img_proj_f=_randn((1, s_kv - 1, d), seed + 3),
        pred_t_img_embed=_randn((1, 1, cfg.projection_class_embeddings_input_dim or d), seed + 4),
        height=h * 8, width=w * 8,
    )


def make_unet_inputs(cfg: UNetConfig, *, batch: int = 2, h: int = 32, w: int = 64, s_kv: int = 258, seed: int = 7):
    """Direct UNet.forward inputs (one step)."""
    d = cfg.cross_attention_dim
    out = dict(
        sample=_randn((batch, cfg.in_channels, h, w), seed),
        encoder_hidden_states=_randn((batch, s_kv, d), seed + 1),
    )
    if cfg.class_embed_type is not None:
        out["class_labels"] = _randn((batch, 1, cfg.projection_class_embeddings_input_dim), seed + 2)
    if cfg.use_pose_cond:
        out["my_pose_cond"] = 0.1 * _randn((batch, cfg.block_out_channels[0], h, w), seed + 3)
    return out
