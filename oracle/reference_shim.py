"""ORACLE support (test infrastructure only): import and drive the reference's own classes from /root/reference
on top of oracle/diffusers_shim.  Only usable where /root/reference exists (the build container)."""
from __future__ import annotations

import sys
from dataclasses import asdict
from pathlib import Path
from types import SimpleNamespace

import torch

REFERENCE_ROOT = Path("/root/reference")
SHIM_ROOT = Path(__file__).resolve().parent / "diffusers_shim"


def reference_available() -> bool:
    return (REFERENCE_ROOT / "src" / "models" / "stage2_inpaint_unet_2d_condition.py").exists()


def _enable():
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    for p in (str(SHIM_ROOT), str(REFERENCE_ROOT), str(Path(__file__).resolve().parent.parent)):
        if p not in sys.path:
            sys.path.insert(0, p)


def build_reference_unet(cfg):
    """Instantiate the reference's Stage2_InapintUNet2DConditionModel (stage2_inpaint_unet_2d_condition.py:61) with
    the oracle's config."""
    _enable()
    from src.models.stage2_inpaint_unet_2d_condition import Stage2_InapintUNet2DConditionModel

    d = asdict(cfg)
    d.pop("use_pose_cond")
    d.pop("_diffusers_version")
    model = Stage2_InapintUNet2DConditionModel(**d)
    model.eval()
    # inference-only: with requires_grad=True torch's CPU conv picks a different (non-mkldnn) kernel even under
    # no_grad, which changes fp32 rounding by ~1e-6 and would make "bit-equal to the oracle" depend on a flag that
    # is not part of the algorithm.
    for p in model.parameters():
        p.requires_grad_(False)
    return model


class _FakeVAE:
    """Stands in for AutoencoderKL: encode() returns the pre-made masked latents, decode() is the identity, so the
    pipeline's loop can be observed without the (out-of-scope) VAE."""

    def __init__(self, masked_latents_unscaled):
        self.config = SimpleNamespace(block_out_channels=(1, 2, 3, 4), scaling_factor=1.0)  # identity scaling: no extra rounding
        self._lat = masked_latents_unscaled

    def encode(self, x):
        lat = self._lat.to(x.dtype)
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda generator=None: lat))

    def decode(self, z, return_dict=False):
        return (z,)


def run_reference_pipeline(cfg, state_dict, pin, *, num_inference_steps, guidance_scale, num_images_per_prompt):
    """Run Stage2_InpaintDiffusionPipeline.__call__ (stage2_inpaint_pipeline.py:391-541) on CPU.  The reference
    hard-codes fp16 for every loop tensor (:431,440,449,452,487,501), so the UNet runs in fp16 here too.  Returns the
    final latents (output of the last scheduler.step; the fake VAE is the identity with scaling_factor 1)."""
    _enable()
    from diffusers.schedulers import DDIMScheduler
    from src.pipelines.stage2_inpaint_pipeline import Stage2_InpaintDiffusionPipeline

    unet = build_reference_unet(cfg)
    unet.load_state_dict(state_dict, strict=True)
    unet = unet.half()
    vae = _FakeVAE(pin["masked_latents"])
    pipe = Stage2_InpaintDiffusionPipeline(vae=vae, unet=unet, scheduler=DDIMScheduler())
    h, w = pin["latents"].shape[-2:]
    out = pipe(height=pin["height"], width=pin["width"], num_inference_steps=num_inference_steps,
               guidance_scale=guidance_scale, num_images_per_prompt=num_images_per_prompt,
               latents=pin["latents"].half(), output_type="pt", vae_image=torch.zeros(1, 3, h * 8, w * 8),
               s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
               pred_t_img_embed=pin["pred_t_img_embed"])
    return out.images  # = final latents through the identity decoder (scaling_factor 1.0)
