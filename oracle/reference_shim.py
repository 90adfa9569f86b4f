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


class _EncodeVAE:
    """Identity-scaled stand-in whose encode() returns given latents (sample ignores the generator) and whose decode()
    is the identity: exposes the loop of a pipeline without the VAE."""

    def __init__(self, latents):
        self.config = SimpleNamespace(block_out_channels=(1, 2, 3, 4), scaling_factor=1.0)
        self._lat = latents

    def encode(self, x):
        lat = self._lat.to(x.dtype)
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda generator=None: lat))

    def decode(self, z, return_dict=False, generator=None):
        return (z,)


class _Stage3UNetAdapter(torch.nn.Module):
    """The stock diffusers UNet2DConditionModel call surface the stage-3 pipeline uses
    (stage3_refined_pipeline.py:541-543) over the oracle UNet with the stage-3 config."""

    def __init__(self, unet, cfg):
        super().__init__()
        self.unet = unet
        self.config = SimpleNamespace(**asdict(cfg))

    @property
    def device(self):
        return torch.device("cpu")

    def forward(self, sample, timestep, encoder_hidden_states=None, cross_attention_kwargs=None, return_dict=True):
        assert cross_attention_kwargs is None and not return_dict
        return self.unet(sample, timestep, encoder_hidden_states)


def run_reference_stage3_pipeline(cfg, oracle_unet, *, latents, gen_t_img_latents, s_img_proj_f, num_inference_steps,
                                  guidance_scale, scheduler):
    """Run Stage3_RefinedPipeline.__call__ (stage3_refined_pipeline.py:443-578) unmodified on CPU: fp16 latents and
    conditioning, fp32 UNet evaluation (:538-542), identity VAE.  Returns the final latents."""
    _enable()
    from src.pipelines.stage3_refined_pipeline import Stage3_RefinedPipeline

    pipe = Stage3_RefinedPipeline(vae=_EncodeVAE(gen_t_img_latents), unet=_Stage3UNetAdapter(oracle_unet, cfg),
                                  scheduler=scheduler)
    h, w = latents.shape[-2:]
    out = pipe(height=h * 8, width=w * 8, num_inference_steps=num_inference_steps, guidance_scale=guidance_scale,
               num_images_per_prompt=1, latents=latents.half(), output_type="pt",
               vae_gen_t_image=torch.zeros(1, 3, h * 8, w * 8), s_img_proj_f=s_img_proj_f)
    return out.images


class _DemoUNetAdapter(torch.nn.Module):
    """The call surface PCDMsPipeline uses on its UNet (PCDMs_pipeline.py:1121-1130; `config.time_cond_proj_dim`
    :1100, `encoder_hid_proj` :1067) over the oracle UNet (9 input channels, no class embedding, pose add)."""

    def __init__(self, unet, cfg):
        super().__init__()
        self.unet = unet
        self.config = SimpleNamespace(**asdict(cfg))
        self.encoder_hid_proj = None

    @property
    def device(self):
        return torch.device("cpu")

    @property
    def dtype(self):
        return next(self.unet.parameters()).dtype

    def forward(self, sample, timestep, encoder_hidden_states=None, my_pose_cond=None, timestep_cond=None,
                cross_attention_kwargs=None, added_cond_kwargs=None, return_dict=True):
        assert timestep_cond is None and cross_attention_kwargs is None and added_cond_kwargs is None
        return self.unet(sample, timestep, encoder_hidden_states, my_pose_cond=my_pose_cond)


def run_reference_demo_pipeline(cfg, oracle_unet_half, *, latents, mask, simg_mask_latents, cond_pose, prompt_embeds,
                                negative_prompt_embeds, num_inference_steps, guidance_scale, scheduler=None,
                                raw_unet=False):
    """Run PCDMsPipeline.__call__ (PCDMs_pipeline.py:889-1180, the pcdms_demo.ipynb driver) unmodified on CPU.  The
    reference hard-codes fp16 for the UNet input (:1115), so the UNet and the conditioning are fp16 here.  Returns the
    final latents (identity VAE).  `raw_unet=True` hands the given object to the pipeline as `unet` without the adapter
    (plug-in tests: a B200UNet2DConditionModel), `scheduler` replaces the shim's DDIM."""
    _enable()
    from diffusers.schedulers import DDIMScheduler
    from src.pipelines.PCDMs_pipeline import PCDMsPipeline

    pipe = PCDMsPipeline(vae=_EncodeVAE(latents), text_encoder=None, tokenizer=None,
                         unet=oracle_unet_half if raw_unet else _DemoUNetAdapter(oracle_unet_half, cfg),
                         scheduler=scheduler or DDIMScheduler(), safety_checker=None,
                         feature_extractor=None, requires_safety_checker=False)
    h, w = latents.shape[-2:]
    out = pipe(simg_mask_latents=simg_mask_latents, mask=mask, cond_pose=cond_pose.half(),
               prompt_embeds=prompt_embeds.half(), negative_prompt_embeds=negative_prompt_embeds.half(),
               height=h * 8, width=w * 8, num_images_per_prompt=1, guidance_scale=guidance_scale,
               latents=latents.half(), num_inference_steps=num_inference_steps, output_type="pt")
    return out.images


def build_reference_prior(**cfg):
    """Instantiate the reference's Stage1_PriorTransformer (src/models/stage1_prior_transformer.py:50) on the shim."""
    _enable()
    from src.models.stage1_prior_transformer import Stage1_PriorTransformer

    model = Stage1_PriorTransformer(**cfg)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def run_reference_prior_pipeline(ref_prior, *, s_embed, s_pose, t_pose, latents, num_inference_steps, guidance_scale,
                                 zero_embed):
    """Run Stage1_PriorPipeline.__call__ (src/pipelines/stage1_prior_pipeline.py:357-505) unmodified on CPU.  The
    variance noise of each step comes from torch's global generator (the reference passes none to scheduler.step,
    :478-483): seed it before calling.  `zero_embed` stands in for the CLIP embedding of a black image
    (get_zero_embed, :282-289) — the image encoder is outside the loop.  Returns (image_embeds, negative_image_embeds)."""
    _enable()
    from diffusers.schedulers import UnCLIPScheduler
    from src.pipelines.stage1_prior_pipeline import Stage1_PriorPipeline

    class _Encoder:
        config = SimpleNamespace(image_size=8)
        dtype = torch.float32

        def __call__(self, x):
            return {"image_embeds": zero_embed}

    pipe = Stage1_PriorPipeline(prior=ref_prior, image_encoder=_Encoder(), scheduler=UnCLIPScheduler(),
                                image_processor=None)
    out = pipe(s_embed=s_embed, s_pose=s_pose, t_pose=t_pose, num_images_per_prompt=1,
               num_inference_steps=num_inference_steps, latents=latents, guidance_scale=guidance_scale)
    return out[0] if isinstance(out, tuple) else (out["image_embeds"], out["negative_image_embeds"])


def run_reference_simple_stage2_pipeline(cfg, oracle_unet_half, *, latents, s_img_proj_f, st_pose_f, masked_latents,
                                         height, width, num_inference_steps, guidance_scale, num_images_per_prompt):
    """Run Simple_Stage2_InpaintDiffusionPipeline.__call__ (stage2_inpaint_pipeline.py:757-877) unmodified on CPU over
    the oracle UNet without class embedding (fp16, as the reference hard-codes :794,803,811,843,856).  Returns the
    final latents (identity VAE)."""
    _enable()
    from diffusers.schedulers import DDIMScheduler
    from src.pipelines.stage2_inpaint_pipeline import Simple_Stage2_InpaintDiffusionPipeline

    pipe = Simple_Stage2_InpaintDiffusionPipeline(vae=_EncodeVAE(masked_latents), unet=_DemoUNetAdapter(oracle_unet_half, cfg),
                                                  scheduler=DDIMScheduler())
    h, w = latents.shape[-2:]
    out = pipe(height=height, width=width, num_inference_steps=num_inference_steps, guidance_scale=guidance_scale,
               num_images_per_prompt=num_images_per_prompt, latents=latents.half(), output_type="pt",
               vae_image=torch.zeros(1, 3, h * 8, w * 8), s_img_proj_f=s_img_proj_f, st_pose_f=st_pose_f)
    return out.images
