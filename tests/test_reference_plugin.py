"""NS-2 — "stage2_batchtest_inpaint_model.py runs unmodified": the REFERENCE's own pipeline classes, imported from
/root/reference and run unmodified over oracle/diffusers_shim, with this repo's objects plugged into them exactly the
way the reference driver plugs its own (stage2_batchtest_inpaint_model.py:123-133):

    pipe = Stage2_InpaintDiffusionPipeline(...)                       # reference class
    pipe.unet = <UNet>.from_config / load_state_dict(unet_dict)       # -> B200UNet2DConditionModel
    pipe.scheduler = UniPCMultistepScheduler.from_config(pipe.scheduler.config)   # -> B200UniPCMultistepScheduler
    pipe.enable_xformers_memory_efficient_attention()
    pipe(height=..., s_img_proj_f=..., st_pose_f=..., pred_t_img_embed=..., ...)   # the reference's own __call__

CPU test: the CUDA kernels are replaced by the torch stand-ins of tests/mock_ops.py (each follows its C-ABI entry
point's documented semantics), so what is proven is the plug-in SURFACE — every attribute, keyword and dtype the
reference's `__call__` (stage2_inpaint_pipeline.py:391-541; PCDMs_pipeline.py:889-1180) touches on the unet and the
scheduler exists and means the same thing.  Skipped where /root/reference is absent (the GPU box).
"""
from dataclasses import asdict, replace

import pytest
import torch

from oracle import reference_shim as rs
from oracle.factory import make_inputs, make_unet
from oracle.unet import UNetConfig
from tests import mock_ops

pytestmark = pytest.mark.skipif(not rs.reference_available(), reason="/root/reference not present")


def _envelope(got, want, frac, label):
    got, want = got.float(), want.float()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    assert not torch.isnan(got).any(), label
    assert err <= frac * scale, f"{label}: max|err| {err:.3e} vs {frac} * max|ref| {scale:.3e}"


def _b200_unet(cfg, sd, dtype):
    from pcdms_b200.unet import B200UNet2DConditionModel
    m = B200UNet2DConditionModel(dtype=dtype, device="cpu", **asdict(cfg))
    m.load_state_dict(sd)
    return m


def _reference_stage2_pipe(cfg, sd, pin):
    rs._enable()
    from diffusers.schedulers import DDIMScheduler
    from src.pipelines.stage2_inpaint_pipeline import Stage2_InpaintDiffusionPipeline
    unet = rs.build_reference_unet(cfg)
    unet.load_state_dict(sd, strict=True)
    return Stage2_InpaintDiffusionPipeline(vae=rs._FakeVAE(pin["masked_latents"]), unet=unet.half(),
                                           scheduler=DDIMScheduler())


def _call(pipe, pin, steps, n):
    h, w = pin["latents"].shape[-2:]
    return pipe(height=pin["height"], width=pin["width"], num_inference_steps=steps, guidance_scale=2.0,
                num_images_per_prompt=n, latents=pin["latents"].half(), output_type="pt",
                vae_image=torch.zeros(1, 3, h * 8, w * 8), s_img_proj_f=pin["s_img_proj_f"],
                st_pose_f=pin["st_pose_f"], pred_t_img_embed=pin["pred_t_img_embed"]).images


@pytest.mark.parametrize("sched", ["ddim", "unipc"])
def test_reference_stage2_pipeline_runs_with_b200_unet_and_scheduler(sched):
    from pcdms_b200.scheduler import B200DDIMScheduler, B200UniPCMultistepScheduler
    cfg = UNetConfig.tiny()
    sd = make_unet(cfg, seed=0).state_dict()
    pin = make_inputs(cfg, n=2, h=16, w=32, s_kv=9)
    steps = 4

    # the reference as it is: its own UNet (fp16, as its loop forces, :431-501) + the shim's scheduler
    ref_pipe = _reference_stage2_pipe(cfg, sd, pin)
    if sched == "unipc":
        from diffusers.schedulers import UniPCMultistepScheduler
        ref_pipe.scheduler = UniPCMultistepScheduler.from_config(ref_pipe.scheduler.config)   # driver :132
    want = _call(ref_pipe, pin, steps, 2)

    # the same reference pipeline object with this repo's UNet / scheduler assigned, driver-style
    pipe = _reference_stage2_pipe(cfg, sd, pin)
    with mock_ops.patched():
        pipe.unet = _b200_unet(cfg, sd, torch.float16)                                         # driver :125-130
        if sched == "unipc":
            pipe.scheduler = B200UniPCMultistepScheduler.from_config(pipe.scheduler.config)    # driver :132
        else:
            pipe.scheduler = B200DDIMScheduler.from_config(pipe.scheduler.config)
        pipe.enable_xformers_memory_efficient_attention()                                      # driver :133
        assert pipe._execution_device == pipe.unet.device                                      # :225-243 walks unet.modules()
        walked = [m for m in pipe.unet.modules()]
        assert walked[0] is pipe.unet and len(walked) == 1 + 32                                # 16 blocks x (attn1, attn2)
        for m in walked:
            m.set_use_memory_efficient_attention_xformers(True, None)                          # harmless no-op
        got = _call(pipe, pin, steps, 2)
    assert got.shape == want.shape and got.dtype == want.dtype
    # both sides run the net in fp16 with different op granularity: envelope of fp16 accumulation, far below any
    # functional slip (a swapped CFG half, a missing pose add or a wrong token order moves the result by O(1))
    _envelope(got, want, 5e-3, f"reference pipeline + B200 unet/{sched}")   # measured 1.8e-3: 1-2 fp16 ulps of the output


def test_reference_demo_pipeline_runs_with_b200_unet():
    """`pipe.unet` inside PCDMsPipeline (pcdms_demo.ipynb): 9-channel UNet without class embedding, cond_pose added
    after conv_in, tokens through prompt_embeds / negative_prompt_embeds."""
    rs._enable()
    from pcdms_b200.scheduler import B200DDIMScheduler
    cfgd = replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True)
    ud = make_unet(cfgd, seed=9)
    g = torch.Generator().manual_seed(5)
    h, w = 8, 16
    kw = dict(latents=torch.randn(1, 4, h, w, generator=g), simg_mask_latents=torch.randn(1, 4, h, w, generator=g),
              mask=torch.cat([torch.ones(1, 1, h, w // 2), torch.zeros(1, 1, h, w // 2)], dim=3),
              cond_pose=0.1 * torch.randn(1, cfgd.block_out_channels[0], h, w, generator=g),
              prompt_embeds=torch.randn(1, 7, cfgd.cross_attention_dim, generator=g),
              negative_prompt_embeds=0.3 * torch.randn(1, 7, cfgd.cross_attention_dim, generator=g),
              num_inference_steps=3, guidance_scale=2.0)
    want = rs.run_reference_demo_pipeline(cfgd, make_unet(cfgd, seed=9).half(), **kw)

    # what the demo notebook's `pipe.unet` has to offer beyond forward(): config.time_cond_proj_dim (:1100),
    # encoder_hid_proj (:1067), config.in_channels
    b200 = _b200_unet(cfgd, ud.state_dict(), torch.float16)
    assert b200.config.time_cond_proj_dim is None and b200.config.in_channels == 9 and b200.encoder_hid_proj is None
    with mock_ops.patched():
        got = rs.run_reference_demo_pipeline(cfgd, b200, scheduler=B200DDIMScheduler(), raw_unet=True, **kw)
    _envelope(got, want, 5e-3, "reference PCDMsPipeline + B200 unet")   # measured 1.3e-3
