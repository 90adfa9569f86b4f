"""The three reference drivers on the fused engine: stage-2 with VAE in and out and the
batch-test scheduler (UniPC), the demo `PCDMsPipeline` surface, and the stage-3 refiner.

CPU: host orchestration against the oracle loops (oracle/pipeline.py, each citing the reference lines it follows) with
the kernels replaced by tests/mock_ops.py.  GPU: the real kernels, CUDA-graph replay included, against the same
oracles; the full image-in -> image-out chain against oracle VAE + oracle loop.
"""
from dataclasses import asdict, replace

import pytest
import torch

from oracle.factory import make_inputs, make_unet
from oracle.pipeline import denoise_loop, denoise_loop_pcdms, denoise_loop_stage3, prepare_conditioning
from oracle.schedulers import OracleDDIMScheduler
from oracle.unet import UNetConfig
from oracle.unipc import UniPCMultistepScheduler as OracleUniPC
from oracle.vae import VAEConfig, make_vae
from pcdms_b200.pipeline import B200PCDMsPipeline, B200Stage2InpaintPipeline, B200Stage3RefinedPipeline
from pcdms_b200.scheduler import B200DDIMScheduler, B200UniPCMultistepScheduler
from pcdms_b200.unet import B200UNet2DConditionModel
from pcdms_b200.vae import B200AutoencoderKL
from tests import mock_ops
from tests.test_unipc import SD21_SCHED


def _unet(cfg, dtype, device, seed=0):
    o = make_unet(cfg, seed=seed)
    m = B200UNet2DConditionModel(dtype=dtype, device=device, **asdict(cfg))
    m.load_state_dict(o.state_dict())
    return o, m


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _demo_case(dtype, device, steps):
    cfg = replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True)
    o, m = _unet(cfg, dtype, device)
    h, w = 8, 16
    lat, msk_l = torch.randn(1, 4, h, w, generator=_g(1)), torch.randn(1, 4, h, w, generator=_g(2))
    mask = torch.cat([torch.ones(1, 1, h, w // 2), torch.zeros(1, 1, h, w // 2)], dim=3)
    pose = 0.1 * torch.randn(1, 64, h, w, generator=_g(3))
    pe, ne = torch.randn(1, 7, 128, generator=_g(4)), 0.3 * torch.randn(1, 7, 128, generator=_g(5))
    want = denoise_loop_pcdms(o, OracleDDIMScheduler(), latents=lat, mask=mask, simg_mask_latents=msk_l,
                              cond_pose=pose, prompt_embeds=pe, negative_prompt_embeds=ne,
                              num_inference_steps=steps)
    pipe = B200PCDMsPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    call = lambda: pipe(simg_mask_latents=msk_l, mask=mask, cond_pose=pose, prompt_embeds=pe,
                        negative_prompt_embeds=ne, height=h * 8, width=w * 8, num_images_per_prompt=1,
                        guidance_scale=2.0, latents=lat, num_inference_steps=steps, output_type="latent").images
    return pipe, call, want


def _simple_case(dtype, device, steps):
    """Simple_Stage2_InpaintDiffusionPipeline (stage2_inpaint_pipeline.py:544-877): tokens only, no class embedding."""
    from oracle.pipeline import denoise_loop_simple
    from pcdms_b200.pipeline import B200SimpleStage2InpaintPipeline
    cfg = replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True)
    o, m = _unet(cfg, dtype, device)
    h, w = 8, 16
    kw = dict(latents=torch.randn(2, 4, h, w, generator=_g(1)), masked_latents=torch.randn(1, 4, h, w, generator=_g(2)),
              st_pose_f=0.1 * torch.randn(1, 64, h, w, generator=_g(3)), s_img_proj_f=torch.randn(1, 7, 128, generator=_g(4)),
              height=h * 8, width=w * 8, num_inference_steps=steps, guidance_scale=2.0, num_images_per_prompt=2)
    want = denoise_loop_simple(o, OracleDDIMScheduler(), **kw)
    pipe = B200SimpleStage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    call = lambda: pipe(output_type="latent", pred_t_img_embed=torch.zeros(1, 1, 128), **kw).images
    return pipe, call, want


def _stage3_case(dtype, device, steps):
    cfg = UNetConfig.tiny(in_channels=8, stage2=False)
    o, m = _unet(cfg, dtype, device)
    h = w = 16
    lat, gl = torch.randn(1, 4, h, w, generator=_g(1)), torch.randn(1, 4, h, w, generator=_g(2))
    f = torch.randn(1, 9, 128, generator=_g(3))
    want = denoise_loop_stage3(o, OracleUniPC.from_config(SD21_SCHED), latents=lat, gen_t_img_latents=gl,
                               s_img_proj_f=f, num_inference_steps=steps)
    pipe = B200Stage3RefinedPipeline(vae=None, unet=m, scheduler=B200UniPCMultistepScheduler.from_config(SD21_SCHED))
    call = lambda: pipe(height=h * 8, width=w * 8, num_inference_steps=steps, guidance_scale=2.0, latents=lat,
                        s_img_proj_f=f, gen_t_img_latents=gl, output_type="latent").images
    return pipe, call, want


@pytest.mark.parametrize("case", [_demo_case, _stage3_case, _simple_case])
def test_driver_orchestration_matches_oracle(case):
    pipe, call, want = case(torch.float32, "cpu", 4)
    pipe.use_cuda_graph = False
    with mock_ops.patched():
        for _ in range(2):
            torch.testing.assert_close(call(), want, rtol=2e-4, atol=2e-5)


def test_driver_surfaces():
    pipe, _, _ = _demo_case(torch.float32, "cpu", 1)
    with pytest.raises(ValueError):
        pipe(height=64, width=128, guidance_scale=2.0)                       # no prompt_embeds
    with pytest.raises(NotImplementedError):
        pipe(prompt="a photo", height=64, width=128, guidance_scale=2.0)
    p3, _, _ = _stage3_case(torch.float32, "cpu", 1)
    with pytest.raises(ValueError):
        p3(height=128, width=128, guidance_scale=2.0, s_img_proj_f=torch.zeros(1, 9, 128))   # no VAE, no latents
    with pytest.raises(ValueError):
        p3(height=100, width=128, guidance_scale=2.0, s_img_proj_f=torch.zeros(1, 9, 128))
    assert p3.vae_scale_factor == 8 and p3.device.type == "cpu"
    p3.enable_xformers_memory_efficient_attention()


def _image_chain_oracle(o_unet, o_vae, sched, pin, vae_image, steps, noise_seed):
    with torch.no_grad():
        d = o_vae.encode(vae_image).latent_dist
        masked = d.sample(generator=_g(noise_seed)) * o_vae.cfg.scaling_factor
        cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                    st_pose_f=pin["st_pose_f"], masked_latents=masked, height=pin["height"],
                                    width=pin["width"], num_images_per_prompt=pin["latents"].shape[0],
                                    guidance_scale=2.0)
        lat = denoise_loop(o_unet, sched, latents=pin["latents"], cond=cond, num_inference_steps=steps,
                           guidance_scale=2.0)
        return lat, o_vae.decode(lat / o_vae.cfg.scaling_factor).sample


def test_image_in_image_out_host_logic():
    """vae.encode -> loop (UniPC, the batch-test scheduler) -> vae.decode -> PIL, host side, against the oracle chain."""
    cfg, vcfg = UNetConfig.tiny(), VAEConfig.tiny()
    o, m = _unet(cfg, torch.float32, "cpu")
    ov = make_vae(vcfg, seed=2)
    pv = B200AutoencoderKL(dtype=torch.float32, device="cpu", **asdict(vcfg))
    pv.load_state_dict(ov.state_dict())
    pv._check_input = lambda *a, **k: None
    pin = make_inputs(cfg, n=2, h=8, w=16, s_kv=6)
    img = torch.rand(1, 3, 64, 128, generator=_g(9)) * 2 - 1
    lat, want = _image_chain_oracle(o, ov, OracleUniPC.from_config(SD21_SCHED), pin, img, 3, 11)
    pipe = B200Stage2InpaintPipeline(vae=pv, unet=m, scheduler=B200UniPCMultistepScheduler.from_config(SD21_SCHED))
    pipe.use_cuda_graph = False
    with mock_ops.patched():
        kw = dict(height=64, width=128, num_inference_steps=3, guidance_scale=2.0, num_images_per_prompt=2,
                  latents=pin["latents"], vae_image=img, generator=_g(11), s_img_proj_f=pin["s_img_proj_f"],
                  st_pose_f=pin["st_pose_f"], pred_t_img_embed=pin["pred_t_img_embed"])
        got = pipe(output_type="pt", **kw).images
        torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-4)
        pil = pipe(output_type="pil", **{**kw, "generator": _g(11)}).images
    assert len(pil) == 2 and pil[0].size == (128, 64)


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
def _rel(got, want):
    return float((got.float().cpu() - want).abs().max() / want.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("case", [_demo_case, _stage3_case, _simple_case])
@pytest.mark.parametrize("use_graph", [False, True])
def test_drivers_gpu(case, use_graph):
    pipe, call, want = case(torch.float16, "cuda", 8)
    pipe.use_cuda_graph = use_graph
    for _ in range(2):
        assert _rel(call(), want) < 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize("sched", ["ddim", "unipc"])
def test_stage2_with_scheduler_gpu(sched):
    """BASELINE config-1 style run (tiny weights) under both schedulers, fused graph path vs oracle loop."""
    cfg = UNetConfig.tiny()
    o, m = _unet(cfg, torch.float16, "cuda")
    pin = make_inputs(cfg, n=2, h=16, w=32, s_kv=9)
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"], height=pin["height"],
                                width=pin["width"], num_images_per_prompt=2, guidance_scale=2.0)
    os_, ps = ((OracleDDIMScheduler(), B200DDIMScheduler()) if sched == "ddim" else
               (OracleUniPC.from_config(SD21_SCHED), B200UniPCMultistepScheduler.from_config(SD21_SCHED)))
    want = denoise_loop(o, os_, latents=pin["latents"], cond=cond, num_inference_steps=20, guidance_scale=2.0)
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=ps)
    for _ in range(2):
        got = pipe(height=pin["height"], width=pin["width"], num_inference_steps=20, guidance_scale=2.0,
                   num_images_per_prompt=2, latents=pin["latents"], output_type="latent",
                   s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
                   pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images
        assert _rel(got, want) < 3e-3


@pytest.mark.gpu
def test_image_in_image_out_gpu():
    cfg, vcfg = UNetConfig.tiny(), VAEConfig.tiny()
    o, m = _unet(cfg, torch.float16, "cuda")
    ov = make_vae(vcfg, seed=2)
    pv = B200AutoencoderKL(dtype=torch.float16, **asdict(vcfg))
    pv.load_state_dict(ov.state_dict())
    pin = make_inputs(cfg, n=2, h=16, w=32, s_kv=9)
    img = torch.rand(1, 3, 128, 256, generator=_g(9)) * 2 - 1
    lat, want = _image_chain_oracle(o, ov, OracleUniPC.from_config(SD21_SCHED), pin, img, 10, 11)
    pipe = B200Stage2InpaintPipeline(vae=pv, unet=m, scheduler=B200UniPCMultistepScheduler.from_config(SD21_SCHED))
    kw = dict(height=128, width=256, num_inference_steps=10, guidance_scale=2.0, num_images_per_prompt=2,
              latents=pin["latents"], vae_image=img, s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
              pred_t_img_embed=pin["pred_t_img_embed"])
    got = pipe(output_type="pt", generator=_g(11), **kw).images
    assert got.shape == (2, 3, 128, 256)
    assert _rel(got, want) < 1e-2
    pil = pipe(output_type="pil", generator=_g(11), **kw).images
    assert len(pil) == 2 and pil[0].size == (256, 128)


@pytest.mark.gpu
def test_drivers_against_reference_goldens_gpu():
    """Product pipelines against tensors produced by the REFERENCE's own pipeline classes (tests/golden/*.pt, made by
    tools/make_golden.py in the build container): stage-3 refiner and the demo driver."""
    from pathlib import Path
    gold = Path(__file__).resolve().parent / "golden"
    g = torch.load(gold / "ref_stage3_tiny.pt")
    cfg = UNetConfig.tiny(in_channels=8, stage2=False)
    _, m = _unet(cfg, torch.float16, "cuda", seed=g["seed"])
    i = g["inputs"]
    pipe = B200Stage3RefinedPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    got = pipe(height=64, width=64, num_inference_steps=i["num_inference_steps"], guidance_scale=i["guidance_scale"],
               latents=i["latents"], s_img_proj_f=i["s_img_proj_f"], gen_t_img_latents=i["gen_t_img_latents"],
               output_type="latent").images
    assert _rel(got, g["latents"].float()) < 4e-3
    g = torch.load(gold / "ref_demo_tiny.pt")
    cfg = replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True)
    _, m = _unet(cfg, torch.float16, "cuda", seed=g["seed"])
    i = g["inputs"]
    pipe = B200PCDMsPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    got = pipe(simg_mask_latents=i["simg_mask_latents"], mask=i["mask"], cond_pose=i["cond_pose"],
               prompt_embeds=i["prompt_embeds"], negative_prompt_embeds=i["negative_prompt_embeds"], height=64,
               width=128, num_images_per_prompt=1, guidance_scale=i["guidance_scale"], latents=i["latents"],
               num_inference_steps=i["num_inference_steps"], output_type="latent").images
    assert _rel(got, g["latents"].float()) < 4e-3
    g = torch.load(gold / "ref_simple_tiny.pt")            # Simple_Stage2_InpaintDiffusionPipeline (no class embedding)
    _, m = _unet(cfg, torch.float16, "cuda", seed=g["seed"])
    i = g["inputs"]
    from pcdms_b200.pipeline import B200SimpleStage2InpaintPipeline
    pipe = B200SimpleStage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    got = pipe(height=i["height"], width=i["width"], num_inference_steps=i["num_inference_steps"],
               guidance_scale=i["guidance_scale"], num_images_per_prompt=i["num_images_per_prompt"], latents=i["latents"],
               s_img_proj_f=i["s_img_proj_f"], st_pose_f=i["st_pose_f"], masked_latents=i["masked_latents"],
               output_type="latent").images
    assert got.shape == (2, 4, 8, 16) and _rel(got, g["latents"].float()) < 4e-3
