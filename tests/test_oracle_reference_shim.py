"""Live pin: run the reference's OWN UNet class and pipeline (unmodified, from /root/reference) on top of
oracle/diffusers_shim and require bit-equality with oracle/unet.py + oracle/pipeline.py.  Skipped where
/root/reference is absent (the GPU box); tests/golden/ carries the same comparison there."""
import pytest
import torch

from oracle import reference_shim as rs
from oracle.factory import make_inputs, make_unet, make_unet_inputs
from oracle.pipeline import denoise_loop, prepare_conditioning
from oracle.schedulers import OracleDDIMScheduler
from oracle.unet import UNetConfig

pytestmark = pytest.mark.skipif(not rs.reference_available(), reason="/root/reference not present")


def test_reference_unet_forward_equals_oracle():
    cfg = UNetConfig.tiny()
    oracle = make_unet(cfg, seed=3)
    ref = rs.build_reference_unet(cfg)
    assert set(ref.state_dict().keys()) == set(oracle.state_dict().keys())
    ref.load_state_dict(oracle.state_dict(), strict=True)
    i = make_unet_inputs(cfg, batch=2, h=8, w=16, s_kv=5, seed=11)
    with torch.no_grad():
        want = ref(i["sample"], 501, i["encoder_hidden_states"], class_labels=i["class_labels"],
                   my_pose_cond=i["my_pose_cond"], return_dict=False)[0]
    got = oracle(i["sample"], 501, i["encoder_hidden_states"], class_labels=i["class_labels"],
                 my_pose_cond=i["my_pose_cond"])[0]
    assert torch.equal(got, want)


def test_reference_unet_config_and_processor_surface():
    ref = rs.build_reference_unet(UNetConfig.tiny())
    assert ref.config.in_channels == 9 and ref.config.class_embed_type == "projection"
    assert len(ref.attn_processors) == 32  # 16 transformer blocks x (attn1, attn2)
    ref.set_default_attn_processor()


def test_reference_pipeline_equals_oracle_loop():
    cfg = UNetConfig.tiny()
    oracle = make_unet(cfg, seed=5)
    pin = make_inputs(cfg, n=1, h=8, w=16, s_kv=6, seed=21)
    want = rs.run_reference_pipeline(cfg, oracle.state_dict(), pin, num_inference_steps=3, guidance_scale=2.0,
                                     num_images_per_prompt=1)
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"],
                                height=pin["height"], width=pin["width"], num_images_per_prompt=1,
                                guidance_scale=2.0, dtype=torch.float16)
    got = denoise_loop(oracle.half(), OracleDDIMScheduler(), latents=pin["latents"], cond=cond,
                       num_inference_steps=3, guidance_scale=2.0, dtype=torch.float16)
    assert torch.equal(got, want)


@pytest.mark.parametrize("sched", ["ddim", "unipc"])
def test_reference_stage3_pipeline_equals_oracle_loop(sched):
    """The reference's own Stage3_RefinedPipeline.__call__ (fp16 loop tensors, fp32 UNet evaluation) against
    oracle.pipeline.denoise_loop_stage3 run with the same casts: bit-equal, under both schedulers."""
    from oracle.pipeline import denoise_loop_stage3
    from oracle.unipc import UniPCMultistepScheduler
    cfg = UNetConfig.tiny(in_channels=8, stage2=False)
    oracle = make_unet(cfg, seed=7)
    g = torch.Generator().manual_seed(31)
    lat, gl = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    f = torch.randn(1, 5, cfg.cross_attention_dim, generator=g)
    mk = (lambda: OracleDDIMScheduler()) if sched == "ddim" else (lambda: UniPCMultistepScheduler(
        beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1, timestep_spacing="leading"))
    want = rs.run_reference_stage3_pipeline(cfg, oracle, latents=lat, gen_t_img_latents=gl, s_img_proj_f=f,
                                            num_inference_steps=4, guidance_scale=2.0, scheduler=mk())

    got = denoise_loop_stage3(oracle, mk(), latents=lat.half(), gen_t_img_latents=gl, s_img_proj_f=f,
                              num_inference_steps=4, guidance_scale=2.0, dtype=torch.float16,
                              unet_dtype=torch.float32)
    assert torch.equal(got.float(), want.float())


def test_reference_demo_pipeline_equals_oracle_loop():
    """The reference's own PCDMsPipeline.__call__ (the pcdms_demo.ipynb driver; fp16 as it hard-codes) against
    oracle.pipeline.denoise_loop_pcdms: bit-equal."""
    from dataclasses import replace
    from oracle.pipeline import denoise_loop_pcdms
    cfg = replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True)
    oracle = make_unet(cfg, seed=9).half()
    h, w = 8, 16
    g = torch.Generator().manual_seed(41)
    lat, msk_l = torch.randn(1, 4, h, w, generator=g), torch.randn(1, 4, h, w, generator=g)
    mask = torch.cat([torch.ones(1, 1, h, w // 2), torch.zeros(1, 1, h, w // 2)], dim=3)
    pose = 0.1 * torch.randn(1, cfg.block_out_channels[0], h, w, generator=g)
    pe = torch.randn(1, 7, cfg.cross_attention_dim, generator=g)
    ne = 0.3 * torch.randn(1, 7, cfg.cross_attention_dim, generator=g)
    kw = dict(latents=lat, mask=mask, simg_mask_latents=msk_l, cond_pose=pose, prompt_embeds=pe,
              negative_prompt_embeds=ne, num_inference_steps=3, guidance_scale=2.0)
    want = rs.run_reference_demo_pipeline(cfg, oracle, **kw)
    got = denoise_loop_pcdms(oracle, OracleDDIMScheduler(), dtype=torch.float16, **kw)
    assert torch.equal(got, want)


def test_reference_simple_stage2_pipeline_equals_oracle_loop():
    """The reference's own Simple_Stage2_InpaintDiffusionPipeline.__call__ (stage2_inpaint_pipeline.py:757-877; fp16 as it
    hard-codes) against oracle.pipeline.denoise_loop_simple: bit-equal."""
    from dataclasses import replace
    from oracle.pipeline import denoise_loop_simple
    cfg = replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True)
    oracle = make_unet(cfg, seed=12).half()
    h, w = 8, 16
    g = torch.Generator().manual_seed(51)
    kw = dict(latents=torch.randn(2, 4, h, w, generator=g), masked_latents=torch.randn(1, 4, h, w, generator=g).half(),
              st_pose_f=(0.1 * torch.randn(1, cfg.block_out_channels[0], h, w, generator=g)).half(),
              s_img_proj_f=torch.randn(1, 7, cfg.cross_attention_dim, generator=g).half(), height=h * 8, width=w * 8,
              num_inference_steps=3, guidance_scale=2.0, num_images_per_prompt=2)
    want = rs.run_reference_simple_stage2_pipeline(cfg, oracle, **kw)
    got = denoise_loop_simple(oracle, OracleDDIMScheduler(), dtype=torch.float16, **kw)
    assert torch.equal(got, want)
