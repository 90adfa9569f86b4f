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
