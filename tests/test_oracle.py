"""Oracle self-checks: the CPU restatement against closed forms, the published parameter count and the golden
fixtures produced by the reference's own classes (tools/make_golden.py)."""
import math
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import blocks as B
from oracle.factory import make_unet
from oracle.pipeline import cfg_combine, denoise_loop, prepare_conditioning
from oracle.schedulers import OracleDDIMScheduler, ddpm_add_noise, scaled_linear_alphas_cumprod
from oracle.unet import OracleUNet, UNetConfig, param_count

GOLD = Path(__file__).resolve().parent / "golden"


def test_parameter_counts_and_key_count():
    # 865 910 724 is the published SD-2.x UNet size: pins the restated topology
    with torch.device("meta"):
        stock = OracleUNet(UNetConfig.sd21_stock())
        s2 = OracleUNet(UNetConfig.stage2())
        s3 = OracleUNet(UNetConfig.stage3())
    assert param_count(stock) == 865_910_724
    assert param_count(s2) == 868_876_804
    assert param_count(s3) == 865_922_244
    assert len(stock.state_dict()) == 686
    assert len(s2.state_dict()) == 690
    assert len(s3.state_dict()) == 686


def test_state_dict_key_skeleton():
    with torch.device("meta"):
        keys = set(OracleUNet(UNetConfig.stage2()).state_dict().keys())
    for k in ["conv_in.weight", "time_embedding.linear_1.weight", "class_embedding.linear_2.bias",
              "down_blocks.0.resnets.0.time_emb_proj.weight", "down_blocks.0.attentions.1.proj_in.weight",
              "down_blocks.2.downsamplers.0.conv.weight", "down_blocks.3.resnets.1.conv2.bias",
              "mid_block.attentions.0.transformer_blocks.0.attn2.to_k.weight", "mid_block.resnets.1.norm2.weight",
              "up_blocks.0.upsamplers.0.conv.weight", "up_blocks.1.resnets.2.conv_shortcut.weight",
              "up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0.proj.weight",
              "up_blocks.3.attentions.0.transformer_blocks.0.ff.net.2.bias",
              "up_blocks.2.attentions.1.transformer_blocks.0.attn1.to_out.0.bias", "conv_norm_out.weight",
              "conv_out.bias"]:
        assert k in keys, k
    assert "down_blocks.3.attentions.0.norm.weight" not in keys
    assert "up_blocks.0.attentions.0.norm.weight" not in keys
    assert "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.bias" not in keys  # q/k/v have no bias


def test_timestep_embedding_closed_form():
    t = torch.tensor([981, 1])
    emb = B.Timesteps(320, True, 0)(t)
    assert emb.shape == (2, 320) and emb.dtype == torch.float32
    i = torch.arange(160, dtype=torch.float32)
    f = torch.exp(-math.log(10000.0) * i / 160)
    want = torch.cat([torch.cos(t[:, None].float() * f), torch.sin(t[:, None].float() * f)], dim=1)
    assert torch.equal(emb, want)


def test_groupnorm_vs_manual():
    x = torch.randn(2, 64, 5, 7)
    gn = torch.nn.GroupNorm(32, 64, eps=1e-5)
    with torch.no_grad():
        gn.weight.normal_()
        gn.bias.normal_()
    xg = x.view(2, 32, -1)
    mean = xg.mean(-1, keepdim=True)
    var = xg.var(-1, unbiased=False, keepdim=True)
    want = ((xg - mean) / torch.sqrt(var + 1e-5)).view_as(x) * gn.weight[None, :, None, None] + gn.bias[None, :, None, None]
    torch.testing.assert_close(gn(x), want, rtol=1e-5, atol=1e-5)


def test_geglu_and_attention_closed_forms():
    torch.manual_seed(0)
    g = B.GEGLU(16, 24)
    x = torch.randn(3, 5, 16)
    y = g.proj(x)
    want = y[..., :24] * (0.5 * y[..., 24:] * (1 + torch.erf(y[..., 24:] / math.sqrt(2))))
    torch.testing.assert_close(g(x), want, rtol=1e-6, atol=1e-6)

    attn = B.Attention(query_dim=128, cross_attention_dim=32, heads=2, dim_head=64)
    h = torch.randn(2, 10, 128)
    ctx = torch.randn(2, 7, 32)
    q = attn.to_q(h).view(2, 10, 2, 64).transpose(1, 2)
    k = attn.to_k(ctx).view(2, 7, 2, 64).transpose(1, 2)
    v = attn.to_v(ctx).view(2, 7, 2, 64).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)  # scale 1/sqrt(64) = 0.125
    want = attn.to_out[0](o.transpose(1, 2).reshape(2, 10, 128))
    torch.testing.assert_close(attn(h, encoder_hidden_states=ctx), want, rtol=1e-5, atol=1e-5)
    assert attn.scale == 0.125


def test_resnet_block_closed_form():
    torch.manual_seed(0)
    r = B.ResnetBlock2D(in_channels=64, out_channels=128, temb_channels=32, eps=1e-5, groups=32)
    x, temb = torch.randn(2, 64, 4, 6), torch.randn(2, 32)
    h = r.conv1(F.silu(r.norm1(x)))
    h = h + r.time_emb_proj(F.silu(temb))[:, :, None, None]
    h = r.conv2(F.silu(r.norm2(h)))
    want = r.conv_shortcut(x) + h
    torch.testing.assert_close(r(x, temb), want)
    assert r.conv_shortcut.kernel_size == (1, 1)
    assert B.ResnetBlock2D(in_channels=64, out_channels=64, temb_channels=32).conv_shortcut is None


def test_ddim_schedule_and_step_closed_form():
    ac = scaled_linear_alphas_cumprod()
    assert abs(float(ac[0]) - 0.99915) < 1e-6
    s = OracleDDIMScheduler()
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(981, 0, -20))
    s.set_timesteps(10)
    assert s.timesteps.tolist() == list(range(901, 0, -100))
    x, eps = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    out = s.step(eps, 901, x, return_dict=False)[0]
    a_t, a_p = ac[901], ac[801]
    want = a_p.sqrt() * (x - (1 - a_t).sqrt() * eps) / a_t.sqrt() + (1 - a_p).sqrt() * eps
    torch.testing.assert_close(out, want)
    last = s.step(eps, 1, x, return_dict=False)[0]  # t_prev < 0 -> alpha_cumprod[0] (set_alpha_to_one=False)
    want = ac[0].sqrt() * (x - (1 - ac[1]).sqrt() * eps) / ac[1].sqrt() + (1 - ac[0]).sqrt() * eps
    torch.testing.assert_close(last, want)
    assert s.init_noise_sigma == 1.0 and s.order == 1
    assert s.scale_model_input(x, 5) is x


def test_ddpm_add_noise():
    x0, n = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8)
    t = torch.tensor([0, 500, 999])
    ac = scaled_linear_alphas_cumprod()
    want = ac[t].sqrt()[:, None, None, None] * x0 + (1 - ac[t]).sqrt()[:, None, None, None] * n
    torch.testing.assert_close(ddpm_add_noise(x0, n, t), want)


def test_cfg_combine_and_conditioning_layout():
    e = torch.randn(4, 4, 2, 2)
    torch.testing.assert_close(cfg_combine(e, 2.0), e[:2] + 2.0 * (e[2:] - e[:2]))
    cond = prepare_conditioning(s_img_proj_f=torch.randn(1, 8, 16), pred_t_img_embed=torch.randn(1, 1, 16),
                                st_pose_f=torch.randn(1, 6, 4, 8), masked_latents=torch.randn(1, 4, 4, 8), height=32,
                                width=64, num_images_per_prompt=3, guidance_scale=2.0)
    assert cond["feature_f"].shape == (6, 9, 16) and cond["prior_embed"].shape == (6, 1, 16)
    assert cond["feature_f"][:3].abs().sum() == 0 and cond["prior_embed"][:3].abs().sum() == 0  # uncond half first
    assert cond["mask"].shape == (6, 1, 4, 8)
    assert torch.all(cond["mask"][..., :4] == 1) and torch.all(cond["mask"][..., 4:] == 0)
    assert cond["pose_cond"].shape == (6, 6, 4, 8)  # pose is NOT dropped for the unconditional half
    assert torch.equal(cond["pose_cond"][0], cond["pose_cond"][5])


def test_oracle_matches_reference_unet_golden():
    """tests/golden/ref_unet_tiny.pt was produced by the reference's own UNet class (over the shim)."""
    g = torch.load(GOLD / "ref_unet_tiny.pt")
    m = make_unet(UNetConfig.tiny(), seed=g["seed"])
    i = g["inputs"]
    out = m(i["sample"], g["timestep"], i["encoder_hidden_states"], class_labels=i["class_labels"],
            my_pose_cond=i["my_pose_cond"])[0]
    assert torch.equal(out, g["out"])
    out_v = m(i["sample"], g["timestep_vec"], i["encoder_hidden_states"], class_labels=i["class_labels"],
              my_pose_cond=i["my_pose_cond"])[0]
    assert torch.equal(out_v, g["out_vec"])
    assert 0.05 < float(out.std()) < 20


def test_oracle_matches_reference_pipeline_golden():
    """tests/golden/ref_pipeline_tiny.pt was produced by the reference's own pipeline __call__ (fp16 loop tensors)."""
    g = torch.load(GOLD / "ref_pipeline_tiny.pt")
    cfg = UNetConfig.tiny()
    m = make_unet(cfg, seed=g["seed"]).half()
    pin = g["inputs"]
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"],
                                height=pin["height"], width=pin["width"], num_images_per_prompt=g["n"],
                                guidance_scale=g["guidance_scale"], dtype=torch.float16)
    lat = denoise_loop(m, OracleDDIMScheduler(), latents=pin["latents"], cond=cond, num_inference_steps=g["steps"],
                       guidance_scale=g["guidance_scale"], dtype=torch.float16)
    assert lat.dtype == torch.float16
    assert torch.equal(lat, g["latents"])


def test_oracle_matches_reference_stage3_and_demo_goldens():
    """tests/golden/ref_stage3_tiny.pt / ref_demo_tiny.pt were produced by the reference's own Stage3_RefinedPipeline
    and PCDMsPipeline `__call__` (tools/make_golden.py); the oracle loops must reproduce them bit for bit."""
    from dataclasses import replace
    from oracle.pipeline import denoise_loop_pcdms, denoise_loop_stage3
    g = torch.load(GOLD / "ref_stage3_tiny.pt")
    u3 = make_unet(UNetConfig.tiny(in_channels=8, stage2=False), seed=g["seed"])
    got = denoise_loop_stage3(u3, OracleDDIMScheduler(), dtype=torch.float16, unet_dtype=torch.float32, **g["inputs"])
    assert torch.equal(got.float(), g["latents"].float())
    g = torch.load(GOLD / "ref_demo_tiny.pt")
    ud = make_unet(replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True), seed=g["seed"]).half()
    got = denoise_loop_pcdms(ud, OracleDDIMScheduler(), dtype=torch.float16, **g["inputs"])
    assert torch.equal(got, g["latents"])


def test_oracle_image_proj_matches_reference_golden():
    """tests/golden/ref_image_proj.pt: state dict, input and output of the reference's own ImageProjModel_p."""
    from oracle.frontend import ImageProjModel_p
    g = torch.load(GOLD / "ref_image_proj.pt")
    m = ImageProjModel_p(128, 64, 96).eval()
    m.load_state_dict(g["state_dict"])
    with torch.no_grad():
        assert torch.equal(m(g["x"]), g["y"])
