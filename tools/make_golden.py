"""Generate tests/golden/*.pt by running the REFERENCE's own classes (unmodified, from /root/reference) on top of
oracle/diffusers_shim.  Run in the build container only (the GPU box has no /root/reference):

    python tools/make_golden.py

Fixtures (tiny topology-preserving config so the files stay small):
  ref_unet_tiny.pt      inputs + output of Stage2_InapintUNet2DConditionModel.forward (fp32)
  ref_pipeline_tiny.pt  inputs + final latents of Stage2_InpaintDiffusionPipeline.__call__ (fp16 loop tensors, as
                        the reference hard-codes them; DDIM, 4 steps, guidance 2.0, 2 images per prompt)
  ref_stage3_tiny.pt    inputs + final latents of Stage3_RefinedPipeline.__call__ (fp16 loop tensors, fp32 UNet; DDIM)
  ref_demo_tiny.pt      inputs + final latents of PCDMsPipeline.__call__ (the pcdms_demo.ipynb driver; fp16; DDIM)
  ref_simple_tiny.pt    inputs + final latents of Simple_Stage2_InpaintDiffusionPipeline.__call__ (fp16; DDIM, 3 steps)
  ref_prior_tiny.pt     inputs + outputs of Stage1_PriorTransformer.forward (plain and test_flag) and of
                        Stage1_PriorPipeline.__call__ (fp32, 4 UnCLIP steps, guidance 0 as the batch-test driver) with
                        the variance noise the global generator produced
  ref_image_proj.pt     state dict + input + output of the reference's ImageProjModel_p class
                        (stage2_batchtest_inpaint_model.py:48-66), at a reduced width
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import reference_shim as rs  # noqa: E402
from oracle.factory import make_inputs, make_unet, make_unet_inputs  # noqa: E402
from oracle.unet import UNetConfig  # noqa: E402

GOLD = ROOT / "tests" / "golden"
GOLD.mkdir(parents=True, exist_ok=True)


def main():
    cfg = UNetConfig.tiny()
    oracle_unet = make_unet(cfg, seed=0)
    sd = oracle_unet.state_dict()

    ref_unet = rs.build_reference_unet(cfg)
    missing, unexpected = ref_unet.load_state_dict(sd, strict=True)
    inp = make_unet_inputs(cfg, batch=2, h=16, w=32, s_kv=9)
    with torch.no_grad():
        out = ref_unet(inp["sample"], 981, inp["encoder_hidden_states"], class_labels=inp["class_labels"],
                       my_pose_cond=inp["my_pose_cond"], return_dict=False)[0]
        out_t = ref_unet(inp["sample"], torch.tensor([21, 501]), inp["encoder_hidden_states"],
                         class_labels=inp["class_labels"], my_pose_cond=inp["my_pose_cond"]).sample
    torch.save({"cfg": "tiny", "seed": 0, "inputs": inp, "timestep": 981, "out": out,
                "timestep_vec": torch.tensor([21, 501]), "out_vec": out_t}, GOLD / "ref_unet_tiny.pt")
    print("ref_unet_tiny", out.shape, float(out.std()))

    pin = make_inputs(cfg, n=2, h=16, w=32, s_kv=9)
    latents = rs.run_reference_pipeline(cfg, sd, pin, num_inference_steps=4, guidance_scale=2.0,
                                        num_images_per_prompt=2)
    torch.save({"cfg": "tiny", "seed": 0, "inputs": pin, "steps": 4, "guidance_scale": 2.0, "n": 2,
                "latents": latents}, GOLD / "ref_pipeline_tiny.pt")
    print("ref_pipeline_tiny", latents.shape, latents.dtype, float(latents.float().std()))

    from dataclasses import replace
    from oracle.schedulers import OracleDDIMScheduler
    g = torch.Generator().manual_seed(31)
    cfg3 = UNetConfig.tiny(in_channels=8, stage2=False)
    u3 = make_unet(cfg3, seed=7)
    kw3 = dict(latents=torch.randn(1, 4, 8, 8, generator=g), gen_t_img_latents=torch.randn(1, 4, 8, 8, generator=g),
               s_img_proj_f=torch.randn(1, 5, cfg3.cross_attention_dim, generator=g), num_inference_steps=4,
               guidance_scale=2.0)
    out3 = rs.run_reference_stage3_pipeline(cfg3, u3, scheduler=OracleDDIMScheduler(), **kw3)
    torch.save({"seed": 7, "inputs": kw3, "latents": out3}, GOLD / "ref_stage3_tiny.pt")
    print("ref_stage3_tiny", out3.shape, out3.dtype)

    cfgd = replace(UNetConfig.tiny(in_channels=9, stage2=False), use_pose_cond=True)
    ud = make_unet(cfgd, seed=9).half()
    h, w = 8, 16
    kwd = dict(latents=torch.randn(1, 4, h, w, generator=g), simg_mask_latents=torch.randn(1, 4, h, w, generator=g),
               mask=torch.cat([torch.ones(1, 1, h, w // 2), torch.zeros(1, 1, h, w // 2)], dim=3),
               cond_pose=0.1 * torch.randn(1, cfgd.block_out_channels[0], h, w, generator=g),
               prompt_embeds=torch.randn(1, 7, cfgd.cross_attention_dim, generator=g),
               negative_prompt_embeds=0.3 * torch.randn(1, 7, cfgd.cross_attention_dim, generator=g),
               num_inference_steps=3, guidance_scale=2.0)
    outd = rs.run_reference_demo_pipeline(cfgd, ud, **kwd)
    torch.save({"seed": 9, "inputs": kwd, "latents": outd}, GOLD / "ref_demo_tiny.pt")
    print("ref_demo_tiny", outd.shape, outd.dtype)

    us = make_unet(cfgd, seed=12).half()
    kws = dict(latents=torch.randn(2, 4, h, w, generator=g), masked_latents=torch.randn(1, 4, h, w, generator=g).half(),
               st_pose_f=(0.1 * torch.randn(1, cfgd.block_out_channels[0], h, w, generator=g)).half(),
               s_img_proj_f=torch.randn(1, 7, cfgd.cross_attention_dim, generator=g).half(), height=h * 8, width=w * 8,
               num_inference_steps=3, guidance_scale=2.0, num_images_per_prompt=2)
    outs = rs.run_reference_simple_stage2_pipeline(cfgd, us, **kws)
    torch.save({"seed": 12, "inputs": kws, "latents": outs}, GOLD / "ref_simple_tiny.pt")
    print("ref_simple_tiny", outs.shape, outs.dtype)

    from oracle.prior import TINY, make_prior, make_prior_inputs
    op = make_prior(seed=13, **TINY)
    rp = rs.build_reference_prior(**TINY)
    rp.load_state_dict(op.state_dict(), strict=True)
    pi = make_prior_inputs(n=1, seed=17, steps=4)
    x, x2 = pi["latents"][:, None], torch.cat([pi["latents"], pi["latents"]])[:, None]
    e2 = torch.cat([torch.zeros_like(pi["s_embed"]), pi["s_embed"]])
    with torch.no_grad():
        fwd = rp(x, 500, pi["s_embed"], pi["s_pose"], pi["t_pose"]).predicted_image_embedding
        fwd_flag = rp(x2, torch.tensor(37), e2, pi["s_pose"], pi["t_pose"], test_flag=True).predicted_image_embedding
    # the reference draws each step's variance noise from the global generator: record what it drew (same seed, same
    # shapes, same order) so that the fixture carries it
    torch.manual_seed(23)
    drawn = torch.stack([torch.randn(1, 1024) for _ in range(3)] + [torch.zeros(1, 1024)])   # last step adds none
    torch.manual_seed(23)
    emb, zero = rs.run_reference_prior_pipeline(rp, s_embed=pi["s_embed"], s_pose=pi["s_pose"], t_pose=pi["t_pose"],
                                                latents=pi["latents"], num_inference_steps=4, guidance_scale=0,
                                                zero_embed=torch.zeros(1, 1024))
    torch.save({"seed": 13, "input_seed": 17, "forward": fwd, "forward_test_flag": fwd_flag, "steps": 4,
                "variance_noise": drawn, "image_embeds": emb}, GOLD / "ref_prior_tiny.pt")
    print("ref_prior_tiny", emb.shape, float(emb.std()))

    import ast
    path = "/root/reference/stage2_batchtest_inpaint_model.py"
    node = next(n for n in ast.parse(open(path).read()).body
                if isinstance(n, ast.ClassDef) and n.name == "ImageProjModel_p")
    ns = {"torch": torch, "nn": torch.nn}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    torch.manual_seed(3)
    proj = ns["ImageProjModel_p"](in_dim=128, hidden_dim=64, out_dim=96).eval()
    x = torch.randn(1, 9, 128, generator=g)
    with torch.no_grad():
        y = proj(x)
    torch.save({"state_dict": proj.state_dict(), "x": x, "y": y}, GOLD / "ref_image_proj.pt")
    print("ref_image_proj", y.shape)


if __name__ == "__main__":
    main()
