"""Generate tests/golden/*.pt by running the REFERENCE's own classes (unmodified, from /root/reference) on top of
oracle/diffusers_shim.  Run in the build container only (the GPU box has no /root/reference):

    python tools/make_golden.py

Fixtures (tiny topology-preserving config so the files stay small):
  ref_unet_tiny.pt      inputs + output of Stage2_InapintUNet2DConditionModel.forward (fp32)
  ref_pipeline_tiny.pt  inputs + final latents of Stage2_InpaintDiffusionPipeline.__call__ (fp16 loop tensors, as
                        the reference hard-codes them; DDIM, 4 steps, guidance 2.0, 2 images per prompt)
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


if __name__ == "__main__":
    main()
