"""Time the other BASELINE configurations (parity-test cases, not the headline bench line) through the public pipeline
classes with synthetic weights: resident-input images/s and ms per CFG UNet step.

  cfg2  stage-2, 8 images, 256x256 (32x64 latents), 258 tokens, 50 DDIM        (the bench.py workload, for reference)
  cfg2_95tokens  the same with BASELINE's "77+18" = 95 conditioning tokens      (secondary figure)
  cfg3  stage-2, 4 images, 512x512 (64x128 latents), 258 tokens, 50 DDIM
  cfg5  stage-3 refiner, 8 images, 512x512 (64x64 latents), 257 tokens, 30 DDIM
  drv   the batch-test driver's defaults (stage2_batchtest_inpaint_model.py:199-211,256-262): 4 images, 512x512,
        20 UniPC steps, fp16, image in -> image out through the B200 VAE
usage: python tools/bench_configs.py [out.json]
"""
import json
import sys

sys.path.insert(0, ".")
import torch

from pcdms_b200.pipeline import B200Stage2InpaintPipeline, B200Stage3RefinedPipeline
from pcdms_b200.scheduler import B200DDIMScheduler, B200UniPCMultistepScheduler
from pcdms_b200.unet import B200UNet2DConditionModel
from pcdms_b200.vae import B200AutoencoderKL

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_configs.json"
dev = "cuda"
SD21 = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1,
            timestep_spacing="leading")
TFLOP_ROW = {(32, 64): 0.387, (64, 128): 1.876, (64, 64): 0.822}   # SURVEY.md §8d, per UNet batch row


def timed(fn, k=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def g(seed):
    return torch.Generator().manual_seed(seed)


def stage2(n, h, w, steps, dt, sched, vae=None, tokens=257):
    unet = B200UNet2DConditionModel(dtype=dt, device=dev, in_channels=9, class_embed_type="projection",
                                    projection_class_embeddings_input_dim=1024)
    unet.load_state_dict(unet.synthetic_state_dict(seed=0))
    pipe = B200Stage2InpaintPipeline(vae=vae, unet=unet, scheduler=sched)
    kw = dict(height=h * 8, width=w * 8, num_inference_steps=steps, guidance_scale=2.0, num_images_per_prompt=n,
              latents=torch.randn(n, 4, h, w, generator=g(1)), s_img_proj_f=torch.randn(1, tokens, 1024, generator=g(2)),
              st_pose_f=0.1 * torch.randn(1, 320, h, w, generator=g(3)),
              pred_t_img_embed=torch.randn(1, 1, 1024, generator=g(4)))
    if vae is None:
        kw.update(masked_latents=torch.randn(1, 4, h, w, generator=g(5)), output_type="latent")
    else:
        kw.update(vae_image=torch.rand(1, 3, h * 8, w * 8, generator=g(5)) * 2 - 1, output_type="pt")
    out = pipe(**kw).images
    assert torch.isfinite(out.float()).all()
    call_ms = timed(lambda: pipe(**kw))
    st = next(iter(pipe._graphs.values()))
    loop_ms = timed(lambda: pipe.replay_fused(st))
    return call_ms, loop_ms


def stage3(n, h, w, steps, dt):
    unet = B200UNet2DConditionModel(dtype=dt, device=dev, in_channels=8)
    unet.load_state_dict(unet.synthetic_state_dict(seed=0))
    pipe = B200Stage3RefinedPipeline(vae=None, unet=unet, scheduler=B200DDIMScheduler())
    kw = dict(height=h * 8, width=w * 8, num_inference_steps=steps, guidance_scale=2.0, num_images_per_prompt=n,
              latents=torch.randn(n, 4, h, w, generator=g(1)), s_img_proj_f=torch.randn(1, 257, 1024, generator=g(2)),
              gen_t_img_latents=torch.randn(1, 4, h, w, generator=g(3)), output_type="latent")
    out = pipe(**kw).images
    assert torch.isfinite(out.float()).all()
    call_ms = timed(lambda: pipe(**kw))
    st = next(iter(pipe._graphs.values()))
    loop_ms = timed(lambda: pipe.replay_fused(st))
    return call_ms, loop_ms


res = {}


def record(name, n, h, w, steps, call_ms, loop_ms, note):
    tf = TFLOP_ROW[(h, w)] * 2 * n
    res[name] = {"images": n, "latent": [h, w], "steps": steps, "call_ms": call_ms, "loop_ms": loop_ms,
                 "unet_step_ms": loop_ms / steps, "images_per_s_resident": n / (loop_ms * 1e-3),
                 "images_per_s_call": n / (call_ms * 1e-3), "tflops_per_unet_fwd": tf,
                 "achieved_tflops": tf / (loop_ms / steps * 1e-3), "note": note}
    print(name, json.dumps(res[name]), flush=True)


c, l = stage2(8, 32, 64, 50, torch.bfloat16, B200DDIMScheduler())
record("cfg2", 8, 32, 64, 50, c, l, "stage-2 b8 256x256 DDIM-50 bf16 (bench.py workload)")
torch.cuda.empty_cache()
# BASELINE's "77+18-token" synthetic conditioning (S_kv = 95, the secondary figure next to the
# reference-faithful 258): 94 image tokens + the appended predicted-embedding token
c, l = stage2(8, 32, 64, 50, torch.bfloat16, B200DDIMScheduler(), tokens=94)
TFLOP_ROW[(32, 64)] = 5.998 / 16
record("cfg2_95tokens", 8, 32, 64, 50, c, l, "stage-2 b8 256x256 DDIM-50 bf16 with 95 conditioning tokens (5.998 TFLOP / fwd)")
TFLOP_ROW[(32, 64)] = 0.387
torch.cuda.empty_cache()
c, l = stage2(4, 64, 128, 50, torch.bfloat16, B200DDIMScheduler())
record("cfg3", 4, 64, 128, 50, c, l, "stage-2 b4 512x512 DDIM-50 bf16")
torch.cuda.empty_cache()
c, l = stage3(8, 64, 64, 30, torch.bfloat16)
record("cfg5", 8, 64, 64, 30, c, l, "stage-3 refiner b8 512x512 DDIM-30 bf16")
torch.cuda.empty_cache()
vae = B200AutoencoderKL(dtype=torch.float16, device=dev)
vae.load_state_dict(vae.synthetic_state_dict(seed=1))
c, l = stage2(4, 64, 128, 20, torch.float16, B200UniPCMultistepScheduler.from_config(SD21), vae=vae)
record("driver_default", 4, 64, 128, 20, c, l,
       "stage2_batchtest defaults: 4 images 512x512, UniPC-20, fp16, VAE encode + decode inside call_ms")
json.dump(res, open(out_path, "w"), indent=1)
