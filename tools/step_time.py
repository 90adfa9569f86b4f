"""ms per CFG UNet step of the bench workload (BASELINE config 2: 8 images, 32x64 latents, 258 tokens, bf16) through the
fused engine — the A/B number for a kernel change (`PCDM_B200_LIB=<other build> python tools/step_time.py`)."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200.pipeline import B200Stage2InpaintPipeline
from pcdms_b200.scheduler import B200DDIMScheduler
from pcdms_b200.unet import B200UNet2DConditionModel

dev, dt, n, h, w, steps = "cuda", torch.bfloat16, 8, 32, 64, 50
g = lambda s: torch.Generator().manual_seed(s)
unet = B200UNet2DConditionModel(dtype=dt, device=dev, in_channels=9, class_embed_type="projection",
                                projection_class_embeddings_input_dim=1024)
unet.load_state_dict(unet.synthetic_state_dict(seed=0))
pipe = B200Stage2InpaintPipeline(vae=None, unet=unet, scheduler=B200DDIMScheduler())
pipe(height=h * 8, width=w * 8, num_inference_steps=steps, guidance_scale=2.0, num_images_per_prompt=n,
     latents=torch.randn(n, 4, h, w, generator=g(1)), s_img_proj_f=torch.randn(1, 257, 1024, generator=g(2)),
     st_pose_f=0.1 * torch.randn(1, 320, h, w, generator=g(3)), pred_t_img_embed=torch.randn(1, 1, 1024, generator=g(4)),
     masked_latents=torch.randn(1, 4, h, w, generator=g(5)), output_type="latent")
st = next(iter(pipe._graphs.values()))
best = 1e9
for rep in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.replay_fused(st)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
print(f"unet_step_ms {best:.4f}")
