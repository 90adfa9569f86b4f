"""Time the widened rows either side of the stage-2 loop at full size with synthetic weights:

  prior   stage-1 prior (20 blocks, 2048 wide, 6 tokens), UnCLIP sampling loop through B200Stage1PriorPipeline:
          ms per step, and the weight stream it amounts to (the prior is weight-bandwidth work: ~2.05 GB of 16-bit
          weights against 6 or 12 activation rows per step) against the measured HBM peak
  clip    CLIP ViT-H/14 image encoder (32 layers, 1280 wide, 257 tokens), batch 2 (source + target image)
  dinov2  DINOv2-giant (40 layers, 1536 wide, 257 tokens), batch 1

usage: python tools/bench_stage1.py [out.json]
"""
import json
import sys

sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch

from pcdms_b200.clip import B200CLIPVisionModelWithProjection
from pcdms_b200.dinov2 import B200Dinov2Model
from pcdms_b200.prior import B200Stage1PriorPipeline, B200Stage1PriorTransformer

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_stage1.json"
dev = "cuda"
try:
    HBM = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs")
except Exception:
    HBM = None


def timed(fn, k=5, warm=3):
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


res = {"hbm_peak_gbps": HBM}
dt = torch.float16
import ctypes as _C
import os
from pcdms_b200 import lib as _lib
_L = _lib.load()
_L.pcdm_set_skinny_gemm(_C.c_int(int(os.environ.get("PCDM_SKINNY", "1"))))          # A/B hooks
_L.pcdm_set_attention_small(_C.c_int(int(os.environ.get("PCDM_ATT_SMALL", "1"))))
ONLY_PRIOR = os.environ.get("PCDM_ONLY_PRIOR") == "1"

prior = B200Stage1PriorTransformer(dtype=dt, device=dev, num_embeddings=2, embedding_dim=1024)
prior.load_state_dict(prior.synthetic_state_dict(seed=0))
weight_bytes = sum(v.numel() * v.element_size() for k, v in prior._w.items()
                   if k.split(".")[0].isdigit() and v.dtype == dt)        # the 20 blocks' matrices, read every step


class _Zero:
    config = type("c", (), {"image_size": 8})
    dtype = dt

    def __call__(self, x):
        return {"image_embeds": torch.zeros(1, 1024, device=x.device)}


pipe = B200Stage1PriorPipeline(prior=prior, image_encoder=_Zero())
g = torch.Generator(device=dev).manual_seed(0)
for name, n, guidance in (("prior_n1", 1, 0.0), ("prior_n1_cfg", 1, 2.0), ("prior_n4", 4, 0.0)):
    steps = 25
    kw = dict(s_embed=torch.randn(n, 1, 1024, device=dev, generator=g), s_pose=torch.rand(n, 1, 36, device=dev, generator=g),
              t_pose=torch.rand(n, 1, 36, device=dev, generator=g), num_inference_steps=steps, guidance_scale=guidance,
              generator=g)
    out = pipe(**kw)
    assert torch.isfinite(out[0]).all()
    call_ms = timed(lambda: pipe(**kw))
    st = pipe._graphs[(n, 1024, guidance > 1.0, dt)]
    loop_ms = timed(lambda: pipe.replay_fused(st))
    step_ms = loop_ms / steps
    res[name] = {"embeddings": n, "rows": (2 if guidance > 1 else 1) * 6 * n, "steps": steps, "call_ms": call_ms,
                 "loop_ms": loop_ms, "step_ms": step_ms, "launches_per_step": st.launches_per_step,
                 "block_weight_bytes": weight_bytes, "weight_stream_gbps": weight_bytes / (step_ms * 1e-3) / 1e9,
                 "frac_of_hbm_peak": (weight_bytes / (step_ms * 1e-3) / 1e9 / HBM) if HBM else None}
    print(name, json.dumps(res[name]), flush=True)
del pipe, prior
torch.cuda.empty_cache()
if ONLY_PRIOR:
    json.dump(res, open(out_path, "w"), indent=1)
    sys.exit(0)

clip = B200CLIPVisionModelWithProjection(dtype=dt, device=dev)
clip.load_state_dict(clip.synthetic_state_dict(seed=1))
x = torch.randn(2, 3, 224, 224, device=dev)
assert torch.isfinite(clip(x).image_embeds.float()).all()
ms = timed(lambda: clip(x))
flops = 2 * 2 * 257 * 32 * (4 * 1280 * 1280 + 2 * 1280 * 5120) + 32 * 2 * 4 * 16 * 257 * 257 * 80
res["clip_vit_h_b2"] = {"images": 2, "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12}
print("clip", json.dumps(res["clip_vit_h_b2"]), flush=True)
del clip
torch.cuda.empty_cache()

dino = B200Dinov2Model(dtype=dt, device=dev)
dino.load_state_dict(dino.synthetic_state_dict(seed=2))
x = torch.randn(1, 3, 224, 224, device=dev)
assert torch.isfinite(dino(x).last_hidden_state.float()).all()
ms = timed(lambda: dino(x))
res["dinov2_giant_b1"] = {"images": 1, "ms": ms}
print("dinov2", json.dumps(res["dinov2_giant_b1"]), flush=True)
json.dump(res, open(out_path, "w"), indent=1)
