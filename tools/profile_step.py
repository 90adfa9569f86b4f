"""Run a few eager UNet evaluations at BASELINE config 2 (B=16, 32x64, 258 tokens, bf16) for ncu launch lists."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200.unet import B200UNet2DConditionModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"
dt = torch.bfloat16
m = B200UNet2DConditionModel(dtype=dt, device=dev, in_channels=9, class_embed_type="projection", projection_class_embeddings_input_dim=1024)
m.load_state_dict(m.synthetic_state_dict(0))
B, h, w = 16, 32, 64
x9 = torch.randn(B, h, w, 64, device=dev).to(dt)
t = torch.tensor([981.0], device=dev)
ctx = torch.randn(B, 258, 1024, device=dev).to(dt)
cls = torch.randn(B, 1024, device=dev).to(dt)
pose = (0.1 * torch.randn(B, h, w, 320, device=dev)).to(dt)
kv = m.context_kv(ctx)
for _ in range(n):
    out = m.forward_nhwc(x9, t, kv, cls, pose)
torch.cuda.synchronize()
print("done", float(out.float().abs().mean()))
