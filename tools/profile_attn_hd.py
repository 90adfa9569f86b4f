"""The head_dim-128 variant of the flash-attention kernel at CLIP ViT-H/14's shape (batch 2, 16 heads, 257 tokens,
80-wide heads zero-padded to 128) and at a long sequence (batch 2, 16 heads, 2048 tokens), for `ncu --set full`."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
dev, dt = "cuda", torch.float16
for B, heads, S in ((2, 16, 257), (2, 16, 2048)):
    C = heads * 128
    qkv = torch.randn(B * S, 3 * C, device=dev, dtype=dt)
    qkv.view(B * S, 3 * heads, 128)[:, :, 80:] = 0     # the padding columns really are zero in the encoder
    for _ in range(3):
        ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads, scale=80 ** -0.5, head_dim=128)
    torch.cuda.synchronize()
print("ok")
