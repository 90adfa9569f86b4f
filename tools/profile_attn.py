"""Self-attention 2048x2048 (B16, 5 heads, bf16) twice, for `ncu --set full -k regex:attention`."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
qkv = torch.randn(16 * 2048, 960, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], 16, 5)
torch.cuda.synchronize()
print("ok")
