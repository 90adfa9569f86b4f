"""Self-attention 2048x2048 (B16, 5 heads, bf16) on the experiment build's v2 kernel, for `ncu --set full -k regex:attention2`."""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401
import ctypes as C
import torch
from pcdms_b200 import ops, lib
L = lib.load()
L.pcdm_set_attention_v2(C.c_int(1))
qkv = torch.randn(16 * 2048, 960, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], 16, 5)
torch.cuda.synchronize()
print("ok")
