"""What paces the v2 attention kernel: self-attention 2048x2048 (B16, 5 heads, bf16) with parts of the softmax switched
off through the experiment build's pcdm_set_attention_debug (results are wrong under a non-zero mask; timing only)."""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401
import ctypes as C
import torch
from pcdms_b200 import ops, lib
from tools.dev_attn2 import graph_us  # noqa: E402  (prints that tool's table first)

L = lib.load()
B, heads, S = 16, 5, 2048
q = torch.randn(B * S, 960, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, 320, device="cuda", dtype=torch.bfloat16)
names = {0: "full", 1: "no ex2", 2: "one S chunk read", 4: "no row max", 8: "one P chunk stored", 16: "no pack",
         17: "no ex2, no pack", 21: "no ex2, no pack, no max", 3: "no ex2 + one S chunk", 7: "no ex2, S chunk, max",
         31: "everything off"}
for v2 in (1, 0):
    L.pcdm_set_attention_v2(C.c_int(v2))
    for m, name in names.items():
        if v2 == 0 and m:
            continue
        L.pcdm_set_attention_debug(C.c_int(m))
        us = graph_us(lambda: ops.attention(q[:, :320], q[:, 320:640], q[:, 640:], B, heads, out=out))
        print(f"v2={v2} mask {m:2d} ({name:24s}): {us:7.2f} us", flush=True)
L.pcdm_set_attention_debug(C.c_int(0))
L.pcdm_set_attention_v2(C.c_int(1))
