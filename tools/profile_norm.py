"""Norm-kernel launches for `ncu --set full`: GroupNorm (single-pass and statistics + apply) and LayerNorm at the bench
shapes, with a same-size device copy as the bandwidth yardstick."""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch
from pcdms_b200 import ops, lib
dev = "cuda"; dt = torch.bfloat16
L = lib.load()
x320 = torch.randn(16, 32, 64, 320, device=dev, dtype=dt); g320 = torch.randn(320, device=dev); b320 = torch.randn(320, device=dev)
x640 = torch.randn(16, 32, 64, 640, device=dev, dtype=dt); g960 = torch.randn(960, device=dev); b960 = torch.randn(960, device=dev)
x1280 = torch.randn(16, 8, 16, 1280, device=dev, dtype=dt); g1280 = torch.randn(1280, device=dev); b1280 = torch.randn(1280, device=dev)
r320 = x320.view(-1, 320); y320 = torch.empty_like(x320)
def run():
    y320.copy_(x320)                                                   # 0: 21 MB -> 21 MB device copy
    ops.groupnorm(x320, g320, b320, 1e-5, silu=True)                   # 1: fused, L0 C320
    ops.groupnorm(x640, g960, b960, 1e-5, x2=x320, silu=True)          # 2: fused, L0 C960 cat
    ops.groupnorm(x1280, g1280, b1280, 1e-5, silu=True)                # 3: fused, L2 C1280
    L.pcdm_set_groupnorm_two_pass(1)
    ops.groupnorm(x320, g320, b320, 1e-5, silu=True)                   # 4,5: stats + apply, L0 C320
    L.pcdm_set_groupnorm_two_pass(0)
    ops.layernorm(r320, g320, b320)                                    # 6: LN (32768, 320)
    torch.cuda.synchronize()
run(); run()
print("ok")
