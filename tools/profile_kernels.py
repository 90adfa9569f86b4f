"""A handful of representative launches for `ncu --set full` (one warm-up + one profiled launch each)."""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch
from pcdms_b200 import ops, lib
dev = "cuda"; dt = torch.bfloat16
L = lib.load()
x320 = torch.randn(16, 32, 64, 320, device=dev, dtype=dt); w320 = (torch.randn(320, 2880, device=dev) / 54).to(dt)
b320 = torch.randn(320, device=dev); r320 = torch.randn(16, 32, 64, 320, device=dev, dtype=dt)
x1280 = torch.randn(16, 8, 16, 1280, device=dev, dtype=dt); w1280 = (torch.randn(1280, 11520, device=dev) / 107).to(dt)
a320 = torch.randn(32768, 320, device=dev, dtype=dt); wg = (torch.randn(2560, 320, device=dev) / 18).to(dt); bg = torch.randn(2560, device=dev)
qkv = torch.randn(16 * 2048, 960, device=dev, dtype=dt)
def run():
    ops.conv3x3(x320, w320, bias=b320, residual=r320)                  # 0: conv 320->320 @32x64 (auto: BN160, CTA pairs)
    ops.conv3x3(x1280, w1280)                                          # 1: conv 1280->1280 @8x16 (auto: BN160, CTA pairs)
    ops.gemm(a320, wg, bias=bg, geglu=True)                            # 2: GEGLU GEMM (auto: BN256, single CTA)
    L.pcdm_set_gemm_cta_group(1)
    ops.conv3x3(x320, w320, bias=b320, residual=r320, bn=160)          # 3: conv 320->320, single-CTA BN160 tiles
    L.pcdm_set_gemm_cta_group(0)
    ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], 16, 5)  # 4: self-attention 2048x2048, 5 heads
    torch.cuda.synchronize()
run(); run()
print("ok")
