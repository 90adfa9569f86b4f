"""Short-K residual GEMMs of the transformer blocks (to_out / proj_out / ff.net.2 at the three UNet levels) for
`ncu --set full`: two warm launches + one to profile, each."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
dev, dt = "cuda", torch.bfloat16
for M, N, K in ((32768, 320, 320), (8192, 640, 640), (2048, 1280, 1280)):
    a = torch.randn(M, K, device=dev, dtype=dt)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev, dtype=dt)
    for _ in range(3):
        ops.gemm(a, w, bias=b, residual=r)
    torch.cuda.synchronize()
print("ok")
