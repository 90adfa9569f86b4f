"""The stage-1 prior's weight-stream GEMMs for `ncu --set full`: ff.net.0 (6 rows x [8192, 2048]) and ff.net.2
(6 rows x [2048, 8192]) through pcdm_gemm -> skinny_gemm_kernel; two warm launches + one to profile, each."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
dev, dt = "cuda", torch.float16
for M, N, K, kw in ((6, 8192, 2048, dict(gelu=True)), (6, 2048, 8192, dict())):
    a = torch.randn(M, K, device=dev, dtype=dt)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev, dtype=dt) if not kw else None
    for _ in range(3):
        ops.gemm(a, w, bias=b, residual=r, **kw)
    torch.cuda.synchronize()
print("ok")
