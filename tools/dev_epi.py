"""Epilogue-bound problems of the UNet step with the automatic tile choice: us per call inside a CUDA graph (10 calls per
graph).  A/B of library builds: `PCDM_B200_LIB=.ab/<variant>.so python tools/dev_epi.py` (tools/build_variant.py)."""
import os
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
from tools.dev_cg import graph_us, rnd, dt, dev   # noqa: E402  (dev_cg's module-level sweep is skipped below)

torch.manual_seed(0)
rows = []
for (M, N, K, kw) in [(32768, 2560, 320, dict(geglu=True)), (8192, 5120, 640, dict(geglu=True)),
                      (2048, 10240, 1280, dict(geglu=True)), (32768, 320, 320, dict(res=True)), (32768, 320, 320, {}),
                      (32768, 960, 320, {}), (32768, 320, 1280, dict(res=True)), (8192, 640, 640, dict(res=True)),
                      (8192, 1920, 640, {}), (2048, 1280, 1280, dict(res=True)), (2048, 3840, 1280, {})]:
    a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), torch.randn(N, device=dev)
    r = rnd(M, N) if kw.get("res") else None
    out = torch.empty(M, N // 2 if kw.get("geglu") else N, device=dev, dtype=dt)
    t = graph_us(lambda: ops.gemm(a, w, bias=b, residual=r, geglu=bool(kw.get("geglu")), out=out))
    rows.append((f"gemm M{M} N{N} K{K} {'+'.join(kw) or '-'}", t))
for (B, H, W, Ci, Co, res) in [(16, 32, 64, 320, 320, True), (16, 32, 64, 640, 320, False), (16, 32, 64, 960, 320, False),
                               (16, 32, 64, 640, 640, False), (16, 16, 32, 640, 640, True), (16, 16, 32, 1280, 640, False), (16, 16, 32, 1920, 640, False),
                               (16, 8, 16, 1280, 1280, True), (16, 8, 16, 2560, 1280, False), (16, 4, 8, 1280, 1280, True)]:
    x = rnd(B, H, W, Ci)
    wp = ops.pack_conv3x3_weight(torch.randn(Co, Ci, 3, 3, device=dev) * (9 * Ci) ** -0.5, dt)
    b = torch.randn(Co, device=dev)
    r = rnd(B, H, W, Co) if res else None
    out = torch.empty(B, H, W, Co, device=dev, dtype=dt)
    t = graph_us(lambda: ops.conv3x3(x, wp, bias=b, residual=r, out=out))
    rows.append((f"conv {H}x{W} {Ci}->{Co}{' +res' if res else ''}", t))
print("library:", os.environ.get("PCDM_B200_LIB", "release"))
for name, t in rows:
    print(f"{name:42s} {t:8.2f} us", flush=True)
