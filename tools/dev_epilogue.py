"""What paces the short-K GEMMs (K = 320..1280: five to twenty k-blocks per tile)?  Times each problem GPU-bound with
parts of the kernel switched off through pcdm_set_gemm_debug (results are wrong under a non-zero mask; timing only):
1 no TMA stores, 2 no residual, 4 no bias, 8 (+2) no epilogue body, 16 no MMAs."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch
from pcdms_b200 import ops, lib
dev, dt = "cuda", torch.bfloat16
L = lib.load()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(5_000_000)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


import os
masks = [0, 1, 2, 4, 7, 10, 16, 26]
if os.environ.get("PCDM_STAGES"):      # ring-depth sweep: do queued mainloop loads delay the epilogue's TMA traffic?
    L.pcdm_set_gemm_max_stages(C.c_int(int(os.environ["PCDM_STAGES"])))
    masks = [0, 10]
print("mask:" + "".join(f"{m:8d}" for m in masks))
for M, N, K, res in ((32768, 320, 320, True), (32768, 320, 320, False), (8192, 640, 640, True), (2048, 1280, 1280, True),
                     (32768, 320, 1280, True), (32768, 960, 320, False)):
    a = torch.randn(M, K, device=dev, dtype=dt)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev, dtype=dt) if res else None
    out = torch.empty(M, N, device=dev, dtype=dt)
    for bn, cg in ((0, 0), (160, 1), (128, 2), (64, 1)):
        L.pcdm_set_gemm_cta_group(C.c_int(cg))
        row = []
        for m in masks:
            L.pcdm_set_gemm_debug(C.c_int(m))
            row.append(timeit(lambda: ops.gemm(a, w, out=out, bias=b, residual=r, bn=bn)))
        L.pcdm_set_gemm_debug(C.c_int(0))
        print(f"M{M} N{N} K{K} res={int(res)} bn={bn} cg={cg}: " + "".join(f"{t:8.1f}" for t in row), flush=True)
L.pcdm_set_gemm_cta_group(C.c_int(0))
