"""What paces the GEGLU GEMMs?  The GEGLU problems of the UNet step with parts of the kernel switched off (experiment
build, pcdm_set_gemm_debug; results are wrong under a non-zero mask — timing only): 1 no TMA stores, 8 no epilogue body,
16 no MMAs, 32 no GEGLU arithmetic (TMEM reads + stores only); us per call inside a CUDA graph of 10 calls."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401
import torch
from pcdms_b200 import ops, lib
from tools.dev_cg import graph_us, rnd, dt, dev

L = lib.load()
masks = [0, 1, 32, 33, 8, 16, 24]
print("mask:".ljust(34) + "".join(f"{m:8d}" for m in masks))
for M, N, K in ((32768, 2560, 320), (8192, 5120, 640), (2048, 10240, 1280)):
    a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), torch.randn(N, device=dev)
    out = torch.empty(M, N // 2, device=dev, dtype=dt)
    row = []
    for m in masks:
        L.pcdm_set_gemm_debug(C.c_int(m))
        row.append(graph_us(lambda: ops.gemm(a, w, bias=b, geglu=True, out=out)))
    L.pcdm_set_gemm_debug(C.c_int(0))
    print(f"gemm M{M} N{N} K{K} geglu".ljust(34) + "".join(f"{t:8.1f}" for t in row), flush=True)
