"""GroupNorm / LayerNorm timing at the bench shapes: back-to-back launches (launch overhead amortised), once on one
buffer pair (L2-warm) and once rotating over 8 buffer pairs (> L2: HBM-cold), next to a device copy of the same size."""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch
from pcdms_b200 import ops, lib
dev = "cuda"; dt = torch.bfloat16
L = lib.load()

def timeit(fn, nbuf, reps=24):
    for i in range(4):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for i in range(reps):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

def bench_gn(B, H, W, C1, C2, label):
    C = C1 + C2
    xs = [torch.randn(B, H, W, C1, device=dev, dtype=dt) for _ in range(8)]
    x2 = [torch.randn(B, H, W, C2, device=dev, dtype=dt) for _ in range(8)] if C2 else None
    ys = [torch.empty(B, H, W, C, device=dev, dtype=dt) for _ in range(8)]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    mb = 2 * B * H * W * C * 2 / 1e6
    def gn(i):
        ops.groupnorm(xs[i], g, b, 1e-5, x2=x2[i] if C2 else None, silu=True, out=ys[i])
    def cp(i):
        ys[i][..., :C1].copy_(xs[i]) if C2 else ys[i].copy_(xs[i])
    res = {}
    for mode, code in (("auto", 0), ("T256", 258), ("T1024", 1026), ("two-pass", 1)):
        L.pcdm_set_groupnorm_two_pass(code)
        try:
            res[mode] = (timeit(gn, 1), timeit(gn, 8))
        except Exception as e:
            res[mode] = None
    L.pcdm_set_groupnorm_two_pass(0)
    res["copy"] = (timeit(cp, 1), timeit(cp, 8)) if not C2 else None
    print(f"GN {label:24s} {mb:6.1f} MB  " + "  ".join(
        f"{k} {v[0]:6.1f}/{v[1]:6.1f}" if v else f"{k} n/a" for k, v in res.items()), flush=True)

def bench_ln(M, C):
    xs = [torch.randn(M, C, device=dev, dtype=dt) for _ in range(8)]
    ys = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    ln = lambda i: ops.layernorm(xs[i], g, b, out=ys[i])
    cp = lambda i: ys[i].copy_(xs[i])
    print(f"LN ({M},{C}) {2 * M * C * 2 / 1e6:6.1f} MB  warm/cold us: ln {timeit(ln, 1):6.1f}/{timeit(ln, 8):6.1f}  "
          f"copy {timeit(cp, 1):6.1f}/{timeit(cp, 8):6.1f}", flush=True)

print("us per launch, warm/cold")
bench_gn(16, 32, 64, 320, 0, "32x64 C320")
bench_gn(16, 32, 64, 320, 320, "32x64 C640 cat")
bench_gn(16, 32, 64, 640, 320, "32x64 C960 cat")
bench_gn(16, 16, 32, 640, 0, "16x32 C640")
bench_gn(16, 16, 32, 1280, 640, "16x32 C1920 cat")
bench_gn(16, 8, 16, 1280, 0, "8x16 C1280")
bench_gn(16, 8, 16, 1280, 1280, "8x16 C2560 cat")
bench_gn(16, 4, 8, 1280, 0, "4x8 C1280")
bench_ln(32768, 320)
bench_ln(8192, 640)
bench_ln(2048, 1280)
