"""The 320-wide CTA-pair tile (two 160-column MMAs from one activation stage, single-buffered 320-column accumulator)
against the automatic choice without it, on the UNet's N % 320 == 0 shapes: us per call inside a CUDA graph, and
agreement of the results."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops

dt, dev = torch.bfloat16, "cuda"


def graph_us(fn, n=10, reps=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)


def rnd(*shape, scale=1.0):
    return (scale * torch.randn(*shape, device=dev)).to(dt)


rows = []
for (M, N, K, kw) in [(32768, 320, 320, dict(res=True)), (32768, 320, 320, {}), (32768, 320, 320, dict(res=True, stats=True)),
                      (32768, 320, 1280, dict(res=True)), (32768, 960, 320, {}), (32768, 2560, 320, dict(geglu=True)),
                      (8192, 640, 640, dict(res=True)), (8192, 640, 640, {}), (8192, 1920, 640, {}), (8192, 640, 2560, dict(res=True)),
                      (8192, 5120, 640, dict(geglu=True)), (2048, 1280, 1280, dict(res=True)), (2048, 3840, 1280, {}),
                      (2048, 1280, 5120, dict(res=True)), (2048, 10240, 1280, dict(geglu=True))]:
    a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), torch.randn(N, device=dev)
    r = rnd(M, N) if kw.get("res") else None
    out = torch.empty(M, N // 2 if kw.get("geglu") else N, device=dev, dtype=dt)
    args = dict(bias=b, residual=r, geglu=bool(kw.get("geglu")), out=out, row_stats=bool(kw.get("stats")))
    t = {}
    ref = None
    for bn in (0, 160, 256, 320):
        if kw.get("geglu") and bn == 160:
            continue
        try:
            t[bn] = graph_us(lambda: ops.gemm(a, w, bn=bn, **args))
        except Exception as ex:
            t[bn] = float("nan")
            continue
        got = out.float().clone()
        if ref is None:
            ref = got
        elif not torch.allclose(got, ref, rtol=2e-2, atol=2e-2):
            print("MISMATCH", M, N, K, kw, bn, float((got - ref).abs().max()))
    rows.append((f"gemm M{M} N{N} K{K} {'+'.join(kw) or '-'}", t))
    print(rows[-1][0].ljust(44), "  ".join(f"bn{k} {v:7.2f}" for k, v in t.items()), flush=True)

for (B, H, W, Cin, Cout, kw) in [(16, 32, 64, 320, 320, dict(res=True)), (16, 32, 64, 320, 320, {}), (16, 32, 64, 640, 320, {}),
                                 (16, 32, 64, 960, 320, {}), (16, 32, 64, 640, 640, {}), (16, 16, 32, 640, 640, dict(res=True)),
                                 (16, 16, 32, 1280, 640, {}), (16, 8, 16, 1280, 1280, dict(res=True)), (16, 16, 32, 320, 640, {})]:
    x, w, b = rnd(B, H, W, Cin), rnd(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5), torch.randn(Cout, device=dev)
    tv = torch.randn(B, Cout, device=dev)
    r = rnd(B, H, W, Cout) if kw.get("res") else None
    out = torch.empty(B, H, W, Cout, device=dev, dtype=dt)
    t = {}
    ref = None
    for bn in (0, 160, 320):
        t[bn] = graph_us(lambda: ops.conv3x3(x, w, bias=b, rowvec=tv, residual=r, out=out, bn=bn))
        got = out.float().clone()
        if ref is None:
            ref = got
        elif not torch.allclose(got, ref, rtol=2e-2, atol=2e-2):
            print("MISMATCH conv", B, H, W, Cin, Cout, bn, float((got - ref).abs().max()))
    print(f"conv {H}x{W} {Cin}->{Cout} {'+'.join(kw) or '-'}".ljust(44), "  ".join(f"bn{k} {v:7.2f}" for k, v in t.items()), flush=True)
