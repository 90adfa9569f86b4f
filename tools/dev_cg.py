"""CTA-pair (cta_group 2) vs single-CTA tiles at every tile width for the wide short-K GEMMs (GEGLU, fused q/k/v):
us per call inside a CUDA graph."""
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



if __name__ == "__main__":
    for (M, N, K, kw) in [(32768, 2560, 320, dict(geglu=True)), (8192, 5120, 640, dict(geglu=True)), (2048, 10240, 1280, dict(geglu=True)),
                          (32768, 960, 320, {}), (8192, 1920, 640, {}), (2048, 3840, 1280, {}), (32768, 320, 1280, dict(res=True)),
                          (32768, 320, 320, dict(res=True))]:
        a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), torch.randn(N, device=dev)
        r = rnd(M, N) if kw.get("res") else None
        out = torch.empty(M, N // 2 if kw.get("geglu") else N, device=dev, dtype=dt)
        args = dict(bias=b, residual=r, geglu=bool(kw.get("geglu")), out=out)
        t = {}
        for bn in (0, 128, 160, 256):
            for cg in (1, 2):
                if kw.get("geglu") and bn == 160:
                    continue
                try:
                    t[(bn, cg)] = graph_us(lambda: ops.gemm(a, w, bn=bn, cta_group=cg, **args))
                except Exception as ex:
                    t[(bn, cg)] = float("nan")
        print(f"gemm M{M} N{N} K{K} {'+'.join(kw) or '-'}".ljust(40), "  ".join(f"bn{k[0]}/cg{k[1]} {v:6.2f}" for k, v in t.items()), flush=True)
