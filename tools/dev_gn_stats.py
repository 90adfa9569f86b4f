"""GroupNorm at the UNet's shapes: stand-alone kernels vs statistics-from-the-producer (fold + apply), us per call inside
a CUDA graph of 20 calls (launch overhead of the eager path would hide the kernels)."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops

dt, dev = torch.bfloat16, "cuda"


def graph_us(fn, n=20, reps=10):
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


for (B, H, W, C1, C2) in [(16, 32, 64, 320, 0), (16, 32, 64, 640, 0), (16, 32, 64, 640, 320), (16, 16, 32, 640, 0),
                          (16, 16, 32, 1280, 640), (16, 8, 16, 1280, 0), (16, 4, 8, 1280, 1280)]:
    C = C1 + C2
    x1 = torch.randn(B, H, W, C1, device=dev).to(dt)
    x2 = torch.randn(B, H, W, C2, device=dev).to(dt) if C2 else None
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)

    def mk(x):
        return ops.ChanStats(torch.rand((B * H * W // 32, x.shape[-1], 2), device=dev) + 1.0, H * W)

    st = (mk(x1), mk(x2) if C2 else None)
    out = torch.empty(B, H, W, C, device=dev, dtype=dt)
    row = {}
    row["auto"] = graph_us(lambda: ops.groupnorm(x1, gamma, beta, 1e-5, x2=x2, silu=True, out=out))
    row["two_pass"] = graph_us(lambda: ops.groupnorm(x1, gamma, beta, 1e-5, x2=x2, silu=True, out=out, path="two_pass"))
    row["stats"] = graph_us(lambda: ops.groupnorm(x1, gamma, beta, 1e-5, x2=x2, silu=True, out=out, stats=st))
    mb = B * H * W * C * 2 / 1e6
    print(f"B{B} {H}x{W} C{C1}+{C2} ({mb:5.1f} MB): " + "  ".join(f"{k} {v:6.2f} us" for k, v in row.items()), flush=True)
