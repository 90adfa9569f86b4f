"""Attention at the UNet's shapes: us per call inside a CUDA graph (10 calls), TFLOP/s, fraction of the measured burst peak."""
import json
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops

dev, dt = "cuda", torch.bfloat16
peak = 1652.7
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"]
except Exception:
    pass


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


for (B, heads, Sq, Skv) in [(16, 5, 2048, 2048), (16, 10, 512, 512), (16, 20, 128, 128), (16, 5, 2048, 258), (16, 10, 512, 258),
                            (16, 20, 128, 258), (16, 20, 32, 258), (8, 5, 8192, 8192), (16, 5, 4096, 4096), (16, 5, 2048, 95)]:
    C = heads * 64
    q = torch.randn(B * Sq, C, device=dev).to(dt)
    k = torch.randn(B * Skv, C, device=dev).to(dt)
    v = torch.randn(B * Skv, C, device=dev).to(dt)
    out = torch.empty_like(q)
    us = graph_us(lambda: ops.attention(q, k, v, B, heads, out=out))
    fl = 4.0 * B * heads * Sq * Skv * 64
    print(f"attention B{B} h{heads} Sq{Sq} Skv{Skv}: {us:8.2f} us  {fl / us / 1e6:7.1f} TFLOP/s  {fl / us / 1e6 / peak:5.3f} of burst peak", flush=True)
