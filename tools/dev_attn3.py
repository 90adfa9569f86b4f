"""Dev check (GPU): the experiment build's persistent two-tile attention kernel (speculative exponentials) under each
FMA-pipe exp2 share (pcdm_set_attention_debug 0..3 = none, 1/4, 1/3, 1/2 of the column pairs): correctness against an
fp32 softmax reference — including inputs whose row maxima keep growing from block to block (the redo / rescale path) —
and timing inside a CUDA graph next to the release kernel."""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import ctypes as C
import torch
from pcdms_b200 import ops, lib
from tools.dev_attn2 import graph_us  # noqa: E402  (prints that tool's table for the default kernel first)

L = lib.load()
dev = "cuda"


def ref(q, k, v, B, heads):
    Sq, Skv = q.shape[0] // B, k.shape[0] // B
    qh = q.float().view(B, Sq, heads, 64).transpose(1, 2)
    kh = k.float().view(B, Skv, heads, 64).transpose(1, 2)
    vh = v.float().view(B, Skv, heads, 64).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B * Sq, heads * 64)


torch.manual_seed(0)
VARIANTS = [(0, 0), (1, 0), (1, 3)]   # (two-tile kernel on, pcdm_set_attention_debug code: 0 release = 1/4 of the exp2 on the FMA pipe, 1 none, 2 = 1/3, 3 = release + early score hand-back, 4 = none + early hand-back, 5 = half-row threads + release share, 6 = half-row threads, no FMA-pipe exp2)
for dt in (torch.float16, torch.bfloat16):
    for (B, heads, Sq, Skv, grow) in [(2, 5, 2048, 2048, 0), (2, 5, 1024, 1024, 1), (2, 10, 512, 512, 0), (2, 5, 2048, 258, 0),
                                      (1, 2, 384, 300, 1), (1, 5, 256, 95, 0), (3, 20, 128, 128, 0), (2, 3, 200, 2048, 1)]:
        Cc = heads * 64
        q = torch.randn(B * Sq, Cc, device=dev)
        k = torch.randn(B * Skv, Cc, device=dev)
        v = torch.randn(B * Skv, Cc, device=dev)
        if grow:   # key norms ramp up along the sequence: later blocks bring much larger scores => rescales
            ramp = torch.linspace(0.2, 6.0, Skv, device=dev).repeat(B)[:, None]
            k = k * ramp
        q, k, v = q.to(dt), k.to(dt), v.to(dt)
        want = ref(q, k, v, B, heads)
        line = f"{str(dt)[6:]:9s} B{B} h{heads} Sq{Sq} Skv{Skv} grow{grow}:"
        for v2, code in VARIANTS:
            L.pcdm_set_attention_v2(C.c_int(v2))
            L.pcdm_set_attention_debug(C.c_int(code))
            got = ops.attention(q, k, v, B, heads).float()
            err = (got - want).abs().max().item()
            line += f"  [{v2}{code}] {err:.2e}{' NAN' if bool(torch.isnan(got).any()) else ''}"
        print(line, flush=True)

dt = torch.bfloat16
shapes = [(16, 5, 2048, 2048), (16, 10, 512, 512), (16, 20, 128, 128), (16, 5, 2048, 258), (16, 10, 512, 258), (16, 20, 128, 258),
          (8, 5, 8192, 8192)]
for (B, heads, Sq, Skv) in shapes:
    Cc = heads * 64
    q = torch.randn(B * Sq, Cc, device=dev).to(dt)
    k = torch.randn(B * Skv, Cc, device=dev).to(dt)
    v = torch.randn(B * Skv, Cc, device=dev).to(dt)
    out = torch.empty_like(q)
    line = f"B{B} h{heads} Sq{Sq} Skv{Skv}:"
    for v2, code in VARIANTS:
        L.pcdm_set_attention_v2(C.c_int(v2))
        L.pcdm_set_attention_debug(C.c_int(code))
        us = graph_us(lambda: ops.attention(q, k, v, B, heads, out=out))
        line += f"  [{v2}{code}] {us:7.2f} us"
    print(line, flush=True)
L.pcdm_set_attention_debug(C.c_int(0))
L.pcdm_set_attention_v2(C.c_int(0))
