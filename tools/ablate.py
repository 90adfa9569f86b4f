"""In-graph cost of each kernel class at BASELINE config 2 (B=16, 32x64 latents, 258 tokens, bf16).

A CUDA graph of one UNet evaluation is captured with one class of `pcdms_b200.ops` calls replaced by a no-op (its
output buffer is left uninitialised) and replayed; the drop against the full graph is what that class really costs
inside the graph — launch gaps, programmatic-dependent-launch overlap and cache state included — which per-kernel
eager timings and ncu (serialised, cold caches) cannot show.  Results are timing-only: the outputs are garbage.
usage: python tools/ablate.py [out.json]
"""
import json
import sys

sys.path.insert(0, ".")
import torch

from pcdms_b200 import ops
from pcdms_b200.unet import B200UNet2DConditionModel

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ablate.json"
dev, dt = "cuda", torch.bfloat16
m = B200UNet2DConditionModel(dtype=dt, device=dev, in_channels=9, class_embed_type="projection",
                             projection_class_embeddings_input_dim=1024)
m.load_state_dict(m.synthetic_state_dict(0))
B, h, w = 16, 32, 64
x9 = torch.randn(B, h, w, 64, device=dev).to(dt)
t = torch.tensor([981.0], device=dev)
ctx = torch.randn(B, 258, 1024, device=dev).to(dt)
cls = torch.randn(B, 1024, device=dev).to(dt)
pose = (0.1 * torch.randn(B, h, w, 320, device=dev)).to(dt)
kv = m.context_kv(ctx)
real = {n: getattr(ops, n) for n in ["gemm", "conv3x3", "conv3x3_up2x", "groupnorm", "layernorm", "attention"]}


def _fake_chan_stats(out, hw):
    M, N = out.numel() // out.shape[-1], out.shape[-1]
    if hw % 32:
        return None
    return ops.ChanStats(torch.ones((M // 32, N, 2), device=out.device, dtype=torch.float32), hw)


def _empty_like_result(name, args, kw):
    if kw.get("out") is not None:
        return kw["out"]
    if name == "gemm":
        a, wgt = args[0], args[1]
        n = wgt.shape[0] // 2 if kw.get("geglu") else wgt.shape[0]
        out = torch.empty((a.shape[0], n), device=a.device, dtype=torch.float32 if kw.get("out_f32") else a.dtype)
        if kw.get("row_stats"):
            return out, ops.RowStats(torch.ones((2, a.shape[0], 2), device=a.device, dtype=torch.float32), 2)
        if kw.get("chan_stats"):
            return out, _fake_chan_stats(out, kw.get("rows_per_image", 1))
        return out
    if name == "conv3x3":
        x, wgt = args[0], args[1]
        s = kw.get("stride", 1)
        out = torch.empty((x.shape[0], x.shape[1] // s, x.shape[2] // s, wgt.shape[0]), device=x.device,
                          dtype=torch.float32 if kw.get("out_f32") else x.dtype)
        return (out, _fake_chan_stats(out, out.shape[1] * out.shape[2])) if kw.get("chan_stats") else out
    if name == "conv3x3_up2x":
        x, wgt = args[0], args[1]
        out = torch.empty((x.shape[0], 2 * x.shape[1], 2 * x.shape[2], wgt.shape[1]), device=x.device, dtype=x.dtype)
        return (out, _fake_chan_stats(out, out.shape[1] * out.shape[2])) if kw.get("chan_stats") else out
    if name == "groupnorm":
        x1, x2 = args[0], kw.get("x2")
        c = x1.shape[-1] + (x2.shape[-1] if x2 is not None else 0)
        return torch.empty((*x1.shape[:-1], c), device=x1.device, dtype=x1.dtype)
    if name == "layernorm":
        return torch.empty_like(args[0])
    if name == "attention":
        q, heads = args[0], args[4]
        return torch.empty((q.shape[0], heads * 64), device=q.device, dtype=q.dtype)
    raise KeyError(name)


def skipper(name, pred=lambda args, kw: True):
    def f(*args, **kw):
        if pred(args, kw):
            return _empty_like_result(name, args, kw)
        return real[name](*args, **kw)
    return f


def time_graph(reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m.forward_nhwc(x9, t, kv, cls, pose)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        m.forward_nhwc(x9, t, kv, cls, pose)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def is_small_k(args, kw):
    a, a2 = args[0], kw.get("a2")
    return a.shape[1] + (a2.shape[1] if a2 is not None else 0) <= 640 and not kw.get("geglu")


cases = [
    ("full", {}),
    ("no groupnorm", {"groupnorm": skipper("groupnorm")}),
    ("no layernorm", {"layernorm": skipper("layernorm")}),
    ("no attention", {"attention": skipper("attention")}),
    ("no self-attention 2048", {"attention": skipper("attention", lambda a, k: a[0].shape[0] // a[3] == 2048 and a[1].shape[0] // a[3] == 2048)}),
    ("no conv3x3", {"conv3x3": skipper("conv3x3"), "conv3x3_up2x": skipper("conv3x3_up2x")}),
    ("no Upsample2D convs", {"conv3x3_up2x": skipper("conv3x3_up2x")}),
    ("no conv3x3 at 32x64", {"conv3x3": skipper("conv3x3", lambda a, k: a[0].shape[1] == 32)}),
    ("no conv3x3 at 16x32", {"conv3x3": skipper("conv3x3", lambda a, k: a[0].shape[1] == 16)}),
    ("no conv3x3 at 8x16", {"conv3x3": skipper("conv3x3", lambda a, k: a[0].shape[1] == 8)}),
    ("no conv3x3 at 4x8", {"conv3x3": skipper("conv3x3", lambda a, k: a[0].shape[1] == 4)}),
    ("no gemm", {"gemm": skipper("gemm")}),
    ("no GEGLU gemm", {"gemm": skipper("gemm", lambda a, k: bool(k.get("geglu")))}),
    ("no gemm with K <= 640 (non-GEGLU)", {"gemm": skipper("gemm", is_small_k)}),
    ("no gemm with M = 32768", {"gemm": skipper("gemm", lambda a, k: a[0].shape[0] == 32768)}),
    ("no gemm with M = 8192", {"gemm": skipper("gemm", lambda a, k: a[0].shape[0] == 8192)}),
    ("no gemm with M = 2048", {"gemm": skipper("gemm", lambda a, k: a[0].shape[0] == 2048)}),
    ("no gemm with M <= 512", {"gemm": skipper("gemm", lambda a, k: a[0].shape[0] <= 512)}),
    ("nothing but launches skipped: all", {n: skipper(n) for n in ["gemm", "conv3x3", "conv3x3_up2x", "groupnorm", "layernorm", "attention"]}),
]
res = {}
full = None
for label, patch in cases:
    for n, f in patch.items():
        setattr(ops, n, f)
    try:
        ms = time_graph()
    finally:
        for n in patch:
            setattr(ops, n, real[n])
    full = ms if full is None else full
    res[label] = ms
    print(f"{label:44s} {ms:8.3f} ms   delta {full - ms:7.3f} ms ({100 * (full - ms) / full:5.1f} %)", flush=True)
json.dump(res, open(out_path, "w"), indent=1)
