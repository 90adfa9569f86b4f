"""Enumerate every distinct conv / GEMM problem of one UNet evaluation (BASELINE config 2: B=16, 32x64, 258 tokens, bf16)
and time it under each (bn, cta_group) variant plus the automatic choice.  Output: gpurun_out/autotune.json.
`python tools/autotune.py auto [out.json]` times only the automatic choice (A/B runs of a kernel change)."""
import json
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch
from pcdms_b200 import ops, lib
from pcdms_b200.unet import B200UNet2DConditionModel

dev, dt = "cuda", torch.bfloat16
L = lib.load()
m = B200UNet2DConditionModel(dtype=dt, device=dev, in_channels=9, class_embed_type="projection", projection_class_embeddings_input_dim=1024)
m.load_state_dict(m.synthetic_state_dict(0))
B, h, w = 16, 32, 64
x9 = torch.randn(B, h, w, 64, device=dev).to(dt); t = torch.tensor([981.0], device=dev)
ctx = torch.randn(B, 258, 1024, device=dev).to(dt); cls = torch.randn(B, 1024, device=dev).to(dt)
pose = (0.1 * torch.randn(B, h, w, 320, device=dev)).to(dt)
kv = m.context_kv(ctx)

shapes = {}
orig_conv, orig_gemm = ops.conv3x3, ops.gemm
def rec_conv(x, wp, out=None, **k):
    key = ("conv", tuple(x.shape), wp.shape[0], k.get("stride", 1), k.get("residual") is not None, bool(k.get("out_f32")))
    shapes.setdefault(key, [0, (x, wp, dict(k))])[0] += 1
    return orig_conv(x, wp, out, **k)
def rec_gemm(a, wt, out=None, **k):
    key = ("gemm", a.shape[0], wt.shape[0], wt.shape[1], k.get("residual") is not None, bool(k.get("geglu")), k.get("a2") is not None, bool(k.get("out_f32")))
    shapes.setdefault(key, [0, (a, wt, dict(k))])[0] += 1
    return orig_gemm(a, wt, out, **k)
ops.conv3x3, ops.gemm = rec_conv, rec_gemm
m.forward_nhwc(x9, t, kv, cls, pose)
ops.conv3x3, ops.gemm = orig_conv, orig_gemm
torch.cuda.synchronize()

def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    torch.cuda._sleep(2_000_000)   # ~1 ms of spin: the CPU queues all n launches behind it (GPU-bound timing)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us

AUTO_ONLY = len(sys.argv) > 1 and sys.argv[1] == "auto"
OUT = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/autotune.json"
res = []
tot_auto = tot_best = 0.0
for key, (count, (a, wt, k)) in sorted(shapes.items(), key=lambda kv_: str(kv_[0])):
    k = {kk: vv for kk, vv in k.items() if kk != "bn"}
    fn0 = (lambda bn: orig_conv(a, wt, bn=bn, **k)) if key[0] == "conv" else (lambda bn: orig_gemm(a, wt, bn=bn, **k))
    L.pcdm_set_gemm_cta_group(0)
    row = {"key": str(key), "count": count, "auto_us": timeit(lambda: fn0(0))}
    best = ("auto", row["auto_us"])
    N = wt.shape[0]
    for cg in (() if AUTO_ONLY else (1, 2)):
        L.pcdm_set_gemm_cta_group(cg)
        for bn in (64, 128, 160, 256):
            if cg == 2 and bn < 128: continue
            if key[0] == "gemm" and key[5] and bn % 64: continue
            if bn == 160 and N % 160: continue
            if bn > 64 and N <= 32: continue
            try:
                us = timeit(lambda: fn0(bn))
            except Exception as ex:
                us = None
            row[f"cg{cg}_bn{bn}"] = us
            if us is not None and us < best[1]: best = (f"cg{cg}_bn{bn}", us)
    L.pcdm_set_gemm_cta_group(0)
    row["best"], row["best_us"] = best
    tot_auto += count * row["auto_us"]; tot_best += count * best[1]
    res.append(row)
    print(json.dumps(row), flush=True)
print("TOTAL auto us", tot_auto, "best us", tot_best)
json.dump(res, open(OUT, "w"), indent=1)
