"""Per-op gap table of one UNet evaluation at BASELINE config 2 (B=16, 32x64 latents, 258 tokens, bf16).

Every `pcdms_b200.ops` call is bracketed by CUDA events while the CPU runs ahead of the GPU (a long spin kernel is
queued first, so launch latency is not in the intervals).  Each distinct (op, shape) is then compared with its own
roofline: max(flops / tensor peak, unique bytes / HBM peak) with the measured peaks of MEASURED_PEAKS.json.  The table
is sorted by the summed gap, i.e. by where the remaining milliseconds are.
usage: python tools/op_table.py [reps] [out.json]
"""
import json
import sys
from collections import OrderedDict

sys.path.insert(0, ".")
import torch

from pcdms_b200 import ops
from pcdms_b200.unet import B200UNet2DConditionModel

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/op_table.json"
try:
    pk = json.load(open("MEASURED_PEAKS.json"))
    TF, GBS = pk["bf16_tflops"], pk["hbm_gbs"]
except Exception:
    TF, GBS = 1690.0, 6570.0

records = []          # (key, flops, bytes, ev0, ev1)


def _nb(*ts):
    n = 0
    for t in ts:
        if isinstance(t, (tuple, list)):      # (tensor, statistics) results of the round-2 fused epilogues
            n += _nb(*t)
        elif isinstance(t, torch.Tensor):
            n += t.numel() * t.element_size()
    return n


def describe(name, args, kw, result):
    if name == "gemm":
        a, w = args[0], args[1]
        a2 = kw.get("a2")
        M, K = a.shape[0], a.shape[1] + (a2.shape[1] if a2 is not None else 0)
        N = w.shape[0]
        tag = "".join(c for c, on in (("g", kw.get("geglu")), ("r", kw.get("residual") is not None),
                                      ("v", kw.get("rowvec") is not None), ("s", kw.get("silu")),
                                      ("f", kw.get("out_f32"))) if on)
        return f"gemm M{M} N{N} K{K} {tag}", 2.0 * M * N * K, _nb(a, a2, w, result, kw.get("residual"))
    if name == "conv3x3":
        x, w = args[0], args[1]
        B, H, W, Cin = x.shape
        s = kw.get("stride", 1)
        Cout = w.shape[0]
        tag = "r" if kw.get("residual") is not None else ""
        return (f"conv {H}x{W} {Cin}->{Cout} s{s} {tag}", 2.0 * B * (H // s) * (W // s) * Cout * 9 * Cin,
                _nb(x, w, result, kw.get("residual")))
    if name == "conv3x3_up2x":
        x, w = args[0], args[1]
        B, H, W, Cin = x.shape
        Cout = w.shape[1] if w.dim() == 3 else w.shape[0] // 4 if w.dim() == 2 and w.shape[0] % 4 == 0 else w.shape[0]
        r0 = result[0] if isinstance(result, (tuple, list)) else result
        Cout = r0.shape[-1]
        return f"up2x+conv {H}x{W}->{2 * H}x{2 * W} {Cin}->{Cout}", 2.0 * B * 4 * H * W * Cout * 9 * Cin, _nb(x, w, result)
    if name == "groupnorm":
        x1, x2 = args[0], kw.get("x2")
        C_ = x1.shape[-1] + (x2.shape[-1] if x2 is not None else 0)
        return f"groupnorm {tuple(x1.shape[1:3])} C{C_}{' cat' if x2 is not None else ''}", 0.0, _nb(x1, x2, result)
    if name == "layernorm":
        return f"layernorm {tuple(args[0].shape)}", 0.0, _nb(args[0], result)
    if name == "attention":
        q, k, v, B, heads = args[:5]
        Sq, Skv = q.shape[0] // B, k.shape[0] // B
        return (f"attention Sq{Sq} Skv{Skv} h{heads}", 4.0 * B * heads * Sq * Skv * 64,
                2 * (2 * B * Sq + 2 * B * Skv) * heads * 64)
    return f"{name}", 0.0, _nb(*[a for a in args if torch.is_tensor(a)], result if torch.is_tensor(result) else None)


def wrap(name):
    fn = getattr(ops, name)

    def inner(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*args, **kw)
        e1.record()
        key, fl, by = describe(name, args, kw, r)
        records.append((key, fl, by, e0, e1))
        return r
    return inner


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
for _ in range(2):
    m.forward_nhwc(x9, t, kv, cls, pose)
torch.cuda.synchronize()

for n in ["gemm", "conv3x3", "conv3x3_up2x", "groupnorm", "layernorm", "attention", "timestep_embedding",
          "upsample_nearest2x"]:
    setattr(ops, n, wrap(n))

table = OrderedDict()
total_ms = 0.0
for _ in range(reps):
    records.clear()
    torch.cuda._sleep(60_000_000)          # ~30 ms of spin: the CPU queues the whole evaluation behind it
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    m.forward_nhwc(x9, t, kv, cls, pose)
    s1.record()
    torch.cuda.synchronize()
    total_ms += s0.elapsed_time(s1)
    for key, fl, by, e0, e1 in records:
        row = table.setdefault(key, {"count": 0, "us": 0.0, "flops": fl, "bytes": by})
        row["count"] += 1
        row["us"] += e0.elapsed_time(e1) * 1e3
rows = []
for key, r in table.items():
    per = r["count"] / reps
    us = r["us"] / r["count"]
    t_fl, t_by = r["flops"] / (TF * 1e6), r["bytes"] / (GBS * 1e3)     # us
    roof = max(t_fl, t_by)
    rows.append({"op": key, "per_step": per, "us": us, "roof_us": roof, "bound": "tensor" if t_fl >= t_by else "hbm",
                 "eff": roof / us, "tflops": r["flops"] / us / 1e6, "gbs": r["bytes"] / us / 1e3,
                 "total_us": us * per, "gap_us": (us - roof) * per})
rows.sort(key=lambda r: -r["gap_us"])
tot = sum(r["total_us"] for r in rows)
print(f"eager UNet evaluation {total_ms / reps:.3f} ms; summed op time {tot / 1e3:.3f} ms; "
      f"summed roofline {sum(r['roof_us'] * r['per_step'] for r in rows) / 1e3:.3f} ms")
print(f"{'op':44s} {'n':>3s} {'us':>8s} {'roof':>7s} {'eff':>5s} {'TF/s':>6s} {'GB/s':>6s} {'tot us':>8s} {'gap us':>8s}")
for r in rows:
    print(f"{r['op']:44s} {r['per_step']:3.0f} {r['us']:8.1f} {r['roof_us']:7.1f} {r['eff']:5.2f} {r['tflops']:6.0f} "
          f"{r['gbs']:6.0f} {r['total_us']:8.1f} {r['gap_us']:8.1f}")
json.dump({"eager_ms": total_ms / reps, "rows": rows, "peaks": {"tflops": TF, "gbs": GBS}}, open(out_path, "w"), indent=1)
