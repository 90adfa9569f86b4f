"""Dev check (GPU): tcgen05 GEMM / conv kernels vs torch CPU fp32 on identical 16-bit inputs."""
import sys, time, json
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from pcdms_b200 import ops

torch.manual_seed(0)
dev = "cuda"
res = []

def rel(a, b):
    return ((a.float().cpu() - b).abs().max() / (b.abs().max() + 1e-9)).item()

def run_gemm(M, N, K, dt, bn=0, bias=True, resid=False, geglu=False, split=0):
    a = torch.randn(M, K).to(dt); w = (torch.randn(N, K) / K ** 0.5).to(dt)
    b = torch.randn(N) if bias else None
    r = torch.randn(M, N).to(dt) if resid else None
    ref = a.float() @ w.float().t()
    if bias: ref = ref + b
    if resid: ref = ref + r.float()
    wd = w
    bd = b
    if geglu:
        h, g = ref.chunk(2, dim=1)
        ref = h * F.gelu(g)
        perm = ops.geglu_row_permutation(N // 2)
        wd = w[perm].contiguous(); bd = b[perm].contiguous() if bias else None
    ad = a.to(dev)
    kw = {}
    if split:
        kw["a2"] = ad[:, split:]
        ad = ad[:, :split]
    out = ops.gemm(ad, wd.to(dev), bias=bd.to(dev) if bias else None, residual=r.to(dev) if resid else None,
                   geglu=geglu, bn=bn, **kw)
    torch.cuda.synchronize()
    e = rel(out, ref)
    res.append(dict(op="gemm", M=M, N=N, K=K, dt=str(dt), bn=bn, geglu=geglu, split=split, relerr=e))
    print(res[-1], flush=True)

def run_conv(B, H, W, Cin, Cout, dt, stride=1, bn=0, temb=True, resid=True):
    x = torch.randn(B, Cin, H * stride, W * stride).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3) / (9 * Cin) ** 0.5).to(dt)
    b = torch.randn(Cout)
    t = torch.randn(B, Cout) if temb else None
    r = torch.randn(B, Cout, H, W).to(dt) if resid else None
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1)
    if temb: ref = ref + t[:, :, None, None]
    if resid: ref = ref + r.float()
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    out = ops.conv3x3(xd, ops.pack_conv3x3_weight(w, dt).to(dev), bias=b.to(dev), rowvec=t.to(dev) if temb else None,
                      residual=r.permute(0, 2, 3, 1).contiguous().to(dev) if resid else None, stride=stride, bn=bn)
    torch.cuda.synchronize()
    e = rel(out.permute(0, 3, 1, 2), ref)
    res.append(dict(op="conv", B=B, H=H, W=W, Cin=Cin, Cout=Cout, dt=str(dt), stride=stride, bn=bn, relerr=e))
    print(res[-1], flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("gemm", "all"):
    for dt in (torch.float16, torch.bfloat16):
        run_gemm(256, 256, 128, dt, bn=128)
        run_gemm(1000, 320, 320, dt, bn=0, resid=True)
        run_gemm(4096, 1280, 1024, dt, bn=256)
        run_gemm(516, 640, 1024, dt, bn=64, bias=False)
        run_gemm(300, 2560, 320, dt, bn=0, geglu=True)
        run_gemm(512, 320, 960, dt, bn=160, split=640)
        run_gemm(2, 1280, 320, dt, bn=0)
if which in ("conv", "all"):
    for dt in (torch.float16, torch.bfloat16):
        run_conv(2, 32, 64, 320, 320, dt)
        run_conv(2, 16, 32, 640, 640, dt, bn=128)
        run_conv(2, 4, 8, 1280, 1280, dt)
        run_conv(6, 4, 8, 128, 64, dt)
        run_conv(1, 64, 128, 64, 64, dt, resid=False)
        run_conv(2, 16, 32, 320, 320, dt, stride=2, temb=False, resid=False)
        run_conv(3, 4, 8, 128, 128, dt, stride=2, temb=False, resid=False)
if which in ("perf", "all"):
    # timing: conv 320->320 @ 16x32x64 and gemm 32768x2560x320
    dt = torch.bfloat16
    x = torch.randn(16, 32, 64, 320, device=dev, dtype=dt); w = torch.randn(320, 2880, device=dev, dtype=dt)
    for bn in (64, 128, 160, 256):
        for _ in range(3): ops.conv3x3(x, w, bn=bn)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
        e0.record()
        for _ in range(20): ops.conv3x3(x, w, bn=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 2 * 16 * 2048 * 320 * 2880
        res.append(dict(op="perf_conv320", bn=bn, ms=ms, tflops=fl / ms / 1e9)); print(res[-1], flush=True)
    x = torch.randn(16, 8, 16, 1280, device=dev, dtype=dt); w = torch.randn(1280, 9 * 1280, device=dev, dtype=dt)
    for bn in (64, 128, 256):
        for _ in range(3): ops.conv3x3(x, w, bn=bn)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
        e0.record()
        for _ in range(20): ops.conv3x3(x, w, bn=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 2 * 16 * 128 * 1280 * 9 * 1280
        res.append(dict(op="perf_conv1280", bn=bn, ms=ms, tflops=fl / ms / 1e9)); print(res[-1], flush=True)
    a = torch.randn(32768, 320, device=dev, dtype=dt); w = torch.randn(2560, 320, device=dev, dtype=dt)
    for bn in (128, 256):
        for _ in range(3): ops.gemm(a, w, bn=bn)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
        e0.record()
        for _ in range(20): ops.gemm(a, w, bn=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 2 * 32768 * 320 * 2560
        res.append(dict(op="perf_gemm_ff", bn=bn, ms=ms, tflops=fl / ms / 1e9)); print(res[-1], flush=True)
    a = torch.randn(8192, 8192, device=dev, dtype=dt); w = torch.randn(8192, 8192, device=dev, dtype=dt)
    for bn in (256,):
        for _ in range(3): ops.gemm(a, w, bn=bn)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
        e0.record()
        for _ in range(10): ops.gemm(a, w, bn=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2 * 8192 ** 3
        res.append(dict(op="perf_gemm_8k", bn=bn, ms=ms, tflops=fl / ms / 1e9)); print(res[-1], flush=True)
json.dump(res, open(f"gpurun_out/dev_igemm_{which}.json", "w"), indent=1)
