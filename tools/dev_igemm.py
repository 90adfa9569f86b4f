"""Dev check (GPU): tcgen05 GEMM / conv kernels vs torch CPU fp32 on identical 16-bit inputs."""
import sys, time, json
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
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
from pcdms_b200 import lib as _plib
def set_cg(mode):
    _plib.check(_plib.load().pcdm_set_gemm_cta_group(mode), kernels=0)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

if which == "stages":
    dt = torch.bfloat16
    L = _plib.load()
    x320 = torch.randn(16, 32, 64, 320, device=dev, dtype=dt); w320 = torch.randn(320, 2880, device=dev, dtype=dt)
    x1280 = torch.randn(16, 8, 16, 1280, device=dev, dtype=dt); w1280 = torch.randn(1280, 11520, device=dev, dtype=dt)
    a8k = torch.randn(8192, 8192, device=dev, dtype=dt); w8k = torch.randn(8192, 8192, device=dev, dtype=dt)
    for name, fn, fl in [("conv320 bn160", lambda: ops.conv3x3(x320, w320, bn=160), 2 * 16 * 2048 * 320 * 2880),
                         ("conv320 bn256", lambda: ops.conv3x3(x320, w320, bn=256), 2 * 16 * 2048 * 320 * 2880),
                         ("conv1280@8x16 bn256", lambda: ops.conv3x3(x1280, w1280, bn=256), 2 * 16 * 128 * 1280 * 11520),
                         ("conv1280@8x16 bn128", lambda: ops.conv3x3(x1280, w1280, bn=128), 2 * 16 * 128 * 1280 * 11520),
                         ("gemm8k bn256", lambda: ops.gemm(a8k, w8k, bn=256), 2 * 8192 ** 3)]:
        for cg in (1, 2):
            set_cg(cg)
            for st in (2, 3, 4, 5, 6, 8):
                L.pcdm_set_gemm_max_stages(st)
                ms = timeit(fn, 10)
                res.append(dict(op="stages", name=name, cg=cg, max_stages=st, ms=ms, tflops=fl / ms / 1e9)); print(res[-1], flush=True)
    L.pcdm_set_gemm_max_stages(8); set_cg(0)
if which == "splitk":
    ops.ensure_workspace("cuda")
    dt = torch.bfloat16
    x = torch.randn(16, 4, 8, 1280, device=dev, dtype=dt); w = torch.randn(1280, 11520, device=dev, dtype=dt)
    x2 = torch.randn(16, 4, 8, 2560, device=dev, dtype=dt); w2 = torch.randn(1280, 23040, device=dev, dtype=dt)
    b = torch.randn(1280, device=dev); r = torch.randn(16, 4, 8, 1280, device=dev, dtype=dt)
    for name, fn, fl in [("conv1280@4x8 auto(splitK)", lambda: ops.conv3x3(x, w, bias=b, residual=r), 2 * 512 * 1280 * 11520),
                         ("conv1280@4x8 bn64", lambda: ops.conv3x3(x, w, bias=b, residual=r, bn=64), 2 * 512 * 1280 * 11520),
                         ("conv2560@4x8 auto(splitK)", lambda: ops.conv3x3(x2, w2, bias=b, residual=r), 2 * 512 * 1280 * 23040),
                         ("conv2560@4x8 bn64", lambda: ops.conv3x3(x2, w2, bias=b, residual=r, bn=64), 2 * 512 * 1280 * 23040)]:
        ms = timeit(fn, 20)
        res.append(dict(op="splitk", name=name, ms=ms, tflops=fl / ms / 1e9)); print(res[-1], flush=True)
if which == "cg2":
    set_cg(2)
    for dt in (torch.float16, torch.bfloat16):
        run_gemm(256, 256, 128, dt, bn=128)
        run_gemm(1000, 320, 320, dt, bn=160, resid=True)
        run_gemm(4096, 1280, 1024, dt, bn=256)
        run_gemm(300, 2560, 320, dt, bn=256, geglu=True)
        run_gemm(512, 320, 960, dt, bn=160, split=640)
        run_gemm(4128, 640, 1024, dt, bn=0, bias=False)
        run_conv(2, 32, 64, 320, 320, dt)
        run_conv(2, 16, 32, 640, 640, dt, bn=128)
        run_conv(16, 4, 8, 1280, 1280, dt)
        run_conv(6, 4, 8, 128, 128, dt, bn=128)
        run_conv(1, 64, 128, 64, 256, dt, resid=False, bn=256)
        run_conv(2, 16, 32, 320, 320, dt, stride=2, temb=False, resid=False)
    dt = torch.bfloat16
    shapes = [("conv320@32x64", lambda bn: ops.conv3x3(x320, w320, bn=bn), 2 * 16 * 2048 * 320 * 2880, (160, 256)),
              ("conv640@16x32", lambda bn: ops.conv3x3(x640, w640, bn=bn), 2 * 16 * 512 * 640 * 5760, (160, 128, 256)),
              ("conv1280@8x16", lambda bn: ops.conv3x3(x1280, w1280, bn=bn), 2 * 16 * 128 * 1280 * 11520, (160, 128, 256)),
              ("conv1280@4x8", lambda bn: ops.conv3x3(x1280s, w1280, bn=bn), 2 * 16 * 32 * 1280 * 11520, (160, 128, 64)),
              ("geglu 32768x2560x320", lambda bn: ops.gemm(a320, wg, bias=bg, geglu=True, bn=bn), 2 * 32768 * 2560 * 320, (256, 128)),
              ("ffout 32768x320x1280", lambda bn: ops.gemm(a1280, wf, residual=r320, bn=bn), 2 * 32768 * 320 * 1280, (160,)),
              ("qkv 32768x960x320", lambda bn: ops.gemm(a320, wq, bn=bn), 2 * 32768 * 960 * 320, (160, 128)),
              ("toout 32768x320x320", lambda bn: ops.gemm(a320, wo, residual=r320, bn=bn), 2 * 32768 * 320 * 320, (160,)),
              ("gemm 8192^3", lambda bn: ops.gemm(a8k, w8k, bn=bn), 2 * 8192 ** 3, (256,))]
    x320 = torch.randn(16, 32, 64, 320, device=dev, dtype=dt); w320 = torch.randn(320, 2880, device=dev, dtype=dt)
    x640 = torch.randn(16, 16, 32, 640, device=dev, dtype=dt); w640 = torch.randn(640, 5760, device=dev, dtype=dt)
    x1280 = torch.randn(16, 8, 16, 1280, device=dev, dtype=dt); w1280 = torch.randn(1280, 11520, device=dev, dtype=dt)
    x1280s = torch.randn(16, 4, 8, 1280, device=dev, dtype=dt)
    a320 = torch.randn(32768, 320, device=dev, dtype=dt); wg = torch.randn(2560, 320, device=dev, dtype=dt); bg = torch.randn(2560, device=dev)
    a1280 = torch.randn(32768, 1280, device=dev, dtype=dt); wf = torch.randn(320, 1280, device=dev, dtype=dt); r320 = torch.randn(32768, 320, device=dev, dtype=dt)
    wq = torch.randn(960, 320, device=dev, dtype=dt); wo = torch.randn(320, 320, device=dev, dtype=dt)
    a8k = torch.randn(8192, 8192, device=dev, dtype=dt); w8k = torch.randn(8192, 8192, device=dev, dtype=dt)
    for name, fn, fl, bns in shapes:
        for bn in bns:
            for cg in (1, 2):
                if cg == 2 and bn < 128: continue
                set_cg(cg)
                ms = timeit(lambda: fn(bn))
                res.append(dict(op="perf", name=name, bn=bn, cg=cg, ms=ms, tflops=fl / ms / 1e9)); print(res[-1], flush=True)
    set_cg(0)
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
