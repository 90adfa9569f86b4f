"""Mainloop pace of the implicit GEMM: conv 320->320 / 640->320 @32x64 (B=16), GEMM M32768 N320 K1280 and 8192^3, per tile
width and CTA-group, GPU-bound back-to-back launches.  (Round-1 finding, with loads / MMAs switched off one at a time:
the loop was paced by the issuing THREAD — ~330 cycles of barrier handshake + ~80 cycles per tcgen05.mma issued from a
`lane == 0` branch — not by loads or MMAs; see profiles/r1_mainloop_experiment.md.)"""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch
from pcdms_b200 import ops, lib
dev = "cuda"; dt = torch.bfloat16
L = lib.load()
x = torch.randn(16, 32, 64, 320, device=dev, dtype=dt); w = (torch.randn(320, 2880, device=dev) / 54).to(dt)
x6 = torch.randn(16, 32, 64, 640, device=dev, dtype=dt); w6 = (torch.randn(320, 5760, device=dev) / 76).to(dt)
a = torch.randn(32768, 1280, device=dev, dtype=dt); wg = (torch.randn(320, 1280, device=dev) / 36).to(dt)
a8 = torch.randn(8192, 8192, device=dev, dtype=dt); w8 = (torch.randn(8192, 8192, device=dev) / 90).to(dt)
yo = torch.empty(16, 32, 64, 320, device=dev, dtype=dt); go = torch.empty(32768, 320, device=dev, dtype=dt)
g8 = torch.empty(8192, 8192, device=dev, dtype=dt)

def timeit(fn, reps=12):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(10_000_000)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

cases = [("conv 320->320 (45 kb)", lambda bn: ops.conv3x3(x, w, out=yo, bn=bn), 45),
         ("conv 640->320 (90 kb)", lambda bn: ops.conv3x3(x6, w6, out=yo, bn=bn), 90),
         ("gemm M32768 N320 K1280 (20 kb)", lambda bn: ops.gemm(a, wg, out=go, bn=bn), 20),
         ("gemm 8192^3 (128 kb)", lambda bn: ops.gemm(a8, w8, out=g8, bn=bn), 128)]
for cg in (1, 2):
    L.pcdm_set_gemm_cta_group(cg)
    for name, fn, nkb in cases:
        for bn in (160, 256) if "8192^3" not in name else (256,):
            row = []
            row.append(f"{timeit(lambda: fn(bn)):7.1f} us")
            print(f"cg{cg} bn{bn:3d} {name:32s} " + "  ".join(row), flush=True)
L.pcdm_set_gemm_cta_group(0)
