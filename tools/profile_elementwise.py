"""K6 — the scheduler / step elementwise kernels at a size where bandwidth is visible, for `ncu --set full` and for an
event-timed GB/s figure: DDIMScheduler.step and DDPMScheduler.add_noise on 256 x 4 x 64 x 128 fp16 latents (16.8 MB per
tensor: 50 MB of algorithmic traffic per call), and the fused CFG + DDIM step at the bench shape (n = 8, 32 x 64) and at
n = 256.  Prints achieved GB/s (algorithmic bytes / event time) against MEASURED_PEAKS.json."""
import json
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
from pcdms_b200.scheduler import B200DDIMScheduler, B200DDPMScheduler

dev = "cuda"
try:
    HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    HBM = None


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)   # ~10 ms of spin: the CPU queues every launch behind it (GPU-bound timing)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps   # us


res = {}
sch = B200DDIMScheduler()
sch.set_timesteps(50)
# three rotating buffer sets (3 x 50 MB > 126 MB L2) so that every call streams from HBM
sets = [(torch.randn(256, 4, 64, 128, device=dev).half(), torch.randn(256, 4, 64, 128, device=dev).half()) for _ in range(3)]
ts = torch.randint(0, 1000, (256,), device=dev)
it = [0]


def step():
    e, s = sets[it[0] % 3]
    it[0] += 1
    return sch.step(e, 501, s, return_dict=False)[0]


def noise():
    e, s = sets[it[0] % 3]
    it[0] += 1
    return B200DDPMScheduler().add_noise(s, e, ts)


nbytes = 3 * sets[0][0].numel() * 2
for name, fn in (("ddim_step fp16 256x4x64x128", step), ("add_noise fp16 256x4x64x128", noise)):
    us = timed(fn)
    res[name] = {"us": us, "algorithmic_bytes": nbytes, "gbs": nbytes / us / 1e3, "frac_of_hbm_peak": (nbytes / us / 1e3 / HBM) if HBM else None}
    print(name, json.dumps(res[name]), flush=True)

for n in (8, 256):
    h, w = 32, 64
    lat = torch.randn(n, 4, h, w, device=dev)
    eps = torch.randn(2 * n, h, w, 32, device=dev)
    x9 = torch.zeros(2 * n, h, w, 64, device=dev, dtype=torch.bfloat16)
    coef = sch.coefficient_table(dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)

    def fused():   # the step counter walks the 50-row coefficient table: 23 calls per timing stay inside it
        ops.cfg_ddim_step(eps, lat, x9, coef, counter, 2.0)
    counter.zero_()
    us = timed(fused)
    b = n * 4 * h * w * 28          # 28 B per latent element (2 eps reads, latents r/w, 2 x9 writes)
    res[f"cfg_ddim_step n={n}"] = {"us": us, "algorithmic_bytes": b, "gbs": b / us / 1e3}
    print(f"cfg_ddim_step n={n}", json.dumps(res[f"cfg_ddim_step n={n}"]), flush=True)
json.dump(res, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/elementwise.json", "w"), indent=1)
