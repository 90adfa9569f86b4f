"""Replay exactly the conv3x3 launches of ONE UNet evaluation (BASELINE config 2: B = 16, 32x64 latents, bf16) — the
dominant kernel class of bench.py's `roofline` — for ncu:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/convs.csv python tools/profile_convs.py
    python tools/extract_traffic.py gpurun_out/convs.csv profiles/roofline_traffic.json

The calls (arguments and all) are recorded by wrapping pcdms_b200.ops.conv3x3 / conv3x3_up2x during one forward, then
issued again between cudaProfilerStart / Stop with the tensors of that forward still alive."""
import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
from pcdms_b200.unet import B200UNet2DConditionModel

dev, dt = "cuda", torch.bfloat16
m = B200UNet2DConditionModel(dtype=dt, device=dev, in_channels=9, class_embed_type="projection",
                             projection_class_embeddings_input_dim=1024)
m.load_state_dict(m.synthetic_state_dict(0))
B, h, w = 16, 32, 64
x9 = torch.randn(B, h, w, 64, device=dev).to(dt)
t = torch.tensor([981.0], device=dev)
kv = m.context_kv(torch.randn(B, 258, 1024, device=dev).to(dt))
cls = torch.randn(B, 1024, device=dev).to(dt)
pose = (0.1 * torch.randn(B, h, w, 320, device=dev)).to(dt)
calls = []
real = (ops.conv3x3, ops.conv3x3_up2x)


def rec(fn):
    def inner(*a, **k):
        calls.append((fn, a, k))
        return fn(*a, **k)
    return inner


ops.conv3x3, ops.conv3x3_up2x = rec(real[0]), rec(real[1])
m.forward_nhwc(x9, t, kv, cls, pose)
ops.conv3x3, ops.conv3x3_up2x = real
torch.cuda.synchronize()
print("conv launches per UNet evaluation:", len(calls))
for fn, a, k in calls:          # warm-up (lazy attribute set-up)
    fn(*a, **k)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for fn, a, k in calls:
    fn(*a, **k)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
