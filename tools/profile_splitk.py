import sys
sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops
ops.ensure_workspace("cuda")
dev="cuda"; dt=torch.bfloat16
x = torch.randn(16, 4, 8, 1280, device=dev, dtype=dt); w = torch.randn(1280, 11520, device=dev, dtype=dt)
b = torch.randn(1280, device=dev); r = torch.randn(16, 4, 8, 1280, device=dev, dtype=dt)
for _ in range(2):
    ops.conv3x3(x, w, bias=b, residual=r)
    ops.conv3x3(x, w, bias=b, residual=r, bn=64)
    ops.conv3x3(x, w, bias=b, residual=r, bn=160)
torch.cuda.synchronize(); print("ok")
