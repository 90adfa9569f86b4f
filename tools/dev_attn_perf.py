import sys; sys.path.insert(0, ".")
import torch
from pcdms_b200 import ops, lib
L = lib.load(); dev = "cuda"; dt = torch.bfloat16
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
qkv = torch.randn(16 * 2048, 960, device=dev, dtype=dt)
q = torch.randn(16 * 2048, 320, device=dev, dtype=dt); kv = torch.randn(16 * 258, 640, device=dev, dtype=dt)
qkv2 = torch.randn(16 * 512, 1920, device=dev, dtype=dt)
for ns in (0, 200, 400, 600, 900, 1200, 2000):
    L.pcdm_set_attention_stagger_ns(ns)
    a = timeit(lambda: ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], 16, 5))
    b = timeit(lambda: ops.attention(q, kv[:, :320], kv[:, 320:], 16, 5))
    c = timeit(lambda: ops.attention(qkv2[:, :640], qkv2[:, 640:1280], qkv2[:, 1280:], 16, 10))
    print(f"stagger {ns:5d} ns: self2048 {a:7.1f} us ({85.9e3/a:6.1f} TF)  cross258 {b:6.1f} us  self512 {c:6.1f} us", flush=True)
