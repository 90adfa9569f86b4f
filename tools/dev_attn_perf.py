"""Dev check (GPU): attention kernel — share of exp2 on the FMA pipe (0..3 of 4 column pairs): correctness vs an fp32
softmax reference and timing."""
import sys; sys.path.insert(0, ".")
import tools._explib  # noqa: F401  (experiment build: pcdm_set_* hooks)
import torch
from pcdms_b200 import ops, lib
L = lib.load(); dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    torch.cuda._sleep(4_000_000)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def ref(q, k, v, B, heads):
    Sq, Skv = q.shape[0] // B, k.shape[0] // B
    qh = q.float().view(B, Sq, heads, 64).transpose(1, 2); kh = k.float().view(B, Skv, heads, 64).transpose(1, 2)
    vh = v.float().view(B, Skv, heads, 64).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B * Sq, heads * 64)
torch.manual_seed(0)
for dt in (torch.float16, torch.bfloat16):
    for (B, heads, Sq, Skv) in [(2, 5, 2048, 2048), (2, 10, 512, 512), (2, 5, 2048, 258), (1, 2, 384, 300), (1, 5, 256, 95), (3, 20, 128, 128)]:
        C = heads * 64
        q = torch.randn(B * Sq, C, device=dev).to(dt); k = torch.randn(B * Skv, C, device=dev).to(dt); v = torch.randn(B * Skv, C, device=dev).to(dt)
        want = ref(q, k, v, B, heads)
        for var in (0, 1):
            L.pcdm_set_attention_poly(var)
            got = ops.attention(q, k, v, B, heads).float()
            err = (got - want).abs().max().item()
            print(f"{str(dt)[6:]:9s} B{B} h{heads} Sq{Sq} Skv{Skv} fma-exp2 {var}: max abs err {err:.2e} nan {bool(torch.isnan(got).any())}", flush=True)
dt = torch.bfloat16
qkv = torch.randn(16 * 2048, 960, device=dev, dtype=dt)
q = torch.randn(16 * 2048, 320, device=dev, dtype=dt); kv = torch.randn(16 * 258, 640, device=dev, dtype=dt)
qkv2 = torch.randn(16 * 512, 1920, device=dev, dtype=dt)
for var in (0, 1):
    L.pcdm_set_attention_poly(var)
    a = timeit(lambda: ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], 16, 5))
    b = timeit(lambda: ops.attention(q, kv[:, :320], kv[:, 320:], 16, 5))
    c = timeit(lambda: ops.attention(qkv2[:, :640], qkv2[:, 640:1280], qkv2[:, 1280:], 16, 10))
    print(f"fma-exp2 {var}: self2048 {a:7.1f} us ({85.9e3/a:6.1f} TF)  cross258 {b:6.1f} us  self512 {c:6.1f} us", flush=True)
L.pcdm_set_attention_poly(0)
