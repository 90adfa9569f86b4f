"""Where a K/V block's time goes inside the two-tile attention kernel: CTA 0 stamps clock64() at every hand-shake of its
first 96 blocks per tile (experiment build, pcdm_set_attention_trace); this prints the per-block intervals, averaged over
the steady-state blocks of the first items, for the softmax warp of TMEM quadrant 0 of each tile and for the MMA warp.

    python tools/dev_attn_trace.py [Sq] [Skv]          (default 2048 2048; B 16, 5 heads, bf16)
"""
import sys
sys.path.insert(0, ".")
import tools._explib  # noqa: F401
import ctypes as C
import torch
from pcdms_b200 import ops, lib

L = lib.load()
Sq = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Skv = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B, heads = 16, 5
N = 96
dev = "cuda"
q = torch.randn(B * Sq, heads * 64, device=dev).to(torch.bfloat16)
k = torch.randn(B * Skv, heads * 64, device=dev).to(torch.bfloat16)
v = torch.randn(B * Skv, heads * 64, device=dev).to(torch.bfloat16)
out = torch.empty_like(q)
for _ in range(3):
    ops.attention(q, k, v, B, heads, out=out)
torch.cuda.synchronize()
trace = torch.zeros(2 * N * 16 + 32 * 8, dtype=torch.int64, device=dev)
L.pcdm_set_attention_trace.argtypes = [C.c_void_p]
L.pcdm_set_attention_trace(C.c_void_p(trace.data_ptr()))
ops.attention(q, k, v, B, heads, out=out)
torch.cuda.synchronize()
L.pcdm_set_attention_trace(C.c_void_p(0))
items = trace.cpu()[2 * N * 16:].view(32, 8)
tr = trace.cpu()[:2 * N * 16].view(2, N, 16)
n_kv = (Skv + 127) // 128
t0 = int(tr[tr > 0].min())
names = ["S landed", "chunk0 in regs", "chunk0 exp done", "PV(n-1) retired seen", "chunk1 done", "chunk2 done", "chunk3 done",
         "p_full arrive", "MMA: before wait", "MMA: wait passed", "MMA: QK(n) issued", "MMA: PV(n-1) issued"]
print(f"Sq {Sq} Skv {Skv}: {n_kv} K/V blocks per item; stamps relative to the CTA's first one, in cycles")
for t in range(2):
    print(f"--- tile {t}: first 2 items, per block: " + " | ".join(f"[{i}] {nm}" for i, nm in enumerate(names)))
    for n in range(min(N, 2 * n_kv + 2)):
        row = tr[t, n]
        if int(row[0]) == 0 and int(row[8]) == 0:
            continue
        print(f"n {n:3d}: " + " ".join(f"{(int(x) - t0) if int(x) else -1:8d}" for x in row[:12]))
# steady state: blocks 2 .. n_kv - 2 of the first two items
for t in range(2):
    sel = [n for n in range(N) if 2 <= (n % n_kv) <= n_kv - 2 and n + 1 < N and int(tr[t, n, 0]) and int(tr[t, n + 1, 0])]
    if not sel:
        continue
    a = tr[t, sel].double()
    nxt = tr[t, [n + 1 for n in sel]].double()
    period = (nxt[:, 0] - a[:, 0]).mean().item()
    print(f"tile {t}: steady-state block period {period:.0f} cycles over {len(sel)} blocks; mean intervals:")
    segs = [("S landed -> chunk0 in regs", 0, 1), ("chunk0 exp", 1, 2), ("wait PV(n-1) retired", 2, 3), ("chunk1", 3, 4),
            ("chunk2", 4, 5), ("chunk3", 5, 6), ("P stores complete + arrive", 6, 7)]
    for nm, i, j in segs:
        print(f"   {nm:32s} {(a[:, j] - a[:, i]).mean().item():8.0f}")
    print(f"   {'arrive -> next S landed (bubble)':32s} {(nxt[:, 0] - a[:, 7]).mean().item():8.0f}")
    print(f"   MMA warp, block n+1: arrive(n) -> wait passed {(nxt[:, 9] - a[:, 7]).mean().item():6.0f}, "
          f"wait passed -> QK issued {(nxt[:, 10] - nxt[:, 9]).mean().item():6.0f}, QK issued -> S landed "
          f"{(nxt[:, 0] - nxt[:, 10]).mean().item():6.0f}, QK issued -> PV issued {(nxt[:, 11] - nxt[:, 10]).mean().item():6.0f}, "
          f"MMA idle before the wait {(nxt[:, 9] - nxt[:, 8]).mean().item():6.0f}")

print("--- per item (cycles since the CTA's first stamp): MMA [0] item start [1] Q landed [2] first K/V landed | TMA [3] Q buffer free "
      "[4] first K/V slot free | softmax tile 0 [5] last PV retired [6] O written [7] first S landed")
for i in range(32):
    if int(items[i].max()) == 0:
        break
    print(f"item {i:2d}: " + " ".join(f"{(int(x) - t0) if int(x) else -1:8d}" for x in items[i]))
