"""GPU micro-benchmark: fused q/k/v projection + RoPE + KV append vs the two separate kernels
(NVILA-8B prefill chunk, M = 279).  Usage: python tools/bench_qkv_rope.py"""
import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200 import ops

torch.cuda.set_device(0)
M, K, Hq, Hkv, D = 279, 3584, 28, 4, 128
N = (Hq + 2 * Hkv) * D
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K) for _ in range(8)]
b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
pos = torch.arange(M, dtype=torch.int32, device="cuda")
inv_freq = (1.0 / (1e6 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))).cuda()
table = torch.arange(4, dtype=torch.int32, device="cuda")
kp = torch.zeros(4, 128, Hkv, D, dtype=torch.bfloat16, device="cuda")
vp = torch.zeros_like(kp)
i = [0]
def nxt():
    i[0] = (i[0] + 1) % len(ws)
    return ws[i[0]]
def timeit(fn, reps=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
def two():
    q = ops.linear(x, nxt(), b, static_w=True)
    ops.rope_kv_append(q, pos, Hq, Hkv, D, inv_freq, kp, vp, table, 0)
tab = ops.rope_table(pos, D, inv_freq)
def fused():
    ops.linear_qkv_rope(x, nxt(), b, tab, Hq, Hkv, D, kp, vp, table, 0, static_w=True)
print("linear only      us", round(timeit(lambda: ops.linear(x, nxt(), b, static_w=True)), 2))
print("linear + rope    us", round(timeit(two), 2))
print("fused            us", round(timeit(fused), 2))
