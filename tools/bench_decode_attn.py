"""GPU micro-benchmark: decode attention at the headline and video contexts, every kernel flavour,
live (CUDA events over 28 layer-sized KV pools back to back, PDL on).  Usage: python tools/bench_decode_attn.py"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from vila_b200 import ops

torch.cuda.set_device(0)
Hq, Hkv, D, L = 28, 4, 128, 28
inv = (1.0 / (1e6 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))).cuda()
qkv = torch.randn((Hq + 2 * Hkv) * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(Hq * D, device="cuda", dtype=torch.bfloat16)
ws = torch.zeros(Hkv * 64 * (Hq // Hkv) * (D + 2), device="cuda")
cnt = torch.zeros(Hkv, dtype=torch.int32, device="cuda")
o_part = torch.zeros(64 * Hq * D, device="cuda")
lse = torch.zeros(64 * Hq, device="cuda")
res = {}


def timeit(fn, reps=10):
    for li in range(L):
        fn(li)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for li in range(L):
            fn(li)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * L)


for ctx in (279, 1000, 2000, 4000, 16448, 65814):
    npg = max(8, (ctx + 1 + 127) // 128 + 1)
    pools = torch.randn(L, 2, npg, 128, Hkv, D, device="cuda").to(torch.bfloat16)
    pt = torch.arange(npg, dtype=torch.int32, device="cuda")
    pos = torch.tensor([ctx], dtype=torch.int32, device="cuda")
    kvb = 2 * 2 * (ctx + 1) * Hkv * D
    row = {}
    cands = ([0] if ctx < 1024 else []) + [8, 16, 37]
    for sp in cands:
        us = timeit(lambda li: ops.decode_attention(qkv, pos, pools[li, 0], pools[li, 1], pt, out, ws, cnt, inv,
                                                    Hq, Hkv, D, sp, D ** -0.5))
        row["simt_splits_%d" % sp] = {"us": round(us, 2), "gbs": round(kvb / us / 1e3, 1)}
    if ctx > 1024:
        pages = (ctx + 1 + 127) // 128
        for per_head in (18, 37):
            pps = (pages + per_head - 1) // per_head
            sp = (pages + pps - 1) // pps
            us = timeit(lambda li: ops.decode_attention_split(qkv, pos, pools[li, 0], pools[li, 1], pt, out,
                                                              o_part, lse, inv, Hq, Hkv, D, sp, pps * 128, D ** -0.5))
            row["fmha_split_%dx%d" % (sp, pps * 128)] = {"us": round(us, 2), "gbs": round(kvb / us / 1e3, 1)}
            us = timeit(lambda li: ops.decode_attention_split(qkv, pos, pools[li, 0], pools[li, 1], pt, out,
                                                              o_part, lse, inv, Hq, Hkv, D, sp, pps * 128, D ** -0.5,
                                                              counters=cnt))
            row["fmha_split_fused_combine_%dx%d" % (sp, pps * 128)] = {"us": round(us, 2), "gbs": round(kvb / us / 1e3, 1)}
    res[str(ctx)] = row
    print(ctx, json.dumps(row), flush=True)
    del pools
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/bench_decode_attn.json").write_text(json.dumps(res, indent=1))
