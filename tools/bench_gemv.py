"""GPU micro-benchmark: decode GEMV variants on the NVILA-8B shapes (weights rotated over 8+ copies
so every launch streams from HBM).  Usage: python tools/bench_gemv.py"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200 import ops

torch.cuda.set_device(0)
peak = 6574.5
shapes = {"qkv": (4608, 3584, False), "o_proj": (3584, 3584, False), "gate_up": (37888, 3584, True),
          "down": (3584, 18944, False), "lm_head": (152064, 3584, False)}
res = {}
for name, (N, K, swiglu) in shapes.items():
    ncopies = max(3, int(2.5e9 // (N * K * 2)))
    ncopies = min(ncopies, 40)
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(ncopies)]
    x = torch.randn(K, device="cuda", dtype=torch.bfloat16)
    nw = torch.ones(K, device="cuda", dtype=torch.bfloat16)
    y = torch.empty(N // 2 if swiglu else N, device="cuda", dtype=torch.bfloat16)
    for variant in (0, 1):
        for w in ws:
            ops.gemv(x, w, norm_w=nw, swiglu=swiglu, out=y, static_w=True, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            for w in ws:
                ops.gemv(x, w, norm_w=nw, swiglu=swiglu, out=y, static_w=True, variant=variant)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * ncopies)
        gbs = N * K * 2 / us / 1e3
        res[f"{name}/v{variant}"] = {"us": round(us, 2), "GBps": round(gbs, 1), "frac": round(gbs / peak, 3)}
        print(name, "variant", variant, f"{us:8.2f} us  {gbs:7.1f} GB/s  {gbs/peak:.3f}", flush=True)
    del ws
    torch.cuda.empty_cache()
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/bench_gemv.json").write_text(json.dumps(res, indent=1))
