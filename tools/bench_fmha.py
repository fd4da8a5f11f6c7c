"""GPU micro-benchmark: tcgen05 FMHA one-tile (v1) vs two-tile (v2) kernels on the NVILA shapes (warm,
CUDA events).  Usage: python tools/bench_fmha.py"""
import sys, os, json, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200 import ops
torch.cuda.set_device(0)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
res = {}
for variant, tag in ((1, "v1"), (2, "v2"), (3, "v2_poly4")):
  for name, (B, S, H, Hkv, D, causal) in {"vit_1tile": (1, 1024, 16, 16, 72, False), "vit_64": (64, 1024, 16, 16, 72, False),
                                        "llm_279": (1, 279, 28, 4, 128, True), "llm_4k": (1, 4096, 28, 4, 128, True),
                                        "llm_16k": (1, 16470, 28, 4, 128, True),
                                        "llm_66k": (1, 65814, 28, 4, 128, True)}.items():
    q = torch.randn(B * S, H, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B * S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B * S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    o = torch.empty_like(q)
    us = timeit(lambda: ops.fmha(q, k, v, B=B, Sq=S, Sk=S, causal=causal, scale=D ** -0.5, out=o, variant=variant), reps=5 if S > 8000 or B > 8 else 20)
    flops = 4.0 * B * S * S * H * D * (0.5 if causal else 1.0)
    res[tag + ":" + name] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1)}
    print(tag, name, res[tag + ":" + name], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/bench_fmha.json").write_text(json.dumps(res, indent=1))
