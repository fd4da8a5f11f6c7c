"""GPU micro-benchmark: tcgen05 GEMM configurations on the NVILA-8B prefill / ViT shapes, next to
torch.matmul (cuBLAS) as the library yardstick.  Usage: python tools/bench_gemm.py"""
import sys, json, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200 import ops

torch.cuda.set_device(0)
ops.ensure_workspace("cuda")
shapes = {
    "vit_qkv": (1024, 3456, 1152), "vit_out": (1024, 1152, 1152), "vit_fc1": (1024, 4304, 1152),
    "vit_fc2": (1024, 1152, 4304), "llm_qkv": (279, 4608, 3584), "llm_o": (279, 3584, 3584),
    "llm_gu": (279, 37888, 3584), "llm_down": (279, 3584, 18944), "proj1": (256, 3584, 4608),
    "big": (8192, 8192, 8192),
    # mid-size: batched ViT (16 tiles) and long-context LLM prefill
    "mid_vit16_qkv": (16384, 3456, 1152), "mid_vit16_fc2": (16384, 1152, 4304),
    "mid_vit4_fc1": (4096, 4304, 1152), "mid_llm2k_o": (2048, 3584, 3584),
    "mid_llm2k_gu": (2048, 37888, 3584), "mid_llm1k_down": (1280, 3584, 18944),
    # batched video / SP shapes (banded rasterisation)
    "vit64_qkv": (65536, 3456, 1152), "vit64_fc1": (65536, 4304, 1152), "sp8_gu": (8448, 37888, 3584),
}
if len(sys.argv) > 1:  # optional name filter, e.g. "llm"
    shapes = {k: v for k, v in shapes.items() if any(a in k for a in sys.argv[1:])}
res = {}
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for name, (M, N, K) in shapes.items():
    ncopies = 1 if name == "big" else max(2, min(12, int(1.5e9 // (N * K * 2))))
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K) for _ in range(ncopies)]
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    idx = [0]
    def nxt():
        idx[0] = (idx[0] + 1) % ncopies
        return ws[idx[0]]
    us = timeit(lambda: torch.matmul(x, nxt().t(), out=out), reps=5 if name == "big" else 20)
    row = {"cublas_us": round(us, 2), "cublas_tflops": round(2 * M * N * K / us / 1e6, 1)}
    for cfg in (None, 2128, 2256, 4128, 4256) + ((3000, 3001) if M <= 512 else ()) + ((5128,) if ((M + 127) // 128) * ((N + 127) // 128) <= 74 else ()):
        try:
            us = timeit(lambda: ops.linear(x, nxt(), out=out, block_n=cfg, static_w=True), reps=5 if name == "big" else 20)
            row[str(cfg)] = round(us, 2)
        except Exception as e:
            row[str(cfg)] = str(e)[:40]
    res[name] = row
    print(name, (M, N, K), row, flush=True)
    del ws
    torch.cuda.empty_cache()
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/bench_gemm.json").write_text(json.dumps(res, indent=1))
