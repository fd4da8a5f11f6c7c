"""One launch of the two-tile FMHA kernel at the 64-frame video shape (causal GQA, d=128, S=16470, paged)
inside a cudaProfilerStart/Stop range — the target of `tools/run_gpu_r2.sh ncu_one`."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from vila_b200 import ops

S, Hq, Hkv, D = 16470, 28, 4, 128
npg = (S + 127) // 128
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(S, Hq, D, device="cuda", generator=g).to(torch.bfloat16)
kp = torch.randn(npg, 128, Hkv, D, device="cuda", generator=g).to(torch.bfloat16)
vp = torch.randn(npg, 128, Hkv, D, device="cuda", generator=g).to(torch.bfloat16)
pt = torch.arange(npg, dtype=torch.int32, device="cuda")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(2):
    ops.fmha(q, kp, vp, B=1, Sq=S, Sk=S, causal=True, scale=D ** -0.5, page_table=pt, variant=variant)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.fmha(q, kp, vp, B=1, Sq=S, Sk=S, causal=True, scale=D ** -0.5, page_table=pt, variant=variant)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
