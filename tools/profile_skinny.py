"""ncu target: the swap-AB skinny GEMM on the NVILA-8B gate/up prefill shape (M = 279)."""
import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200 import ops
ops.ensure_workspace("cuda")
for (M, N, K) in [(279, 37888, 3584)]:
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K)
    for _ in range(3):
        ops.linear(x, w, swiglu=True, block_n=3000, static_w=True)
    torch.cuda.synchronize()
