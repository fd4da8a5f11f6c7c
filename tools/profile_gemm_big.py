"""ncu target: the CTA-pair tcgen05 GEMM on a large problem (8192^3, the shape MEASURED_PEAKS.json's
cuBLAS figure is quoted on)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200 import ops
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
w = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16) / 90.0
for _ in range(3):
    ops.linear(x, w, static_w=True)
torch.cuda.synchronize()
