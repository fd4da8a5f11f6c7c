"""CUDA-event timing (cold L2: 256 MB scratch write between launches, median of 20) of the data-movement kernels
at their NVILA shapes: dynamic-S2 merge of 35 tiles, TSP pooling of 64 frames, chessboard merge, RoPE + KV append
(table flavour), space-to-depth.  Writes gpurun_out/datamove.json."""
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vila_b200 import ops  # noqa: E402

dev = torch.device("cuda")
ops.ensure_workspace(dev)
g = torch.Generator(device="cuda").manual_seed(0)
bf = torch.bfloat16
PEAK = 6574.5
pk = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
if pk.exists():
    PEAK = float(json.loads(pk.read_text())["hbm_gbs"])


def rn(*shape):
    return torch.randn(*shape, device=dev, generator=g).to(bf)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=20):
    fn()
    ts = []
    for i in range(n):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


C, Hd, Hq, Hkv, D = 1152, 3584, 28, 4, 128
rows = []


def case(label, nbytes, fn):
    us = timed(fn)
    rows.append({"kernel": label, "algorithmic_bytes": nbytes, "us": round(us, 2), "gbs": round(nbytes / us / 1e3, 1),
                 "frac": round(nbytes / us / 1e3 / PEAK, 3)})
    print(rows[-1], flush=True)


tiles = rn(35, 1024, C)
case("s2_merge 35 tiles -> (5,6) x 3456", 2 * (35 * 1024 * C + 30 * 1024 * 3 * C),
     lambda: ops.s2_merge(tiles, [1, 2, 5], [1, 2, 6], 5, 6))
vfe = rn(64, 16, 16, Hd)
case("tsp_pool 64x16x16x3584 (8,1,1)", 2 * (64 + 8) * 256 * Hd, lambda: ops.tsp_pool(vfe, 8, 1, 1))
case("tsp_pool 64x16x16x3584 (4,2,2)", 2 * 64 * 256 * Hd + 2 * 16 * 64 * Hd, lambda: ops.tsp_pool(vfe, 4, 2, 2))
ptiles = rn(30, 256, Hd)
case("chessboard_merge 30x256x3584", 2 * 2 * 30 * 256 * Hd, lambda: ops.chessboard_merge(ptiles, 5, 6))
feat = rn(64, 1024, C)
case("space_to_depth 64x1024x1152 r=2", 2 * 2 * 64 * 1024 * C, lambda: ops.space_to_depth(feat, 32, 32, 2))
Sv = 64 * 257 + 22
npg = (Sv + 127) // 128
kpv, vpv = rn(npg, 128, Hkv, D), rn(npg, 128, Hkv, D)
ptv = torch.arange(npg, dtype=torch.int32, device=dev)
qkv_big = rn(Sv, (Hq + 2 * Hkv) * D)
inv = (1.0 / (1e6 ** (torch.arange(0, D, 2, device=dev, dtype=torch.float32) / D))).contiguous()
posv = torch.arange(Sv, dtype=torch.int32, device=dev)
tbl = ops.rope_table(posv, D, inv)
case("rope_kv_append_table S=16470", 2 * (2 * Sv * (Hq + Hkv) * D + 2 * Sv * Hkv * D),
     lambda: ops.rope_kv_append_table(qkv_big, tbl, Hq, Hkv, D, kpv, vpv, ptv, 0))
out = Path("gpurun_out")
if out.is_dir():
    (out / "datamove.json").write_text(json.dumps({"hbm_peak_gbs": PEAK, "note": "CUDA events, cold L2, median of 20",
                                                   "rows": rows}, indent=1))
