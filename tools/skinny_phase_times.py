"""Profiling aid: per-CTA phase timings of the swap-AB skinny GEMM (single CTA vs CTA pair), read from
the kernel's own clock stamps (VILA_B200_GEMM_DEBUG).  Usage: python tools/skinny_phase_times.py"""
import json, math, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200 import ops

ops.ensure_workspace("cuda")
dbg = torch.zeros(16 * 4096, dtype=torch.int64, device="cuda")
os.environ["VILA_B200_GEMM_DEBUG"] = "%x" % dbg.data_ptr()
out = {}
shapes = {"llm_gu": (279, 37888, 3584), "llm_down": (279, 3584, 18944), "llm_qkv": (279, 4608, 3584)}
SWIGLU = "--swiglu" in sys.argv
if "--sweep" in sys.argv:  # tensor-pipe time per k-block as a function of the token count (UMMA N)
    shapes = {f"m{m}": (m, 18944, 3584) for m in (64, 128, 144, 192, 256, 288, 320, 384, 512)}
for name, (M, N, K) in shapes.items():
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K) for _ in range(3)]
    for cfg in (3000, 3001):
        for i in range(3):
            dbg.zero_()
            ops.linear(x, ws[i], block_n=cfg, static_w=True, swiglu=SWIGLU)
        torch.cuda.synchronize()
        d = dbg.view(-1, 16).cpu()
        d = d[d[:, 0] > 0]
        t0 = int(d[:, 0].min())
        lead = d[d[:, 3] > 0]  # CTAs that issued MMAs
        row = {
            "ctas": int(d.shape[0]), "kernel_us": round((int(d[:, 1].max()) - t0) / 1e3, 1),
            "cta_life_us_mean": round(float((d[:, 1] - d[:, 0]).float().mean()) / 1e3, 1),
            "start_us_p50_p100": [round(float((d[:, 0] - t0).float().quantile(q)) / 1e3, 1) for q in (0.5, 1.0)],
            "mma_loop_kcyc_mean": round(float(lead[:, 3].float().mean()) / 1e3, 1),
            "mma_wait_full_kcyc_mean": round(float(lead[:, 2].float().mean()) / 1e3, 1),
            "prod_wait_empty_kcyc_mean": round(float(d[:, 4].float().mean()) / 1e3, 1),
            "epi_wait_acc_kcyc_mean": round(float(d[:, 7].float().mean()) / 1e3, 1),
            "epilogue_kcyc_mean": round(float(d[:, 5].float().mean()) / 1e3, 1),
            "sms_used": int(d[:, 6].unique().numel()),
            "epi_kcyc[phase1,clusterbar,src,compute,bar,out]": [round(float(d[:, k].float().mean()) / 1e3, 2) for k in (8, 9, 10, 11, 12, 13)],
        }
        nkb_cta = (K + 63) // 64 / max(1, row["ctas"] // ((N + 127) // 128))
        row["mma_busy_cyc_per_kblock"] = round((row["mma_loop_kcyc_mean"] - row["mma_wait_full_kcyc_mean"]) * 1e3 / nkb_cta)
        out[f"{name}/{cfg}"] = row
        print(name, cfg, row, flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/skinny_phase_times.json").write_text(json.dumps(out, indent=1))
