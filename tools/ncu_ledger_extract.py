"""Read gpurun_out/ledger.ncu-rep (made by `tools/run_gpu_r2.sh ledger`) here — no GPU needed — and
write profiles/r02_ncu_ledger.json + a markdown table: per kernel launch the ncu duration (cold L2,
serialised), DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), DRAM / tensor-pipe
utilisation, registers, shared memory, and beside them the ALGORITHMIC bytes / flops of the entry
(tools/ncu_ledger.py) with the achieved GB/s or TFLOP/s and the roofline fraction."""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out/ledger_raw.csv")
labels = json.loads((rep.parent / "ledger_labels.json").read_text())
peaks = {"hbm_gbs": 6574.5, "bf16_tflops": 1724.0}
pk = ROOT / "MEASURED_PEAKS.json"
if pk.exists():
    d = json.loads(pk.read_text())
    peaks = {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"])}

if rep.suffix == ".csv":
    raw = rep.read_text()
else:
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name, default=None):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def to_bytes(r, name):
    v = num(r, name)
    if v is None:
        return None
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(r, name):
    v = num(r, name)
    if v is None:
        return None
    u = units[col[name]].lower()
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "second": 1e6}.get(u, 1e-3)


launches = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("vb::(anonymous namespace)::", "")
    launches.append({
        "kernel": short[:90], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
        "us": to_us(r, "gpu__time_duration.sum"),
        "dram_bytes": (to_bytes(r, "dram__bytes_read.sum") or 0) + (to_bytes(r, "dram__bytes_write.sum") or 0),
        "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pipe_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                               num(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")),
        "sm_busy_pct": num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_pct": num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "regs": num(r, "launch__registers_per_thread"),
        "smem_dyn": num(r, "launch__shared_mem_per_block_dynamic"),
        "achieved_occupancy_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
    })

# map launches to ledger entries in order; an entry may own several launches (label says "[3 launches")
out, i = [], 0
for e in labels:
    mm = re.search(r"\[(\d+) launches", e["label"])
    n = int(mm.group(1)) if mm else 1
    mine = launches[i:i + n]
    i += n
    if not mine:
        continue
    main = max(mine, key=lambda x: x["us"] or 0)
    us = sum(x["us"] or 0 for x in mine)
    row = dict(e)
    row.update({"launches": mine, "ncu_us_cold": round(us, 2), "dram_bytes": int(sum(x["dram_bytes"] for x in mine))})
    if e["algorithmic_flops"] and e["algorithmic_flops"] / max(e["algorithmic_bytes"], 1) > 200:
        tf = e["algorithmic_flops"] / us / 1e6
        row.update({"bound": "tensor", "achieved_tflops": round(tf, 1), "frac": round(tf / peaks["bf16_tflops"], 3)})
    else:
        gbs = e["algorithmic_bytes"] / us / 1e3
        row.update({"bound": "hbm", "achieved_gbs": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 3)})
    row["traffic_over_algorithmic"] = round(row["dram_bytes"] / max(e["algorithmic_bytes"], 1), 3)
    row["main_kernel"] = main["kernel"]
    row["tensor_pipe_pct"], row["dram_pct"], row["regs"] = main["tensor_pipe_pct"], main["dram_pct"], main["regs"]
    out.append(row)
if i != len(launches):
    print(f"WARNING: {len(launches)} launches in the report, {i} consumed by {len(labels)} entries", file=sys.stderr)

dst = ROOT / "profiles"
(dst / "r02_ncu_ledger.json").write_text(json.dumps({"peaks": peaks, "note": "ncu --set full --clock-control none, one launch per "
                                                     "entry, cold L2 (256 MB flush before each), serialised (no PDL overlap)",
                                                     "entries": out}, indent=1))
lines = ["| kernel (shape) | ncu us (cold) | algorithmic | achieved | frac of peak | DRAM bytes / algorithmic | tensor pipe % | DRAM % | regs |",
         "|---|---|---|---|---|---|---|---|---|"]
for r in out:
    alg = ("%.1f GF" % (r["algorithmic_flops"] / 1e9)) if r["bound"] == "tensor" else ("%.2f MB" % (r["algorithmic_bytes"] / 1e6))
    ach = ("%.0f TF/s" % r["achieved_tflops"]) if r["bound"] == "tensor" else ("%.0f GB/s" % r["achieved_gbs"])
    lines.append("| %s | %.1f | %s | %s | %.2f (%s) | %.2f | %s | %s | %s |" % (
        r["label"], r["ncu_us_cold"], alg, ach, r["frac"], r["bound"], r["traffic_over_algorithmic"],
        "-" if r["tensor_pipe_pct"] is None else "%.0f" % r["tensor_pipe_pct"],
        "-" if r["dram_pct"] is None else "%.0f" % r["dram_pct"], "-" if r["regs"] is None else "%d" % r["regs"]))
(dst / "r02_ncu_ledger.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
