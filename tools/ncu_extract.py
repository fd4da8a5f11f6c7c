"""Extract the headline metrics from an .ncu-rep (read here, no GPU needed) as CSV-ish text."""
import csv, subprocess, sys, io, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
tens = [i for i, h in enumerate(hdr) if "tensor" in h and "pct" in h]
out = []
for r in rows[2:]:
    d = {w: (r[i] + (" " + units[i] if units[i] else "")) for w, i in idx}
    for i in tens[:6]:
        d[hdr[i]] = r[i]
    out.append(d)
print(json.dumps(out, indent=1))
