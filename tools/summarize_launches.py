"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean."""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
seq = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"void |vb::|\(anonymous namespace\)::", "", short)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    grid = r.get("Grid Size", "")
    key = short + " grid=" + grid
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += us
    seq.append((short, grid, us))
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {len(seq)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:5d} mean={t/n:8.2f} us  {k}")
if len(sys.argv) > 2:
    for s in seq[: int(sys.argv[2])]:
        print(s)
