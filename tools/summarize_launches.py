"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and mean time."""
import csv, re, sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((int(r["ID"]), name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
agg = OrderedDict()
for _, name, us, grid, blk in rows:
    key = f"{name} grid={grid}"
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {len(rows)} launches, {tot/1e3:.2f} ms total (cold-cache, serialised: compare shares)")
print(f"{'kernel':100s} {'n':>7s} {'total_ms':>10s} {'mean_us':>9s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{k[:100]:100s} {n:7d} {t/1e3:10.3f} {t/n:9.2f} {100*t/tot:6.2f}%")
