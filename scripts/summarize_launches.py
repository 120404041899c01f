"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v, r.get("Grid Size", ""), r.get("Block Size", "")))
agg = defaultdict(lambda: [0, 0.0])
for n, v, *_ in rows:
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v, *_ in rows)
print(f"{len(rows)} launches, total {tot/1e3:.1f} us")
print(f"{'kernel':70s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {c:5d} {v/1e3:10.1f} {v/c/1e3:8.2f} {100*v/tot:5.1f}%")
