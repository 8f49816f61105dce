"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v * scale, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(r[1] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for name, us, *_ in rows:
    agg[name][0] += 1
    agg[name][1] += us
print(f"launches {len(rows)}  total {tot/1e3:.3f} ms (ncu per-launch times: cold cache, serialised -> compare SHARES)")
print(f"{'kernel':60s} {'count':>6s} {'total_ms':>10s} {'share':>7s} {'avg_us':>9s}")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {c:6d} {us/1e3:10.3f} {100*us/tot:6.1f}% {us/c:9.1f}")
