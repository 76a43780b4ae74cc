"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time
and share of the profiled region.  usage: python profiles/summarize.py gpurun_out/launches.csv [skip_launches]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        val_us = val / 1e3 if unit in ("nsecond", "ns") else (val if unit in ("usecond", "us") else val * 1e3)
        rows.append((int(r["ID"]), r["Kernel Name"], val_us))
rows = [r for r in rows if r[0] >= skip]
agg = defaultdict(lambda: [0, 0.0])
for _, k, us in rows:
    k = re.sub(r"\(.*", "", k)[:90]
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised: compare shares, not absolutes)")
print(f"{'share':>7} {'ms':>9} {'n':>5}  kernel")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{100 * us / tot:6.2f}% {us / 1e3:9.3f} {n:5d}  {k}")
mine = sum(us for k, (n, us) in agg.items() if "pcnbr" in k)
print(f"# libpcnbr kernels: {100 * mine / tot:.1f}% of the region")
