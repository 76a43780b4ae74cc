"""Turn an `ncu -i X.ncu-rep --page raw --csv` dump into the per-kernel table kept under profiles/
(duration, DRAM bytes, SM / issue / warp / L2 / L1 utilisation, L2 hit rate, registers).

    python profiles/ncu_table.py gpurun_out/X_raw.csv > profiles/X_ncu.md"""
import csv
import re
import sys

COLS = [("us", "gpu__time_duration.sum", 1.0), ("dram_rd_MB", "dram__bytes_read.sum", 1.0), ("dram_wr_MB", "dram__bytes_write.sum", 1.0),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
        ("warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0), ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("l1_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", 1.0), ("l2_hit", "lts__t_sector_hit_rate.pct", 1.0),
        ("regs", "launch__registers_per_thread", 1.0)]
UNIT = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print("| kernel | grid | " + " | ".join(c[0] for c in COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("pcnbr::", "")
    vals = []
    for _, metric, _ in COLS:
        i = col.get(metric)
        if i is None or r[i] == "":
            vals.append("-")
            continue
        v = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        vals.append(f"{v:.1f}")
    print(f"| `{name}` | {r[col['Grid Size']]} | " + " | ".join(vals) + " |")
