"""Summarise `ncu --page source --csv` (gzip ok): per kernel section, warp-stall samples by reason and the hottest SASS lines.
    python tools/ncu_stalls.py gpurun_out/X_source.csv.gz [kernel substring] [top N]"""
import csv, gzip, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 8
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
sections, cur = [], None
for row in csv.reader(f):
    if not row: continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; sections.append(cur)
    elif row[0] == "Address" and cur is not None:
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
seen = set()
for s in sections:
    short = s["name"].split("(")[0].replace("void ", "").replace("pcnbr::", "")
    if want not in s["name"]: continue
    h = {k: i for i, k in enumerate(s["hdr"])}
    stall_cols = [k for k in s["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    tot = {k: 0 for k in stall_cols}; total = 0; inst = 0
    lines = []
    for r in s["rows"]:
        n = int(r[h["# Samples"]] or 0); total += n; inst += int(r[h["Instructions Executed"]] or 0)
        for k in stall_cols: tot[k] += int(r[h[k]] or 0)
        lines.append((n, r[h["Source"]].strip(), r))
    key = (short, total)
    if key in seen: continue
    seen.add(key)
    print(f"== {short}: {total} samples, {inst/1e6:.2f} M warp-instructions")
    print("   " + ", ".join(f"{k[6:]} {100*v/max(1,total):.0f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7] if v))
    for n, src, r in sorted(lines, key=lambda t: -t[0])[:top]:
        why = sorted(((int(r[h[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
        print(f"   {100*n/max(1,total):5.1f}%  {src[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
