#!/usr/bin/env python
"""Per-source-line stall summary from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: tools/ncu_lines.py src.csv [lo hi]   (optional line range of fjsp_core.cuh to aggregate)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
sections, cur = [], None
for i, r in enumerate(rows):
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0] not in ("Function Name",):
        cur["rows"].append(r)
tot_all = 0
lines = []
for s in sections:
    h = s["hdr"]
    col = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    for r in s["rows"]:
        if not r[0]:
            continue   # SASS row
        try:
            ns = int(r[col["# Samples"]]); ie = int(r[col["Instructions Executed"]])
        except ValueError:
            continue
        st = {n: int(r[col[n]] or 0) for n in stalls}
        lines.append((s["file"].split("/")[-1], int(r[0]), r[1].strip(), ns, ie, st))
        tot_all += ns
print("total samples", tot_all, "total warp instructions", sum(l[4] for l in lines))
sel = [l for l in lines if l[0] == "fjsp_core.cuh" and lo <= l[1] <= hi] if len(sys.argv) > 3 else lines
agg = {}
for l in sel:
    for k, v in l[5].items():
        agg[k] = agg.get(k, 0) + v
ns = sum(l[3] for l in sel)
print("selected samples %d (%.1f%%), instructions %d" % (ns, 100.0 * ns / tot_all, sum(l[4] for l in sel)))
print("stalls:", ", ".join("%s %d" % (k[6:], v) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for l in sorted(sel, key=lambda l: -l[3])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    top = sorted(l[5].items(), key=lambda x: -x[1])[:3]
    print("%s:%d  samples %d  inst %d  %s   %s" % (l[0], l[1], l[3], l[4], l[2][:80], " ".join("%s=%d" % (k[6:], v) for k, v in top if v)))
