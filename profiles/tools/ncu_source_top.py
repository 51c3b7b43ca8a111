#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (one block per kernel): position in the stream,
samples, dominant stall reasons, times executed.  Usage: ncu_source_top.py file.csv [min_pct]"""
import csv, sys
path = sys.argv[1]; min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr) - 2: body.append(rows[j])
            j += 1
        col = {h: k for k, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[col["# Samples"]] or 0) for r in body)
        print("==== %s: %d instructions, %d samples" % (name[:80], len(body), tot))
        for n, r in enumerate(body):
            s = int(r[col["# Samples"]] or 0)
            if tot and 100.0 * s / tot >= min_pct:
                st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
                print("%5d %5.1f%% x%-7s %-60s %s" % (n, 100.0 * s / tot, r[col["Instructions Executed"]], r[col["Source"]].strip()[:60], " ".join("%s:%d" % (h, v) for v, h in st if v)))
        i = j
    else:
        i += 1
