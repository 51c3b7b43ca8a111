#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[ui] == "us":
        v *= 1e3
    elif r[ui] == "ms":
        v *= 1e6
    agg[r[ki][:90]][0] += 1
    agg[r[ki][:90]][1] += v
tot = sum(v for _n, v in agg.values())
print("total %.1f us over %d launches (cold-cache, serialised: compare SHARES)" % (tot / 1e3, sum(n for n, _v in agg.values())))
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%6d  %12.1f us  %5.1f%%  %s" % (n, v / 1e3, 100 * v / tot, k))
