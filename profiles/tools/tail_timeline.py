#!/usr/bin/env python
"""Per-CTA timeline of the tail kernels of ONE serial pass (selection, depth-capped pileup, consensus) on configs[1], from the marks the kernels
store when a buffer is registered with mmlst_debug_timeline (global timer, ns).  Cold L2: the pass follows passes over two other samples.
Prints, per kernel and mark, min / median / max over the CTAs relative to the first mark of the pass, once for eager launches and once for a
CUDA-graph replay.  Usage (GPU box): python profiles/tools/tail_timeline.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from metamlst_b200 import api, native, pipeline  # noqa: E402

OUT = sys.argv[1] if len(sys.argv) > 1 else None
sys.argv = sys.argv[:1]

args = bench.parse()
dev = "cuda:0"
torch.cuda.set_device(0)
lib = native.lib()
db = bench.make_db(args)
index = api.AlleleIndex(db.ref_names())
sts = [bench.gen_streams(db, args, dev, 8000, None, seed=1002 + 100 * i)[0] for i in range(3)]
pipes = [pipeline.DevicePipeline(x, index, db.row_seq, **bench.PARAMS) for x in sts]
for p in pipes:
    for _ in range(3):
        want = p.step()
TLW = 8192
buf = torch.zeros(2 * TLW, dtype=torch.int64, device=dev)
NAMES = {"select": (0, ["entry", "rows arrived", "locus reduced", "ticket back", "finalized (last CTA)", "last CTA: results read back", "last CTA: loci ranked", "last CTA: header written"]),
         "pileup": (128, ["entry", "descriptor arrived", "flushed (warp 0)", "tile 0 landed", "tile 1 landed", "tile 2 landed", "counted (warp 0)", "flushed (all warps)"]),
         "consensus": (896, ["entry", "header arrived", "done"])}


def run(mode):
    buf.zero_()
    native.check(lib.mmlst_debug_timeline(buf.data_ptr()))
    p = pipes[0]
    if mode == "graph":
        p.capture()
    for q in pipes[1:]:
        q.step()
    torch.cuda.synchronize()
    if mode == "graph":
        p.step_graph()
    else:
        p.graph = None
        p.step()
    torch.cuda.synchronize()
    native.check(lib.mmlst_debug_timeline(0))
    p.graph = None
    t = buf.cpu().numpy().astype(np.int64)
    if OUT:
        np.save(OUT.replace(".json", "_%s_raw.npy" % mode), t)   # [2][1024 rows][8 marks]: global timer (ns), then the SM cycle counter
    g = t[:TLW].reshape(-1, 8)
    t0 = g[g > 0].min()
    res = {}
    for k, (row0, marks) in NAMES.items():
        row1 = min([r for r, _ in NAMES.values() if r > row0] + [TLW // 8])
        rows = g[row0:row1]
        live = rows[rows[:, 0] > 0]
        res[k] = {"ctas": int(live.shape[0]), "marks": {}}
        for m, name in enumerate(marks):
            v = live[:, m]
            v = v[v > 0] - t0
            if v.size:
                res[k]["marks"][name] = {"n": int(v.size), "min_us": float(v.min()) / 1e3, "median_us": float(np.median(v)) / 1e3, "max_us": float(v.max()) / 1e3}
    return res


out = {}
for mode in ("eager", "graph"):
    out[mode] = run(mode)
    print("====", mode)
    for k, r in out[mode].items():
        print("  %s: %d CTAs" % (k, r["ctas"]))
        for name, s in r["marks"].items():
            print("    %-24s n=%-4d min %7.2f  median %7.2f  max %7.2f us" % (name, s["n"], s["min_us"], s["median_us"], s["max_us"]))
if OUT:
    json.dump(out, open(OUT, "w"), indent=1)
