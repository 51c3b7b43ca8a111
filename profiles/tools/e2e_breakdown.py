"""Where the host-buffer pass (bench.py `e2e`) spends its time: wall clock of each of its three calls, plain and compressed streams."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from metamlst_b200 import api, native

class A: pass
args = A(); args.reads = 10_000_000; args.read_len = 150; args.k = 4; args.alleles = 1024; args.max_depth = 8000
db = bench.make_db(args)
st, _ = bench.gen_streams(db, args, "cuda:0", 8000)
index = api.AlleleIndex(db.ref_names())
soa = st.to_host(pinned=True)
ctx = native.Context(0)
P = bench.PARAMS
res = {}
for form in ("plain", "deflated_1M", "deflated_256K", "deflated_64K", "deflated_16K"):
    if form != "plain":
        soa.deflate(block={"1M": 1 << 20, "256K": 1 << 18, "64K": 1 << 16, "16K": 1 << 14}[form.split("_")[1]])
    t = {"score": [], "select": [], "pileup_consensus": []}
    for it in range(8):
        t0 = time.perf_counter()
        raw = api.score_soa_raw(ctx, soa, index, P["minscore"], P["max_xM"], P["min_read_len"])
        t1 = time.perf_counter()
        chosen = api.fast_select(index, raw[0], raw[1], raw[2], P["penalty"])
        ts = [x for _sp, tt in chosen for x in tt]
        t2 = time.perf_counter()
        api.pileup_consensus(ctx, soa, ts, [db.row_seq(x) for x in ts], P["minscore"], P["max_xM"], 1, 0)
        t3 = time.perf_counter()
        if it >= 3:
            t["score"].append(t1 - t0); t["select"].append(t2 - t1); t["pileup_consensus"].append(t3 - t2)
    res[form] = {k: round(float(np.median(v)) * 1e3, 3) for k, v in t.items()}
print(json.dumps(res))
