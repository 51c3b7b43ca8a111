"""Host-buffer pass under MMLST_TRACE=1: where the time of mmlst_sample (one call) and of mmlst_score / mmlst_pileup_consensus (two seams) goes.
Prints one JSON line; the library's own trace lines go to stderr."""
import os, sys, time, json
os.environ["MMLST_TRACE"] = "1"
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
sidx = api.SampleIndex(ctx, index, st.ref_lens, db.row_seq)
P = bench.PARAMS
res = {}
for cover in (1.0, 0.9):
    soa.deflate(cover=cover)
    t = {"one_call": [], "score": [], "select": [], "pileup_consensus": []}
    for it in range(7):
        sys.stderr.write("-- cover %.2f it %d\n" % (cover, it))
        t0 = time.perf_counter()
        api.type_soa(sidx, soa, P["minscore"], P["max_xM"], P["min_read_len"], P["penalty"], 100)
        t1 = time.perf_counter()
        raw = api.score_soa_raw(ctx, soa, index, P["minscore"], P["max_xM"], P["min_read_len"])
        t2 = time.perf_counter()
        chosen = api.fast_select(index, raw[0], raw[1], raw[2], P["penalty"])
        ts = [x for _sp, tt in chosen for x in tt]
        t3 = time.perf_counter()
        api.pileup_consensus(ctx, soa, ts, [db.row_seq(x) for x in ts], P["minscore"], P["max_xM"], 1, 0)
        t4 = time.perf_counter()
        if it >= 2:
            t["one_call"].append(t1 - t0); t["score"].append(t2 - t1); t["select"].append(t3 - t2); t["pileup_consensus"].append(t4 - t3)
    res["cover_%.2f" % cover] = {k: round(float(np.median(v)) * 1e3, 3) for k, v in t.items()}
print(json.dumps(res))
