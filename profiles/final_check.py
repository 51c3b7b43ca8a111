#!/usr/bin/env python3
"""Ten-second GPU sanity check of the library as finally built (no torch import): the default run-length score kernel (form 5,
TMA ring) through the host entry on a 100 003-record stream against numpy, and a golden BAM through the native unpacker
and both pileup implementations against the Python oracle."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metamlst_b200 import api, bam, native, packing
from oracle import bamio, mlst_oracle as orc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ctx = native.Context(0)
lib = native.lib()
assert lib.mmlst_set_score_variant(-1) == 5
rng = np.random.default_rng(3)
n, n_ref = 100003, 300
lens = rng.integers(1, 900, 4000); lens = lens[np.cumsum(lens) <= n]; lens = np.concatenate([lens, [n - lens.sum()]])
rt = rng.integers(0, n_ref, lens.shape[0]); rt[1:][rt[1:] == rt[:-1]] += 1; rt %= n_ref
tid = np.repeat(rt, lens).astype(np.uint32)
as0 = rng.integers(-50, 301, n).astype(np.int16); xm3 = rng.integers(0, 8, n).astype(np.uint8)
qlen = np.repeat(rng.integers(30, 160, (n + 255) // 256), 256)[:n].astype(np.uint16)
allow = (rng.random(n_ref) < 0.8).astype(np.uint8)
names = ["o%d_g%d_%d" % (t % 3, t % 7, t) for t in range(n_ref)]
index = api.AlleleIndex(names)
soa = packing.SoaHost(names, np.full(n_ref, 500, np.int32), tid, as0, xm3, qlen, None, np.zeros(0, packing.PREC_DTYPE), np.zeros(0, np.uint32), 0,
                      np.zeros(n_ref + 1, np.uint64)).build_runs(max_fraction=1.0)
assert soa.chunk_qlen is not None
s = np.zeros(n_ref, np.int64); c = np.zeros(n_ref, np.uint32); f = np.full(n_ref, 0xFFFFFFFF, np.uint32); k = np.zeros(2, np.uint64)
cs = soa.c_struct(); prm = native.ScoreParams(100, 5, 50)
native.check(lib.mmlst_score(ctx.handle, C.byref(cs), native.ptr(allow), native.ptr(index.locus_of), index.n_loci, C.byref(prm), native.ptr(s), native.ptr(c), native.ptr(f), native.ptr(k)))
al = allow[tid] != 0
ok = al & (as0 >= 100) & (qlen >= 50) & (xm3 <= 5)
ws = np.zeros(n_ref, np.int64); wc = np.zeros(n_ref, np.int64); wf = np.full(n_ref, 0xFFFFFFFF, np.int64)
np.add.at(ws, tid[ok], as0[ok].astype(np.int64)); np.add.at(wc, tid[ok], 1); np.minimum.at(wf, tid[ok], np.arange(n)[ok])
assert np.array_equal(s, ws) and np.array_equal(c, wc.astype(np.uint32)) and np.array_equal(f, wf.astype(np.uint32)) and int(k[0]) == int(al.sum())
print("score (form 5 ring) ok")
path = os.path.join(ROOT, "tests", "golden", "basic", "sample.bam")
soa2 = bam.unpack_bam(path)
h, recs = bamio.read_bam(path)
srt = sorted(recs, key=lambda r: (r.tid, r.pos, (r.flag >> 4) & 1))
per = np.bincount([r.tid for r in recs], minlength=len(h.ref_names))
tids = [int(t) for t in np.argsort(-per)[:3]]
tf = [("AS", "loc_gte", 80), ("XM", "loc_lte", 5)]
wants = []
for t in tids:
    full, _ = orc.get_base_stats([r for r in srt if r.tid == t], t, 0, 20, tf, 8000)
    w = np.zeros((h.ref_lens[t], 5), np.int64)
    for pos1, d in full.items():
        fq = d["base_freq"]
        w[pos1 - 1] = [fq["A"], fq["C"], fq["G"], fq["T"], fq["N"]]
    wants.append(w)
for impl in (1, 2):
    seqs, holes, snps, counts, col_off = api.pileup_consensus(ctx, soa2, tids, ["A" * h.ref_lens[t] for t in tids], 80, 5, 1, impl, True)
    for i, t in enumerate(tids):
        assert np.array_equal(counts[col_off[i]:col_off[i + 1]], wants[i]), (impl, t)
print("unpack (own inflate) + pileup (atomic, bit-sliced) ok")
ctx.close()
