"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/c/mlst_oracle.c working on AlnTable-style numpy arrays."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "liboracle.so")
_lib = None


def build():
    src = os.path.join(_DIR, "mlst_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _DIR, "liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_depth_cap_sim.restype = C.c_int
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def positional_aux(tab):
    """(aux0, aux3) by POSITION (metamlst.py:109-110): AS, and XM when XS:i is present else XO (H4)."""
    return tab.AS.astype(np.int32), np.where(tab.has_xs, tab.XM, tab.XO).astype(np.int32)


def score(tab, allow, locus_of, n_loci, minscore=80, max_xm=5, min_read_len=50, orig_idx=None):
    n_ref = len(tab.ref_names)
    aux0, aux3 = positional_aux(tab)
    qlen = np.full(tab.n, max(tab.read_len, 1), dtype=np.int32)
    sum_as = np.zeros(n_ref, np.int64)
    n_hit = np.zeros(n_ref, np.uint32)
    first = np.full(n_ref, 0xFFFFFFFF, np.uint32)
    counters = np.zeros(2, np.uint64)
    tid = np.ascontiguousarray(tab.tid, dtype=np.int32)
    oi = None if orig_idx is None else np.ascontiguousarray(orig_idx, dtype=np.uint32)
    allow_a = np.ascontiguousarray(allow, np.uint8)  # keep temporaries alive across the call
    locus_a = np.ascontiguousarray(locus_of, np.uint32)
    lib().orc_score(C.c_uint64(tab.n), _p(tid), _p(aux0), _p(aux3), _p(qlen), _p(oi), _p(allow_a),
                    _p(locus_a), minscore, max_xm, min_read_len, _p(sum_as), _p(n_hit), _p(first), _p(counters), C.c_uint64(0))
    return sum_as, n_hit, first, counters


def ref_lengths(tab):
    op = tab.cig_ops & 0xF
    ln = (tab.cig_ops >> 4).astype(np.int64)
    rec = np.repeat(np.arange(tab.n), (tab.cig_off[1:] - tab.cig_off[:-1]))
    return np.bincount(rec, weights=ln * np.isin(op, (0, 2, 3, 7, 8)), minlength=tab.n).astype(np.int32)


def contig_counts(tab_sorted, tid, minqual=20, minscore=80, max_xm=5, max_depth=8000, sentinel_nodes=1):
    """counts[len][5] (A,C,G,T,N) for one contig of a COORDINATE-SORTED table, with the htslib depth cap simulated."""
    idx = np.nonzero((tab_sorted.tid == tid) & ((tab_sorted.flag & 4) == 0))[0]
    sub = tab_sorted.take(idx)
    n = sub.n
    clen = int(tab_sorted.ref_lens[tid])
    counts = np.zeros((clen, 5), np.uint32)
    if n == 0:
        return counts, np.zeros(0, np.uint8)
    pos = np.ascontiguousarray(sub.pos, np.int32)
    rl = ref_lengths(sub)
    admitted = np.ones(n, np.uint8)
    if max_depth is not None:
        rc = lib().orc_depth_cap_sim(C.c_uint64(n), C.c_int32(int(tid)), _p(pos), _p(rl), C.c_uint32(max_depth), C.c_uint32(sentinel_nodes), _p(admitted))
        if rc != 0:
            raise ValueError("The input is not sorted (reads out of order)")
    seq = np.ascontiguousarray(sub.seq)
    qual = np.ascontiguousarray(sub.qual)
    cig_off = np.ascontiguousarray(sub.cig_off, np.int64)
    cig_ops = np.ascontiguousarray(sub.cig_ops, np.uint32)
    as_n, xm_n = sub.AS.astype(np.int32), sub.XM.astype(np.int32)
    lib().orc_pileup(C.c_uint64(n), _p(pos), _p(cig_off), _p(cig_ops),
                     _p(seq), _p(qual), C.c_int(sub.read_len), _p(as_n), _p(xm_n), _p(admitted),
                     minqual, minscore, max_xm, C.c_int32(clen), _p(counts))
    return counts, admitted


def consensus(counts, dbseq: bytes, mincov=1):
    ln = counts.shape[0]
    cons = np.zeros(ln, np.uint8)
    h = C.c_uint32()
    s = C.c_uint32()
    db = np.frombuffer(dbseq, dtype=np.uint8)
    assert db.shape[0] >= ln
    cnt = np.ascontiguousarray(counts, np.uint32)
    lib().orc_consensus(_p(cnt), _p(db), C.c_int32(ln), C.c_uint32(mincov), _p(cons), C.byref(h), C.byref(s))
    return cons.tobytes().decode(), h.value, s.value


def hamming_min(queries, rows_flat, row_off, ranges):
    """queries: list[bytes]; rows_flat uint8, row_off int64 [A+1]; ranges: list[(r0, r1)] per query."""
    out_d = np.zeros(len(queries), np.uint32)
    out_a = np.zeros(len(queries), np.uint32)
    d = C.c_uint32()
    a = C.c_uint32()
    ro = np.ascontiguousarray(row_off, np.int64)
    for i, q in enumerate(queries):
        qa = np.frombuffer(q, dtype=np.uint8)
        lib().orc_hamming_min(_p(qa), C.c_int32(len(q)), _p(rows_flat), _p(ro), C.c_uint32(ranges[i][0]), C.c_uint32(ranges[i][1]), C.byref(d), C.byref(a))
        out_d[i], out_a[i] = d.value, a.value
    return out_d, out_a
