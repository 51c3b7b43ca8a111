"""Shared test helpers (TEST side only)."""
import numpy as np

from metamlst_b200 import synth
from oracle import bamio


def small_case(seed=5, n_reads=600, L=100, K=4, orgs=("ecoli",), apl=6, schemes=None, **kw):
    db = synth.make_db(orgs, alleles_per_locus=apl, n_profiles=10, seed=seed, schemes=schemes)
    tab = synth.make_sample(db, n_reads, L, seed=seed, K=K, **kw)
    return db, tab


def table_to_records(tab):
    h = bamio.BamHeader("", list(tab.ref_names), [int(x) for x in tab.ref_lens])
    return h, list(bamio.table_records(tab))


def lut_from_db(db, species_filter=None):
    """allow[tid], locus_of[tid] from a SynthDB (what the host derives from the BAM header names)."""
    allow = np.ones(db.n_rows, np.uint8)
    if species_filter:
        keep = set(species_filter.split(","))
        allow = np.array([1 if db.locus_names[int(l)][0] in keep else 0 for l in db.row_locus], np.uint8)
    return allow, db.row_locus.astype(np.uint32), len(db.locus_names)


def planes_to_counts(soa, tid, contig_len, minscore, max_xm):
    """Numpy decode of the packed pileup stream of one contig (test-side check of the PACKER, not a product path)."""
    counts = np.zeros((contig_len, 5), np.int64)
    r0, r1 = int(soa.contig_start[tid]), int(soa.contig_start[tid + 1])
    for r in range(r0, r1):
        off = int(soa.p_row_off[r]); rl = int(soa.p_reflen[r]); p = int(soa.p_pos[r])
        ok = int(soa.p_as[r]) >= minscore and int(soa.p_xm[r]) <= max_xm
        nw = int(soa.p_recs["nw"][r])
        assert nw == (((p & 31) + rl + 31) // 32 if rl else 0)
        row = soa.planes[off:off + 3 * nw].reshape(nw, 3)
        for j in range(nw):
            v, b1, b0 = int(row[j, 0]), int(row[j, 1]), int(row[j, 2])
            for i in range(32):
                vb, h, l = (v >> i) & 1, (b1 >> i) & 1, (b0 >> i) & 1
                if not (vb | l):
                    continue
                col = (p >> 5) * 32 + 32 * j + i  # rows are aligned to the contig's 32-column words
                if 0 <= col < contig_len:
                    counts[col, (2 * h + l) if (vb and ok) else 4] += 1
    return counts


def records_to_table(h, recs):
    """oracle.bamio records -> AlnTable (test side: lets GPU tests consume the committed golden BAMs)."""
    from metamlst_b200.synth import AlnTable
    n = len(recs)
    L = max((len(r.seq) for r in recs), default=1)
    seq = np.full((n, L), ord("N"), np.uint8)
    qual = np.zeros((n, L), np.uint8)
    cig_off = np.zeros(n + 1, np.int64)
    ops = []
    def tag(r, t, default=0):
        for k, _ty, v in r.aux:
            if k == t:
                return int(v)
        return default
    AS = np.zeros(n, np.int32); XS = np.zeros(n, np.int32); has_xs = np.zeros(n, bool)
    XN = np.zeros(n, np.int32); XM = np.zeros(n, np.int32); XO = np.zeros(n, np.int32); XG = np.zeros(n, np.int32); NM = np.zeros(n, np.int32)
    for i, r in enumerate(recs):
        assert len(r.seq) == L, "helper handles fixed-length reads"
        seq[i] = np.frombuffer(r.seq.encode(), np.uint8)
        qual[i] = np.frombuffer(bytes(r.qual), np.uint8)
        ops.extend((l << 4) | op for op, l in r.cigar)
        cig_off[i + 1] = len(ops)
        AS[i] = tag(r, "AS"); XM[i] = tag(r, "XM"); XO[i] = tag(r, "XO"); XG[i] = tag(r, "XG"); NM[i] = tag(r, "NM"); XN[i] = tag(r, "XN")
        has_xs[i] = any(k == "XS" for k, _t, _v in r.aux)
        XS[i] = tag(r, "XS")
        assert [k for k, _t, _v in r.aux][0] == "AS"
    return AlnTable(list(h.ref_names), np.asarray(h.ref_lens, np.int32), np.array([r.tid for r in recs], np.int32),
                    np.array([r.pos for r in recs], np.int32), np.array([r.flag for r in recs], np.uint16),
                    np.array([int(r.qname[1:]) for r in recs], np.int64), cig_off, np.asarray(ops, np.uint32), L, seq, qual,
                    AS, XS, has_xs, XN, XM, XO, XG, NM)
