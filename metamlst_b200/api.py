"""Host-side mirror of the reference's four Python seams on top of libmmlst.

    S1  score_soa          <-> metamlst.py:96-151          (cel / totalReads / ignoredReads)
    S2  build_consensus    <-> metaMLST_functions.py:249-281 (same name, argument meaning and return shape)
    S3  HammingIndex       <-> metamlst-merge.py:174-181 + metaMLST_functions.py:224-234
    S4  define_profile     <-> metaMLST_functions.py:205-216 (SQL kept: ms-scale, H11 quirks preserved verbatim)

All per-record / per-base / per-pair arithmetic runs in the CUDA kernels; what stays here is what the reference
also does once per allele or per locus in Python floats and strings (penalty, round(.,1), dict ordering, formatting),
so that those are bit-identical by construction (H5, H6).
"""
from __future__ import annotations

import ctypes as C
import queue
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import native, packing

NO_IDX = 0xFFFFFFFF
XM_SATURATION = 255  # the packed XM fields are 8 bits and saturate here: a threshold at or above it could let a larger XM through


def check_max_xm(max_xM: int) -> int:
    """--max_xM against the 8-bit XM fields of the streams: up to 254 is exact; anything above is refused, never silently different
    (the lower-level pileup_consensus takes 255 as "no XM filter at all", which is exact too: cmseq with BAM_tagFilter=None)."""
    if int(max_xM) >= XM_SATURATION:
        raise ValueError("max_xM must be at most %d (XM is carried as 8 bits, saturated at %d)" % (XM_SATURATION - 1, XM_SATURATION))
    return int(max_xM)


class AlleleIndex:
    """BAM reference names `organism_gene_allele` (metamlst.py:107) -> species / locus / allele tables."""

    def __init__(self, ref_names: Sequence[str]):
        self.ref_names = list(ref_names)
        self.species: List[str] = []
        self.gene: List[str] = []
        self.allele: List[str] = []
        self.locus_names: List[Tuple[str, str]] = []
        loc: Dict[Tuple[str, str], int] = {}
        locus_of = np.zeros(len(self.ref_names), dtype=np.uint32)
        for i, name in enumerate(self.ref_names):
            parts = name.split("_")
            if len(parts) != 3:
                # the reference dies with "ValueError: not enough/too many values to unpack" at metamlst.py:107 as soon
                # as a record on this reference is read; we refuse the header up front
                raise ValueError("reference name %r is not organism_gene_allele (metamlst.py:107)" % name)
            s, g, a = parts
            self.species.append(s)
            self.gene.append(g)
            self.allele.append(a)
            key = (s, g)
            if key not in loc:
                loc[key] = len(self.locus_names)
                self.locus_names.append(key)
            locus_of[i] = loc[key]
        self.locus_of = locus_of
        self.n_loci = len(self.locus_names)
        self.name_to_tid = {n: i for i, n in enumerate(self.ref_names)}

    def allow_mask(self, species_filter: Optional[str]) -> np.ndarray:
        """metamlst.py:114: `(args.filter and species in args.filter.split(',')) or not args.filter`."""
        if not species_filter:
            return np.ones(len(self.ref_names), dtype=np.uint8)
        keep = set(species_filter.split(","))
        return np.fromiter((1 if s in keep else 0 for s in self.species), dtype=np.uint8, count=len(self.species))


def finish_scores(index: AlleleIndex, sum_as: np.ndarray, n_hit: np.ndarray, first_idx: np.ndarray, penalty: int):
    """Integer tables -> `cel` exactly as metamlst.py holds it after line 151.
    Dict insertion order (H5) = ascending first passing record index of species, locus, allele."""
    hit = np.nonzero(n_hit)[0]
    hit = hit[np.argsort(first_idx[hit], kind="stable")]
    cel: Dict[str, Dict[str, Dict[str, tuple]]] = {}
    raw: Dict[Tuple[str, str], List[Tuple[str, int, int]]] = {}
    for t in hit:
        t = int(t)
        s, g, a = index.species[t], index.gene[t], index.allele[t]
        cel.setdefault(s, {}).setdefault(g, {})
        raw.setdefault((s, g), []).append((a, int(sum_as[t]), int(n_hit[t])))
    for (s, g), lst in raw.items():
        maxLen = max(n for (_a, _s, n) in lst)  # metamlst.py:137
        for a, localScore, geneLen in lst:
            if geneLen != maxLen:
                localScore = localScore - (maxLen - geneLen) * penalty  # :146-147
            averageScore = float(localScore) / float(geneLen)
            cel[s][g][a] = (localScore, geneLen, round(averageScore, 1))  # :151 (H6: Python float rounding)
    return cel


def refuse_lenient(soa) -> None:
    """A stream unpacked with lenient_tags carries made-up positional fields for the records metamlst.py:109-110 would have crashed on: it is for
    cmseq-style pileups only, the score seams refuse it."""
    if getattr(soa, "lenient", False):
        raise ValueError("this stream was unpacked with lenient_tags (cmseq use): metamlst.py's read loop would have crashed on such a BAM, it cannot be scored")


def score_soa_raw(ctx: native.Context, soa: packing.SoaHost, index: AlleleIndex, minscore: int = 80, max_xM: int = 5,
                  min_read_len: int = 50, species_filter: Optional[str] = None):
    """Seam S1, integer half: (sum_as, n_hit, first_idx, totalReads, ignoredReads) straight from mmlst_score."""
    n_ref = len(index.ref_names)
    check_max_xm(max_xM)
    refuse_lenient(soa)
    allow = index.allow_mask(species_filter)
    sum_as = np.zeros(n_ref, np.int64)
    n_hit = np.zeros(n_ref, np.uint32)
    first_idx = np.full(n_ref, NO_IDX, np.uint32)
    counters = np.zeros(2, np.uint64)
    cs = soa.c_struct()
    prm = native.ScoreParams(int(minscore), int(max_xM), int(min_read_len))
    native.check(native.lib().mmlst_score(ctx.handle, C.byref(cs), native.ptr(allow), native.ptr(index.locus_of), index.n_loci,
                                          C.byref(prm), native.ptr(sum_as), native.ptr(n_hit), native.ptr(first_idx),
                                          native.ptr(counters)))
    return sum_as, n_hit, first_idx, int(counters[0]), int(counters[1])


def score_soa(ctx: native.Context, soa: packing.SoaHost, index: AlleleIndex, minscore: int = 80, max_xM: int = 5,
              min_read_len: int = 50, species_filter: Optional[str] = None, penalty: int = 100):
    """Seam S1.  Returns (cel, totalReads, ignoredReads, raw) with raw = (sum_as, n_hit, first_idx) numpy tables."""
    sum_as, n_hit, first_idx, total, ignored = score_soa_raw(ctx, soa, index, minscore, max_xM, min_read_len, species_filter)
    cel = finish_scores(index, sum_as, n_hit, first_idx, penalty)
    return cel, total, ignored, (sum_as, n_hit, first_idx)


def coverage_sums(ctx: native.Context, soa: packing.SoaHost, index: AlleleIndex, minscore: int = 80, max_xM: int = 5,
                  min_read_len: int = 50, species_filter: Optional[str] = None, stream_resident: bool = False) -> Dict[str, int]:
    """Seam S1, coverage column (H7): {'species_gene': sum over unique QNAMEs of len(SEQ) of the last passing record} ==
    `sum(sequenceBank[key].values())` at metamlst.py:228; keys exist for loci with at least one passing record.
    `stream_resident`: the score stream is still on the device from the score_soa call just made on this context."""
    if soa.qhash is None:
        raise ValueError("stream was unpacked without QNAME hashes (unpack_bam(want_qhash=True))")
    allow = index.allow_mask(species_filter)
    cov = np.zeros(max(index.n_loci, 1), np.uint64)
    cs = soa.c_struct()
    prm = native.ScoreParams(int(minscore), int(max_xM), int(min_read_len))
    qh = np.ascontiguousarray(soa.qhash, dtype=np.uint64)
    native.check(native.lib().mmlst_coverage(ctx.handle, C.byref(cs), native.ptr(qh), native.ptr(allow), native.ptr(index.locus_of),
                                             index.n_loci, C.byref(prm), native.COVERAGE_STREAM_RESIDENT if stream_resident else 0,
                                             native.ptr(cov)))
    return {s + "_" + g: int(cov[l]) for l, (s, g) in enumerate(index.locus_names) if cov[l]}


def select_alleles(species_cel: Dict[str, Dict[str, tuple]]) -> List[Tuple[str, str]]:
    """metamlst.py:244: per locus (dict order) the lowest-numbered allele among those whose rounded average equals
    the locus maximum."""
    out = []
    for g1, g2 in species_cel.items():
        best = max(avg for (_v, _l, avg) in g2.values())
        out.append((g1, sorted((k for k, (_v, _l, avg) in g2.items() if avg == best), key=lambda x: int(x))[0]))
    return out


class ConsRecord:
    """What buildConsensus' callers touch on a Bio.SeqRecord (metamlst.py:254-285): id, seq (mutable), description."""

    def __init__(self, seq: str, id: str, description: str):
        self.seq = seq
        self.id = id
        self.description = description

    def __repr__(self):
        return "ConsRecord(id=%r, %s, len=%d)" % (self.id, self.description, len(self.seq))


def pileup_consensus(ctx: native.Context, soa: packing.SoaHost, chosen_tid: Sequence[int], dbseqs: Sequence[str],
                     minscore: int, max_xM: int, mincov: int = 1, impl: int = 0, want_counts: bool = False):
    """Counts + consensus for the chosen contigs.  Returns (cons strings, holes, snps, counts or None, col_off)."""
    n = len(chosen_tid)
    lens = [int(soa.ref_lens[t]) for t in chosen_tid]
    col_off = np.zeros(n + 1, dtype=np.uint32)
    col_off[1:] = np.cumsum(lens)
    total = int(col_off[-1])
    db = np.zeros(max(total, 1), dtype=np.uint8)
    for i, (t, s) in enumerate(zip(chosen_tid, dbseqs)):
        if len(s) < lens[i]:
            # metaMLST_functions.py:267/269 index dbSequen[i] for i < BAM LN (H10)
            raise IndexError("string index out of range: BAM LN %d > DB sequence length %d for %s" % (lens[i], len(s), soa.ref_names[t]))
        db[col_off[i]:col_off[i + 1]] = np.frombuffer(s[:lens[i]].encode("latin-1"), dtype=np.uint8)
    tid_arr = np.asarray(chosen_tid, dtype=np.uint32)
    counts = np.zeros((max(total, 1), 5), dtype=np.uint32) if want_counts else None
    cons = np.zeros(max(total, 1), dtype=np.uint8)
    holes = np.zeros(max(n, 1), dtype=np.uint32)
    snps = np.zeros(max(n, 1), dtype=np.uint32)
    cs = soa.c_struct()
    native.check(native.lib().mmlst_pileup_consensus(ctx.handle, C.byref(cs), native.ptr(tid_arr), n, native.ptr(db),
                                                     native.ptr(col_off), int(minscore), int(max_xM), int(mincov), int(impl),
                                                     native.ptr(counts), native.ptr(cons), native.ptr(holes), native.ptr(snps)))
    seqs = [cons[col_off[i]:col_off[i + 1]].tobytes().decode("latin-1") for i in range(n)]
    return seqs, holes[:n].copy(), snps[:n].copy(), (counts[:total] if want_counts else None), col_off


class SampleIndex:
    """What `type_soa` keeps resident in a context: the allele index, the DB sequences and the BAM header lengths (include/mmlst.h, mmlst_index).
    genes_in_db: {organism: rows of `genes`} (metamlst.py:184); default = loci of the organism among the references."""

    def __init__(self, ctx: native.Context, index: AlleleIndex, ref_lens: Sequence[int], dbseq_of, genes_in_db: Optional[Dict[str, int]] = None):
        self.ctx, self.index = ctx, index
        n_ref = len(index.ref_names)
        self.species_names: List[str] = []
        sp_id: Dict[str, int] = {}
        sol = np.zeros(index.n_loci, dtype=np.uint32)
        for l, (sp, _g) in enumerate(index.locus_names):
            if sp not in sp_id:
                sp_id[sp] = len(self.species_names)
                self.species_names.append(sp)
            sol[l] = sp_id[sp]
        gdb = np.asarray([(genes_in_db or {}).get(sp, int((sol == i).sum())) for i, sp in enumerate(self.species_names)], dtype=np.uint32)
        seqs = [dbseq_of(t).encode("latin-1") for t in range(n_ref)]
        db_off = np.zeros(n_ref + 1, dtype=np.uint64)
        db_off[1:] = np.cumsum([len(x) for x in seqs])
        db_ascii = np.frombuffer(b"".join(seqs) + b"\0" * 8, dtype=np.uint8)
        allele_num = np.asarray([int(a) for a in index.allele], dtype=np.int64).astype(np.uint32)
        self.ref_lens = np.ascontiguousarray(ref_lens, dtype=np.uint32)
        self.n_loci = index.n_loci
        self.max_cols = int(np.sort(self.ref_lens)[::-1][: self.n_loci].sum())
        ix = native.Index(native.ptr(index.locus_of), native.ptr(allele_num), n_ref, native.ptr(sol), index.n_loci, native.ptr(gdb), len(self.species_names),
                          native.ptr(db_ascii), native.ptr(db_off), native.ptr(self.ref_lens))
        native.check(native.lib().mmlst_index_upload(ctx.handle, C.byref(ix)))
        nl = self.n_loci
        self._tid = np.zeros(nl, np.uint32); self._sp = np.zeros(nl, np.uint32); self._col = np.zeros(nl + 1, np.uint32)
        self._holes = np.zeros(nl, np.uint32); self._snps = np.zeros(nl, np.uint32); self._cons = np.zeros(self.max_cols + 16, np.uint8)
        self._allow: Dict[Optional[str], np.ndarray] = {}


def type_soa(sidx: SampleIndex, soa: packing.SoaHost, minscore: int = 80, max_xM: int = 5, min_read_len: int = 50, penalty: int = 100, nloci: int = 100,
             species_filter: Optional[str] = None, mincov: int = 1, impl: int = 0, want_tables: bool = False):
    """One sample held in host buffers, ONE library call (mmlst_sample): score (seam S1, metamlst.py:96-151), selection in the reference's dict order
    (metamlst.py:133-220, H5/H6, --nloci gate), pileup + consensus of the chosen contigs (seam S2, metaMLST_functions.py:249-281).
    Returns {"species": [(organism, [(ref name, consensus, holes, snps)])], "totalReads", "ignoredReads", "tids", "tables"}."""
    check_max_xm(max_xM)
    refuse_lenient(soa)
    if soa.minqual != 20:
        raise ValueError("buildConsensus needs a stream unpacked with minqual=20 (metaMLST_functions.py:258)")
    allow = sidx._allow.get(species_filter)
    if allow is None:
        allow = sidx._allow[species_filter] = sidx.index.allow_mask(species_filter)
    res = native.SampleResult()
    res.chosen_tid, res.chosen_species, res.col_off = native.ptr(sidx._tid), native.ptr(sidx._sp), native.ptr(sidx._col)
    res.cons, res.cons_capacity, res.holes, res.snps = native.ptr(sidx._cons), sidx._cons.shape[0], native.ptr(sidx._holes), native.ptr(sidx._snps)
    tables = None
    if want_tables:
        n_ref = len(sidx.index.ref_names)
        tables = (np.zeros(n_ref, np.int64), np.zeros(n_ref, np.uint32), np.zeros(n_ref, np.uint32))
        res.sum_as, res.n_hit, res.first_idx = (native.ptr(t) for t in tables)
    prm = native.SampleParams(int(minscore), int(max_xM), int(min_read_len), int(penalty), int(nloci), int(mincov), int(impl))
    cs = soa.c_struct()
    rc = native.lib().mmlst_sample(sidx.ctx.handle, C.byref(cs), native.ptr(allow), C.byref(prm), C.byref(res))
    if rc == native.E_RANGE and native.lib().mmlst_last_error().startswith(b"string index out of range"):
        raise IndexError(native.lib().mmlst_last_error().decode())   # H10
    native.check(rc)
    if res.error_bits & 1:
        raise RuntimeError("Database is broken")   # metamlst.py:188-190
    n = int(res.n_chosen)
    cons = sidx._cons.tobytes()
    names = sidx.index.ref_names
    species: List[Tuple[str, list]] = []
    for i in range(n):
        sp = sidx.species_names[int(sidx._sp[i])]
        if not species or species[-1][0] != sp:
            species.append((sp, []))
        species[-1][1].append((names[int(sidx._tid[i])], cons[int(sidx._col[i]):int(sidx._col[i + 1])].decode("latin-1"), int(sidx._holes[i]), int(sidx._snps[i])))
    return {"species": species, "totalReads": int(res.total_reads), "ignoredReads": int(res.ignored_reads), "tids": [int(t) for t in sidx._tid[:n]],
            "tables": tables}


class SampleLanes:
    """Several samples in flight through the host-buffer path (the cohort form of `type_soa`; `pipeline.CohortLanes` is the same idea for streams that
    are already resident): `lanes` contexts on one device -- each with its own stream, device buffers, page-locked staging block and resident index
    (mmlst_index_upload) -- and one host thread per lane inside the synchronous `mmlst_sample` call (ctypes drops the interpreter lock for its duration).
    While one sample waits for its selection block or runs its pileup, the next one's score stream is already on the bus, so the PCIe link and the
    decompression engine stay busy across the two synchronisation points of a call.  Results are those of `type_soa`, in submission order."""

    def __init__(self, device: int, index: AlleleIndex, ref_lens: Sequence[int], dbseq_of, lanes: int = 2, genes_in_db: Optional[Dict[str, int]] = None):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.ctxs = [native.Context(device) for _ in range(lanes)]
        self.sidx = [SampleIndex(c, index, ref_lens, dbseq_of, genes_in_db) for c in self.ctxs]
        self._free: "queue.SimpleQueue[SampleIndex]" = queue.SimpleQueue()
        for s in self.sidx:
            self._free.put(s)
        self._pool = ThreadPoolExecutor(max_workers=lanes, thread_name_prefix="mmlst-lane")

    def _one(self, soa: packing.SoaHost, kw: dict):
        s = self._free.get()
        try:
            return type_soa(s, soa, **kw)
        finally:
            self._free.put(s)

    def submit(self, soa: packing.SoaHost, **kw) -> Future:
        """One sample; the Future's result is `type_soa`'s.  Keyword arguments as `type_soa`."""
        return self._pool.submit(self._one, soa, kw)

    def map(self, soas: Iterable[packing.SoaHost], **kw) -> list:
        futs = [self.submit(soa, **kw) for soa in soas]
        return [f.result() for f in futs]

    def close(self):
        self._pool.shutdown(wait=True)
        for c in self.ctxs:
            c.close()
        self.ctxs, self.sidx = [], []


def build_consensus(ctx: native.Context, soa: packing.SoaHost, chromosomeList: Dict[str, str], filterScore: int, max_xM: int,
                    debugMode: bool = False, impl: int = 0) -> List[ConsRecord]:
    """Seam S2 -- same contract as metaMLST_functions.buildConsensus(bamFile, chromosomeList, filterScore, max_xM,
    debugMode): list order = chromosomeList order, rec.description == 'CI::<holes>_SP::<snps>'.  The BAM is replaced by
    its unpacked SoaHost (consensus rule hard-wired upstream: dominant 0.4 (output-dead), mincov 1, minqual 20)."""
    if soa.minqual != 20:
        raise ValueError("buildConsensus needs a stream unpacked with minqual=20 (metaMLST_functions.py:258)")
    check_max_xm(max_xM)
    name2tid = {n: i for i, n in enumerate(soa.ref_names)}
    contigs = list(chromosomeList.keys())
    for c in contigs:
        if c not in name2tid:
            # cmseq get_contig_by_label returns None -> AttributeError upstream (cmseq/cmseq.py:88)
            raise AttributeError("'NoneType' object has no attribute 'reference_free_consensus' (contig %s not in BAM)" % c)
    tids = [name2tid[c] for c in contigs]
    seqs, holes, snps, _, _ = pileup_consensus(ctx, soa, tids, [chromosomeList[c] for c in contigs], filterScore, max_xM, 1, impl)
    return [ConsRecord(seqs[i], contigs[i], "CI::" + str(int(holes[i])) + "_SP::" + str(int(snps[i]))) for i in range(len(contigs))]


# ----------------------------------------------------------------------------------------------------------------
# Seam S3
# ----------------------------------------------------------------------------------------------------------------

class HammingIndex:
    """Known alleles resident on the GPU as 2-bit planes; rows grouped by (bacterium, gene) in sequencesGetAll order
    (metaMLST_functions.py:224-228: rows of `alleles` by rowid)."""

    def __init__(self, ctx: native.Context, rows: Sequence[Tuple[str, str, int, str]]):
        """rows: (bacterium, gene, alleleVariant, sequence) in table order."""
        self.ctx = ctx
        order = sorted(range(len(rows)), key=lambda i: (rows[i][0], rows[i][1]))  # stable: keeps table order inside a locus
        self.rows = [rows[i] for i in order]
        self.table_pos = np.asarray(order, dtype=np.int64)  # position in table (rowid) order of every resident row
        self.block: Dict[Tuple[str, str], Tuple[int, int]] = {}
        for i, (b, g, _v, _s) in enumerate(self.rows):
            lo, hi = self.block.get((b, g), (i, i))
            self.block[(b, g)] = (lo, i + 1)
        seqs = [r[3].encode("latin-1") for r in self.rows]
        self.W = packing._w_for(max((len(s) for s in seqs), default=1))
        hi, lo, ln, xids, xx, xb = packing.encode_2bit_x(seqs, self.W)  # rows with IUPAC / N / lower case: exact path (H9)
        self._hi, self._lo = packing.tile_db(hi, lo)
        self._len = ln
        self.n_flagged_rows = int(xids.size)
        native.check(native.lib().mmlst_db_upload_x(ctx.handle, native.ptr(self._hi), native.ptr(self._lo), native.ptr(self._len),
                                                    len(self.rows), self.W, native.ptr(xids), native.ptr(xx), native.ptr(xb), int(xids.size)))
        self._keys = self.table_pos.astype(np.uint32)
        self._row_of_key = np.argsort(self.table_pos)
        native.check(native.lib().mmlst_db_row_keys(ctx.handle, native.ptr(self._keys), len(self.rows)))

    @classmethod
    def from_sqlite(cls, ctx, conn, bacterium: Optional[str] = None):
        q = "SELECT bacterium, gene, alleleVariant, sequence FROM alleles"
        args = ()
        if bacterium is not None:
            q += " WHERE bacterium = ?"
            args = (bacterium,)
        return cls(ctx, [(r[0], r[1], r[2], r[3]) for r in conn.execute(q + " ORDER BY recID", args)])

    def search(self, queries: Sequence[str], ranges: Sequence[Tuple[int, int]]):
        """min/argmin of stringDiff(query, row) over rows[ranges[i]] for every query; ties -> lowest row."""
        nq = len(queries)
        if nq == 0:
            return np.zeros(0, np.uint32), np.zeros(0, np.uint32)
        # group queries with identical row ranges into blocks (queries of a block must be contiguous)
        order = sorted(range(nq), key=lambda i: ranges[i])
        qs = [queries[i].encode("latin-1") for i in order]
        if any(len(q) > self.W * 32 for q in qs):
            # zip truncation (H9): columns beyond the longest DB row never take part in a comparison
            qs = [q[: self.W * 32] for q in qs]
        hi, lo, ln, xids, xx, xb = packing.encode_2bit_x(qs, self.W)
        blocks = []
        i = 0
        while i < nq:
            j = i
            while j < nq and ranges[order[j]] == ranges[order[i]]:
                j += 1
            blocks.append((i, j, ranges[order[i]][0], ranges[order[i]][1]))
            i = j
        blk = np.asarray(blocks, dtype=np.uint32).reshape(-1)
        md = np.zeros(nq, np.uint32)
        am = np.zeros(nq, np.uint32)
        native.check(native.lib().mmlst_hamming_min_x(self.ctx.handle, native.ptr(hi), native.ptr(lo), native.ptr(ln), nq,
                                                      native.ptr(xids), native.ptr(xx), native.ptr(xb), int(xids.size),
                                                      native.ptr(blk), len(blocks), native.ptr(md), native.ptr(am)))
        out_d = np.zeros(nq, np.uint32)
        out_a = np.zeros(nq, np.uint32)
        out_d[order] = md
        out_a[order] = am
        return out_d, out_a

    def organism_range(self, bacterium: str) -> Optional[Tuple[int, int]]:
        """Row range of every allele of the organism (rows are grouped by (bacterium, gene): one contiguous range)."""
        lo = [r[0] for (b, _g), r in self.block.items() if b == bacterium]
        hi = [r[1] for (b, _g), r in self.block.items() if b == bacterium]
        return (min(lo), max(hi)) if lo else None

    def exact_first(self, queries: Sequence[str], ranges: Sequence[Tuple[int, int]]) -> np.ndarray:
        """Row a10 -- for every query the lowest row of its range whose sequence EQUALS it (same length, same characters, case-
        sensitive), or 0xFFFFFFFF: `SELECT .. WHERE sequence = ? AND bacterium = ?` + fetchone() (metaMLST_functions.py:168-172,
        196-203, 218-222) with the range = organism_range(bacterium).  Rows are resident grouped by (bacterium, gene); the kernel
        minimises the row's position in TABLE order (mmlst_db_row_keys), so a sequence held by several genes resolves to the
        row fetchone() returns; the value handed back is the resident row index (`self.rows[row]`)."""
        nq = len(queries)
        out = np.full(nq, NO_IDX, np.uint32)
        if nq == 0:
            return out
        order = sorted(range(nq), key=lambda i: ranges[i])
        qs = [queries[i].encode("latin-1") for i in order]
        long = [i for i, q in enumerate(qs) if len(q) > self.W * 32]  # longer than every row: cannot equal any
        if long:
            qs = [q if len(q) <= self.W * 32 else b"" for q in qs]
        hi, lo, ln, xids, _xx, xb = packing.encode_2bit_x(qs, self.W)
        blocks = []
        i = 0
        while i < nq:
            j = i
            while j < nq and ranges[order[j]] == ranges[order[i]]:
                j += 1
            blocks.append((i, j, ranges[order[i]][0], ranges[order[i]][1]))
            i = j
        blk = np.asarray(blocks, dtype=np.uint32).reshape(-1)
        got = np.full(nq, NO_IDX, np.uint32)
        native.check(native.lib().mmlst_exact_match(self.ctx.handle, native.ptr(hi), native.ptr(lo), native.ptr(ln), nq, native.ptr(xids),
                                                    native.ptr(xb), int(xids.size), native.ptr(blk), len(blocks), native.ptr(got)))
        for i in long:
            got[i] = NO_IDX
        for i, q in enumerate(qs):
            if len(q) == 0:  # the empty string equals only an empty DB sequence; never asked by the reference (`geneSeq == ''` short-cuts)
                got[i] = NO_IDX
        hit = got != NO_IDX
        got[hit] = self._row_of_key[got[hit]].astype(np.uint32)  # table position -> resident row
        out[order] = got
        return out

    def closest_allele(self, bacterium: str, gene: str, seq: str) -> Tuple[int, int]:
        """(min distance, alleleVariant) over the rows of (bacterium, gene); flag of metamlst-merge.py:178 = d <= z."""
        rng = self.block.get((bacterium, gene))
        if rng is None:
            return (1 << 30, -1)
        d, a = self.search([seq], [rng])
        return int(d[0]), int(self.rows[int(a[0])][2])


def stringDiff(s1: str, s2: str, ctx: Optional[native.Context] = None) -> int:
    """metaMLST_functions.py:230-234, kept as a name; one pair through the same kernel."""
    ctx = ctx or native.Context(0)
    idx = HammingIndex(ctx, [("x", "x", 1, s2)])
    return int(idx.search([s1], [(0, 1)])[0][0])


def define_profile(conn, geneList: Sequence[str]):
    """Seam S4 -- metaMLST_functions.py:205-216 including H11: unknown labels shrink the denominator, [(0,0)] only
    when the LAST lookup failed.  Stays SQL (ms-scale)."""
    recs = []
    if not geneList:
        # upstream never binds `result` when the loop body does not run and dies on `if result` (metaMLST_functions.py:216)
        raise UnboundLocalError("cannot access local variable 'result' where it is not associated with a value (defineProfile with an empty gene list)")
    result = None
    for allele in geneList:
        result = conn.execute("SELECT recID FROM alleles WHERE bacterium||'_'||gene||'_'||alleleVariant = ?", (allele,)).fetchone()
        if result:
            recs.append(str(result[0]))
    if not result:
        return [(0, 0)]
    inl = ",".join(recs)
    q = ("SELECT profileCode, COUNT(*) as T FROM profiles WHERE alleleCode IN (" + inl + ") GROUP BY profileCode HAVING T = "
         "(SELECT COUNT(*) FROM profiles WHERE alleleCode IN (" + inl + ") GROUP BY profileCode ORDER BY COUNT(*) DESC LIMIT 1) ORDER BY T DESC")
    return [(row[0], int((float(row[1]) / float(len(recs))) * 100)) for row in conn.execute(q)]


class ProfileIndex:
    """Seam S4 on the device (row a11): the `profiles` table grouped by profileCode, resident in the context; label -> recID
    resolved through a host dict built once (first row in table order, what the reference's fetchone() returns)."""

    def __init__(self, ctx: native.Context, conn):
        self.ctx = ctx
        self.rec_of: Dict[str, int] = {}
        for r in conn.execute("SELECT recID, bacterium, gene, alleleVariant FROM alleles ORDER BY recID"):
            self.rec_of.setdefault("%s_%s_%s" % (r[1], r[2], r[3]), int(r[0]))
        groups: Dict[int, List[int]] = {}
        for r in conn.execute("SELECT profileCode, alleleCode FROM profiles"):
            if r[1] is not None:
                groups.setdefault(r[0], []).append(int(r[1]))
        self.codes = sorted(groups)  # GROUP BY profileCode emits the groups in key order
        start = np.zeros(len(self.codes) + 1, np.uint32)
        start[1:] = np.cumsum([len(groups[c]) for c in self.codes])
        flat = np.asarray([a for c in self.codes for a in groups[c]], dtype=np.int64)
        if flat.size and (flat.min() < 0 or flat.max() >= 0xFFFFFFFF):
            raise ValueError("alleleCode outside uint32")
        self._start, self._flat = start, flat.astype(np.uint32)
        native.check(native.lib().mmlst_profiles_upload(ctx.handle, native.ptr(self._start), native.ptr(self._flat), len(self.codes)))

    def define_profiles(self, gene_lists: Sequence[Sequence[str]], max_out: int = 64) -> List[List[Tuple[int, int]]]:
        """defineProfile(conn, geneList) for every list, ONE device call.  H11 kept: labels the DB does not know are dropped (the
        denominator shrinks); [(0, 0)] when the LAST label of a list is unknown (or the list is empty: the reference raises
        NameError there -- callers never pass one)."""
        nq = len(gene_lists)
        if nq == 0:
            return []
        recs = [[self.rec_of[l] for l in gl if l in self.rec_of] for gl in gene_lists]
        lmax = max(1, max(len(r) for r in recs))
        qa = np.full((nq, lmax), 0xFFFFFFFF, np.uint32)
        qn = np.zeros(nq, np.uint32)
        for i, r in enumerate(recs):
            qa[i, :len(r)] = r
            qn[i] = len(r)
        best = np.zeros(nq, np.uint32); nb = np.zeros(nq, np.uint32); out = np.zeros((nq, max_out), np.uint32)
        native.check(native.lib().mmlst_st_match(self.ctx.handle, native.ptr(qa), native.ptr(qn), lmax, nq, native.ptr(best), native.ptr(nb),
                                                 native.ptr(out), max_out))
        if int(nb.max()) > max_out:  # more tied profiles than asked for: once more with room for all of them
            return self.define_profiles(gene_lists, int(nb.max()))
        res = []
        for i, gl in enumerate(gene_lists):
            if not gl or gl[-1] not in self.rec_of:
                res.append([(0, 0)])
                continue
            pct = int((float(best[i]) / float(len(recs[i]))) * 100)
            # ties: `ORDER BY T DESC` over the groups leaves equal counts in DESCENDING profileCode order (SQLite 3.x reads its sorter
            # backwards for DESC; pinned against sqlite3 itself in tests/test_st_match.py) -- the device list is ascending
            res.append([(self.codes[int(p)], pct) for p in out[i, :int(nb[i])][::-1]])
        return res


def fast_select(index: AlleleIndex, sum_as: np.ndarray, n_hit: np.ndarray, first_idx: np.ndarray, penalty: int = 100):
    """Vectorised equivalent of finish_scores + select_alleles for the hot loop (bench, cohort driver): per detected
    locus the chosen allele row, in the reference's dict order (species by first record, loci by first record).

    Exactness: Python's round(x, 1) is monotone in x, so the alleles whose rounded average equals the locus maximum
    all lie within 0.11 of the best raw average; only those few candidates go through Python round() (H6).
    Returns [(species, [tid per locus in dict order])] in species dict order."""
    hit = n_hit > 0
    loc = index.locus_of.astype(np.int64)
    nl = index.n_loci
    maxlen = np.zeros(nl, dtype=np.int64)
    np.maximum.at(maxlen, loc[hit], n_hit[hit].astype(np.int64))
    n = n_hit.astype(np.int64)
    score = sum_as - (maxlen[loc] - n) * penalty
    avg = np.full(n.shape[0], -np.inf)
    avg[hit] = score[hit].astype(np.float64) / n[hit].astype(np.float64)
    best = np.full(nl, -np.inf)
    np.maximum.at(best, loc[hit], avg[hit])
    cand = np.nonzero(hit & (avg >= best[loc] - 0.11))[0]
    chosen: Dict[int, Tuple[float, int, int]] = {}
    for t in cand:
        t = int(t)
        r = round(float(score[t]) / float(n[t]), 1)
        l = int(loc[t])
        a = int(index.allele[t])
        cur = chosen.get(l)
        if cur is None or r > cur[0] or (r == cur[0] and a < cur[1]):
            chosen[l] = (r, a, t)
    lfirst = np.full(nl, NO_IDX, dtype=np.uint32)
    np.minimum.at(lfirst, loc[hit], first_idx[hit])
    out: Dict[str, List[Tuple[int, int]]] = {}
    for l in sorted(chosen, key=lambda l: int(lfirst[l])):
        out.setdefault(index.locus_names[l][0], []).append(chosen[l][2])
    return list(out.items())
