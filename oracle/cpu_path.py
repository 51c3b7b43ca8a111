"""TEST INFRASTRUCTURE ONLY -- the hot path on the HOST cores over a whole synthetic sample, through the C restatement
(oracle/c/mlst_oracle.c) and the plain-Python selection below.  Used by bench.py's `cpu_baseline` / `--impl reference`
legs and by the full-workload parity check (bench.py, tests/): never imported by the product.

The workload is held UNPACKED, the way the reference sees a BAM: one row per alignment record (tid, pos, positional aux
fields, named AS / XM), SEQ / QUAL / CIGAR once per read (bowtie2 -k emits K records per read), ASCII bases and phred
bytes.  Nothing of the product's packing (run-length score stream, bit-planes, depth cap resolved at unpack) is shared:
the depth cap is the htslib event simulation (orc_depth_cap_sim), run inside the timed region on every record of a
chosen contig, exactly the work pysam does per `pileup()` call (cmseq/cmseq.py:527).

Reference lines: stage 1 metamlst.py:101-130, aggregation :133-151, best allele :244, consensus driver
metaMLST_functions.py:249-281, pileup cmseq/cmseq.py:527-569.
"""
from __future__ import annotations

import ctypes as C
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import corc

NO_IDX = 0xFFFFFFFF


def cel_from_tables(ref_names: Sequence[str], sum_as: np.ndarray, n_hit: np.ndarray, first_idx: np.ndarray, penalty: int = 100):
    """metamlst.py:133-151 on integer tables: cel[species][gene][allele] = (localScore, n, round(avg, 1)); dict order =
    first passing record of species / gene / allele (H5)."""
    hit = np.nonzero(n_hit)[0]
    hit = hit[np.argsort(first_idx[hit], kind="stable")]
    lists: Dict[str, Dict[str, Dict[str, Tuple[int, int]]]] = {}
    for t in hit.tolist():
        species, gene, allele = ref_names[t].split("_")
        lists.setdefault(species, {}).setdefault(gene, {})[allele] = (int(sum_as[t]), int(n_hit[t]))
    cel: Dict[str, Dict[str, Dict[str, tuple]]] = {}
    for species, genes in lists.items():
        cel[species] = {}
        for gene, info in genes.items():
            maxLen = max(n for (_s, n) in info.values())
            row = {}
            for allele, (localScore, geneLen) in info.items():
                if geneLen != maxLen:
                    localScore = localScore - (maxLen - geneLen) * penalty
                row[allele] = (localScore, geneLen, round(float(localScore) / float(geneLen), 1))
            cel[species][gene] = row
    return cel


def chosen_contigs(cel) -> List[Tuple[str, List[str]]]:
    """metamlst.py:244 per species in dict order: lowest int(allele) among the alleles whose rounded average is the max."""
    out = []
    for species, genes in cel.items():
        names = []
        for gene, info in genes.items():
            best = max(avg for (_v, _l, avg) in info.values())
            a = sorted((k for k, (_v, _l, avg) in info.items() if avg == best), key=lambda x: int(x))[0]
            names.append(species + "_" + gene + "_" + a)
        out.append((species, names))
    return out


@dataclass
class CpuWorkload:
    """A coordinate-sorted sample (what `samtools sort` leaves), unpacked, on the host."""
    ref_names: List[str]
    ref_lens: np.ndarray          # int32 [n_ref]
    db_seq: List[bytes]           # DB sequence per allele row (the chromosomeList values, metamlst.py:244-247)
    L: int
    # per record, sorted by (tid, pos, reverse strand), stable
    tid: np.ndarray               # int32
    aux0: np.ndarray              # int32: 1st aux field by POSITION (metamlst.py:109)
    aux3: np.ndarray              # int32: 4th aux field by POSITION (:110)
    qlen: np.ndarray              # int32 len(SEQ)
    pos: np.ndarray               # int32
    reflen: np.ndarray            # int32
    as_named: np.ndarray          # int32 AS:i by name (metaMLST_functions.py:259)
    xm_named: np.ndarray          # int32 XM:i by name
    read_of: np.ndarray           # int64 -> row of bases / qual / cig3
    contig_start: np.ndarray      # int64 [n_ref+1] record range of each contig (the .bai of the sorted BAM)
    # per read
    bases: np.ndarray             # uint8 [n_reads, L] ASCII
    qual: np.ndarray              # uint8 [n_reads, L]
    cig3: np.ndarray              # uint32 [n_reads, 3] BAM CIGAR words
    n_cig: np.ndarray             # uint8 [n_reads]

    @property
    def n(self) -> int:
        return int(self.tid.shape[0])


def workload_from_cores(db, cores: List[dict]) -> CpuWorkload:
    """Same records, same order as metamlst_b200.devpack.pack_cores makes of the same `cores` (synth.gen_core chunks of one
    sample): K records per read, `samtools sort` order (stable by tid, pos, reverse strand)."""
    import torch
    K, L = cores[0]["K"], cores[0]["L"]
    tid = torch.cat([c["rows"].reshape(-1) for c in cores])
    pos = torch.cat([c["start"][:, None].expand(-1, K).reshape(-1) for c in cores])
    rev = torch.cat([((c["flag"] >> 4) & 1).reshape(-1) for c in cores])
    AS = torch.cat([c["AS"].reshape(-1) for c in cores])
    xm = torch.cat([c["xm"].reshape(-1) for c in cores])
    rtype = torch.cat([c["rtype"] for c in cores])
    a = torch.cat([c["a_split"] for c in cores])
    n = int(tid.shape[0])
    n_reads = n // K
    key = (tid << 33) | ((pos + 1) << 1) | rev
    order = torch.sort(key, stable=True).indices
    read_of = (order // K)
    span = torch.full((n_reads,), L, dtype=torch.int64, device=tid.device)
    span[rtype == 1] = L - 10
    span[rtype == 2] = L - 1
    span[rtype == 3] = L + 1
    M, I, D, S = 0, 1, 2, 4
    cig = torch.zeros((n_reads, 3), dtype=torch.int64, device=tid.device)
    cig[:, 0] = (L << 4) | M
    m1, m2, m3 = rtype == 1, rtype == 2, rtype == 3
    cig[m1] = torch.tensor([(5 << 4) | S, ((L - 10) << 4) | M, (5 << 4) | S], device=tid.device)
    cig[m2, 0] = (a[m2] << 4) | M; cig[m2, 1] = (1 << 4) | I; cig[m2, 2] = ((L - 1 - a[m2]) << 4) | M
    cig[m3, 0] = (a[m3] << 4) | M; cig[m3, 1] = (1 << 4) | D; cig[m3, 2] = ((L - a[m3]) << 4) | M
    ncig = torch.where(rtype == 0, 1, 3)

    def h(x, dt):
        return np.ascontiguousarray(x.detach().cpu().numpy().astype(dt, copy=False))

    tid_s = h(tid[order], np.int32)
    n_ref = db.n_rows
    w = CpuWorkload(
        ref_names=db.ref_names(), ref_lens=db.row_len().astype(np.int32),
        db_seq=[db.seq[db.seq_off[r]:db.seq_off[r + 1]].tobytes() for r in range(n_ref)], L=L,
        tid=tid_s, aux0=h(AS[order], np.int32), aux3=h(xm[order], np.int32),  # synthetic records carry XS:i => 4th aux is XM
        qlen=np.full(n, L, np.int32), pos=h(pos[order], np.int32), reflen=h(span[read_of], np.int32),
        as_named=h(AS[order], np.int32), xm_named=h(xm[order], np.int32), read_of=h(read_of, np.int64),
        contig_start=np.searchsorted(tid_s, np.arange(n_ref + 1)).astype(np.int64),
        bases=np.ascontiguousarray(torch.cat([c["bases"] for c in cores]).cpu().numpy()),
        qual=np.ascontiguousarray(torch.cat([c["qual"] for c in cores]).cpu().numpy()),
        cig3=h(cig, np.uint32), n_cig=h(ncig, np.uint8))
    return w


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def score(w: CpuWorkload, allow: np.ndarray, locus_of: np.ndarray, minscore: int, max_xm: int, min_read_len: int, threads: int = 1,
          pool: Optional[ThreadPoolExecutor] = None):
    """metamlst.py:101-130 over the whole record table: `threads` contiguous record ranges, partial tables added / min-ed."""
    n_ref = len(w.ref_names)
    lib = corc.lib()
    bounds = np.linspace(0, w.n, threads + 1).astype(np.int64)
    allow_a = np.ascontiguousarray(allow, np.uint8)
    locus_a = np.ascontiguousarray(locus_of, np.uint32)

    def part(i):
        b0, b1 = int(bounds[i]), int(bounds[i + 1])
        s = np.zeros(n_ref, np.int64); c = np.zeros(n_ref, np.uint32); f = np.full(n_ref, NO_IDX, np.uint32); k = np.zeros(2, np.uint64)
        # file order index == position in the sorted stream (a --presorted BAM): idx_base = b0
        lib.orc_score(C.c_uint64(b1 - b0), _p(w.tid[b0:b1]), _p(w.aux0[b0:b1]), _p(w.aux3[b0:b1]), _p(w.qlen[b0:b1]), _p(None), _p(allow_a),
                      _p(locus_a), minscore, max_xm, min_read_len, _p(s), _p(c), _p(f), _p(k), C.c_uint64(b0))
        return s, c, f, k

    res = list(pool.map(part, range(threads))) if (pool is not None and threads > 1) else [part(i) for i in range(threads)]
    s = sum(r[0] for r in res)
    c = sum(r[1].astype(np.int64) for r in res).astype(np.uint32)
    f = np.minimum.reduce([r[2] for r in res])
    k = sum(r[3] for r in res)
    return s, c, f, k


def contig_consensus(w: CpuWorkload, t: int, minqual: int, minscore: int, max_xm: int, max_depth: Optional[int], mincov: int = 1,
                     sentinel_nodes: int = 1):
    """One chosen contig: htslib depth-cap simulation over ALL its records, pileup of the admitted ones, consensus vs the DB
    allele.  Returns (consensus str, holes, snps, counts, n_admitted)."""
    lib = corc.lib()
    b0, b1 = int(w.contig_start[t]), int(w.contig_start[t + 1])
    n = b1 - b0
    clen = int(w.ref_lens[t])
    counts = np.zeros((clen, 5), np.uint32)
    admitted = np.ones(max(n, 1), np.uint8)
    if n:
        pos = w.pos[b0:b1]
        if max_depth:
            rc = lib.orc_depth_cap_sim(C.c_uint64(n), C.c_int32(t), _p(pos), _p(w.reflen[b0:b1]), C.c_uint32(max_depth), C.c_uint32(sentinel_nodes),
                                       _p(admitted))
            if rc != 0:
                raise ValueError("The input is not sorted (reads out of order)")
        lib.orc_pileup_reads(C.c_uint64(n), _p(pos), _p(w.read_of[b0:b1]), _p(w.cig3), _p(w.n_cig), _p(w.bases), _p(w.qual), C.c_int(w.L),
                             _p(w.as_named[b0:b1]), _p(w.xm_named[b0:b1]), _p(admitted), minqual, minscore, max_xm, C.c_int32(clen), _p(counts))
    cons, holes, snps = corc.consensus(counts, w.db_seq[t], mincov)
    return cons, holes, snps, counts, int(admitted[:n].sum())


def run(w: CpuWorkload, minscore: int = 80, max_xM: int = 5, min_read_len: int = 50, penalty: int = 100, minqual: int = 20,
        max_depth: Optional[int] = 8000, threads: int = 1, species_filter: Optional[str] = None, pool: Optional[ThreadPoolExecutor] = None):
    """One pass of the hot path on the host.  Returns a dict: tables (sum_as, n_hit, first_idx, counters), `result`
    {species: [(contig, consensus, holes, snps)]} in the reference's dict order, and the seconds of each phase."""
    own = pool is None and threads > 1
    if own:
        pool = ThreadPoolExecutor(threads)
    try:
        names = w.ref_names
        sp = [n.split("_")[0] for n in names] if species_filter else None
        keep = set(species_filter.split(",")) if species_filter else None
        allow = np.ones(len(names), np.uint8) if not species_filter else np.fromiter((1 if s in keep else 0 for s in sp), np.uint8, len(names))
        t0 = time.perf_counter()
        s, c, f, k = score(w, allow, np.zeros(len(names), np.uint32), minscore, max_xM, min_read_len, threads, pool)
        t1 = time.perf_counter()
        cel = cel_from_tables(names, s, c, f, penalty)
        chosen = chosen_contigs(cel)
        t2 = time.perf_counter()
        name2tid = {n: i for i, n in enumerate(names)}
        tids = [name2tid[n] for _sp, ns in chosen for n in ns]
        fn = lambda t: contig_consensus(w, t, minqual, minscore, max_xM, max_depth)
        per = list(pool.map(fn, tids)) if (pool is not None and threads > 1) else [fn(t) for t in tids]
        t3 = time.perf_counter()
        result: Dict[str, list] = {}
        i = 0
        for species, ns in chosen:
            for n in ns:
                cons, holes, snps, _cnt, _adm = per[i]
                result.setdefault(species, []).append((n, cons, holes, snps))
                i += 1
        return {"sum_as": s, "n_hit": c, "first_idx": f, "counters": k, "cel": cel, "result": result,
                "piled_records": int(sum(p[4] for p in per)), "chosen_contig_records": int(sum(int(w.contig_start[t + 1] - w.contig_start[t]) for t in tids)),
                "seconds": {"score": t1 - t0, "select": t2 - t1, "depthcap_pileup_consensus": t3 - t2, "total": t3 - t0}}
    finally:
        if own:
            pool.shutdown()
