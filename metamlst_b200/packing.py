"""Host-side packing of alignment records / allele sequences into the HBM layouts of include/mmlst.h.

`pack_table` is the numpy route used for synthetic tables (tests, bench); BAM files go through the C++ unpacker
(`metamlst_b200.bam.unpack_bam`), which emits the same structure.  Everything here is layout work: no result of the
reference is computed on the host except the sequential depth-cap admission (H1), done by the native host function
`mmlst_depth_cap`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import native

BAM_FUNMAP = 0x4
BAM_FPROPER_PAIR = 0x2
DEFAULT_MAX_DEPTH = 8000  # pysam pileup() default reaching cmseq/cmseq.py:527 (H1)
DEFAULT_MINQUAL = 20  # metaMLST_functions.py:258
PLANE_SLACK_WORDS = 8
# mmlst_prec (include/mmlst.h): 16 bytes per pileup-stream record
PREC_DTYPE = np.dtype([("pos", "<i4"), ("row_off", "<u4"), ("reflen", "<u2"), ("as_named", "<i2"), ("xm_named", "u1"), ("pad", "u1"), ("nw", "<u2")])
assert PREC_DTYPE.itemsize == 16


@dataclass
class SoaHost:
    """Score stream + pileup stream on the host (numpy; pin() moves them to page-locked memory)."""

    ref_names: List[str]
    ref_lens: np.ndarray
    # score stream (all records, coordinate-sorted)
    tid: np.ndarray
    as0: np.ndarray
    xm3: np.ndarray
    qlen: np.ndarray
    orig_idx: Optional[np.ndarray]
    # pileup stream (mapped + admitted records, coordinate-sorted): 16-byte records + plane rows
    p_recs: np.ndarray  # PREC_DTYPE [P]
    planes: np.ndarray
    max_row_words: int
    contig_start: np.ndarray  # uint64 [n_ref+1]
    minqual: int = DEFAULT_MINQUAL
    max_depth: int = DEFAULT_MAX_DEPTH
    n_dropped_by_cap: int = 0
    _keep: tuple = ()
    qhash: Optional[np.ndarray] = None  # u64 [n_rec][2] 128-bit QNAME hash (coverage column, H7); None when not unpacked
    header_text: str = ""
    unpack_seconds: Optional[dict] = None
    # run-length form of the score stream (include/mmlst.h, mmlst_score_runs_dev): tid per run of equal tid
    run_tid: Optional[np.ndarray] = None
    run_start: Optional[np.ndarray] = None
    chunk_run: Optional[np.ndarray] = None
    chunk_qlen: Optional[np.ndarray] = None  # u16 per 256-record chunk when every chunk has one len(SEQ) (3 B / record form)
    z_bytes: Optional[np.ndarray] = None     # DEFLATE blocks of as0[] / xm3[] (mmlst_zstream) + their table [n_blocks][4] u64: deflate()
    z_table: Optional[np.ndarray] = None
    z_as_xm_coeff: int = 0                   # the as0 blocks hold as0 + coeff * xm3 (mmlst_zstream.as_xm_coeff)
    zp_bytes: Optional[np.ndarray] = None    # DEFLATE blocks of the pileup stream, contig by contig (mmlst_zpileup): deflate(pileup=True)
    zp_table: Optional[np.ndarray] = None    # [n_blocks][2] u64
    zp_contig_block: Optional[np.ndarray] = None   # [n_ref + 1] u32
    lenient: bool = False                    # unpacked with lenient_tags (bam.unpack_bam): pileup only, the score seams refuse it
    n_untagged: int = 0                      # ... pileup records without integer AS:i / XM:i tags

    @property
    def n_rec(self) -> int:
        return int(self.tid.shape[0])

    @property
    def n_prec(self) -> int:
        return int(self.p_recs.shape[0])

    # column views of the records (tests / tools)
    p_pos = property(lambda self: self.p_recs["pos"])
    p_reflen = property(lambda self: self.p_recs["reflen"])
    p_as = property(lambda self: self.p_recs["as_named"])
    p_xm = property(lambda self: self.p_recs["xm_named"])

    @property
    def p_row_off(self) -> np.ndarray:
        """[P+1] word offsets (last entry = end of the last row)."""
        off = np.zeros(self.n_prec + 1, dtype=np.int64)
        off[:-1] = self.p_recs["row_off"]
        if self.n_prec:
            off[-1] = int(off[-2]) + int(row_words(self.p_recs["nw"][-1:])[0])
        return off

    def c_struct(self) -> native.Soa:
        s = native.Soa()
        s.tid, s.as0, s.xm3, s.qlen = native.ptr(self.tid), native.ptr(self.as0), native.ptr(self.xm3), native.ptr(self.qlen)
        s.orig_idx = native.ptr(self.orig_idx)
        s.n_rec = self.n_rec
        s.p_recs, s.planes = native.ptr(self.p_recs), native.ptr(self.planes)
        s.n_prec = self.n_prec
        s.n_plane_words = int(self.planes.shape[0])
        s.max_row_words = int(self.max_row_words)
        s.contig_start = native.ptr(self.contig_start)
        s.n_ref = len(self.ref_names)
        if self.run_tid is not None and self.n_rec:
            s.n_runs = int(self.run_tid.shape[0])
            s.run_tid, s.run_start, s.chunk_run = native.ptr(self.run_tid), native.ptr(self.run_start), native.ptr(self.chunk_run)
            if self.chunk_qlen is not None:
                s.chunk_qlen = native.ptr(self.chunk_qlen)
            if self.z_bytes is not None:
                zs = native.ZStream(native.ptr(self.z_bytes), int(self.z_bytes.shape[0]), native.ptr(self.z_table), int(self.z_table.shape[0]),
                                    int(self.z_as_xm_coeff))
                s._zs = zs  # kept alive by the struct that points at it
                s.z = C.addressof(zs)
        if self.zp_bytes is not None:
            zp = native.ZPileup(native.ptr(self.zp_bytes), int(self.zp_bytes.shape[0]), native.ptr(self.zp_table), int(self.zp_table.shape[0]),
                                native.ptr(self.zp_contig_block))
            s._zp = zp
            s.zp = C.addressof(zp)
        return s

    def deflate(self, level: int = 3, block: int = 1 << 16, threads: int = 0, pinned: bool = True, strategy: int = 0,
                cover: float = 1.0, pileup: bool = False, as_xm_coeff: Optional[int] = None) -> "SoaHost":
        """Attach the DEFLATE-compressed copy of as0[] / xm3[] (include/mmlst.h, mmlst_zstream): the host-buffer path then ships these
        bytes and the device inflates them with the hardware decompression engine.  Done once per sample, like the unpacking; needs the
        run-length form (coordinate-sorted streams).  block = inflated bytes per DEFLATE stream: the engine works on many streams at once,
        64 KiB blocks (a BGZF block's size) run at its full rate, 1 MiB blocks at a third of it (B200: 1.56 ms against 2.56 ms for the
        120 MB of configs[1]; the plain arrays take 2.40 ms over PCIe; profiles/r2o_e2e_breakdown.json).  strategy = zlib strategy (0 default,
        Z_FIXED, Z_RLE, Z_HUFFMAN_ONLY).  cover < 1: only the first `cover` of each array is compressed and the rest crosses PCIe plain, so that
        the bus keeps working while the engine (the slower of the two on this data) drains its queue.  pileup=True additionally attaches the pileup
        stream as DEFLATE blocks per contig (mmlst_zpileup: `mmlst_sample` then ships the chosen contigs' blocks instead of their records and
        plane rows).  as_xm_coeff: the as0 blocks hold as0 + coeff * xm3 (an alignment score is a match bonus minus a penalty per mismatch: with the
        aligner's penalty as coefficient the remainder takes a handful of values; the device subtracts it again); None = try 0..8 on a slice of the
        arrays and keep the smallest, 0 = off."""
        import os
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        self.z_bytes = self.z_table = None
        self.zp_bytes = self.zp_table = self.zp_contig_block = None
        if pileup and self.n_prec:
            self._deflate_pileup(level, block, threads, pinned, strategy)
        if self.run_tid is None or self.n_rec == 0:
            return self
        self.z_as_xm_coeff = 0
        as0 = np.ascontiguousarray(self.as0)
        stop_a = 2 * as0.shape[0] if cover >= 1.0 else int(2 * as0.shape[0] * max(cover, 0.0)) // block * block
        n_tr = stop_a // 2   # records of as0[] that travel in blocks
        if n_tr and as_xm_coeff != 0:
            xm = np.ascontiguousarray(self.xm3)[:n_tr].astype(np.int32)
            a32 = as0[:n_tr].astype(np.int32)
            if as_xm_coeff is None:   # smallest zlib output over a slice from the middle of the arrays
                m0, m1 = max(0, n_tr // 2 - (1 << 17)), min(n_tr, n_tr // 2 + (1 << 17))
                best = None
                for c in range(0, 9):
                    t = a32[m0:m1] + c * xm[m0:m1]
                    if t.max(initial=0) > 32767 or t.min(initial=0) < -32768:
                        continue
                    size = len(zlib.compress(t.astype(np.int16).tobytes(), 1))
                    if best is None or size < best[0]:
                        best = (size, c)
                as_xm_coeff = best[1] if best else 0
            if as_xm_coeff:
                t = a32 + int(as_xm_coeff) * xm
                if t.max(initial=0) <= 32767 and t.min(initial=0) >= -32768:
                    as0 = np.concatenate([t.astype(np.int16), as0[n_tr:]])
                    self.z_as_xm_coeff = int(as_xm_coeff)
        jobs = []
        for kind, arr in ((0, as0.view(np.uint8)), (1, np.ascontiguousarray(self.xm3).view(np.uint8))):
            mv = memoryview(arr)
            stop = arr.shape[0] if cover >= 1.0 else int(arr.shape[0] * max(cover, 0.0)) // block * block
            for off in range(0, stop, block):
                jobs.append((kind, off, mv[off:min(off + block, stop)]))
        if not jobs:
            return self

        def one(job):
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
            return co.compress(job[2]) + co.flush()
        with ThreadPoolExecutor(threads or (os.cpu_count() or 1)) as pool:
            comp = list(pool.map(one, jobs))
        table = np.zeros((len(jobs), 4), np.uint64)
        pos = 0
        for i, ((kind, off, raw), c) in enumerate(zip(jobs, comp)):
            table[i] = (kind, off, pos, (len(c) << 32) | len(raw))
            pos += len(c)
        z = np.frombuffer(b"".join(comp), np.uint8)
        if pinned:
            import torch
            t = torch.from_numpy(z.copy()).pin_memory()
            self._keep = tuple(self._keep) + (t,)
            z = t.numpy()
        self.z_bytes, self.z_table = z, table
        return self

    def _deflate_pileup(self, level: int, block: int, threads: int, pinned: bool, strategy: int) -> None:
        import os
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        cs = np.asarray(self.contig_start, dtype=np.int64)
        recs = np.ascontiguousarray(self.p_recs)
        rec_bytes = memoryview(recs.view(np.uint8).reshape(-1))
        plane_bytes = memoryview(np.ascontiguousarray(self.planes).view(np.uint8).reshape(-1))
        row_off = recs["row_off"].astype(np.int64)             # first plane word of every record
        row_end = row_off + row_words(recs["nw"]).astype(np.int64)
        jobs = []          # (contig, kind, bytes)
        per_contig = np.zeros(len(cs), np.int64)
        for t in np.nonzero(cs[1:] > cs[:-1])[0]:
            r0, r1 = int(cs[t]), int(cs[t + 1])
            for kind, mv, lo, hi in ((0, plane_bytes, int(row_off[r0]) * 4, int(row_end[r1 - 1]) * 4), (1, rec_bytes, r0 * 16, r1 * 16)):
                for off in range(lo, hi, block):
                    jobs.append((int(t), kind, mv[off:min(off + block, hi)]))
                    per_contig[t + 1] += 1

        def one(job):
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
            return co.compress(job[2]) + co.flush()
        with ThreadPoolExecutor(threads or (os.cpu_count() or 1)) as pool:
            comp = list(pool.map(one, jobs))
        table = np.zeros((len(jobs), 2), np.uint64)
        pos = 0
        for i, ((_t, kind, raw), c) in enumerate(zip(jobs, comp)):
            table[i] = (pos, (kind << 63) | (len(c) << 32) | len(raw))
            pos += len(c)
        z = np.frombuffer(b"".join(comp) + b"\0" * 64, np.uint8)
        if pinned:
            import torch
            t = torch.from_numpy(z.copy()).pin_memory()
            self._keep = tuple(self._keep) + (t,)
            z = t.numpy()
        self.zp_bytes, self.zp_table, self.zp_contig_block = z, table, np.cumsum(per_contig).astype(np.uint32)

    def build_runs(self, max_fraction: float = 0.125) -> "SoaHost":
        """Attach the run-length form (mmlst_build_runs) when it is the smaller one: at most `max_fraction` runs per
        record (coordinate-sorted streams; a name-grouped stream keeps the explicit tid form)."""
        n = self.n_rec
        self.run_tid = self.run_start = self.chunk_run = self.chunk_qlen = None
        if n == 0:
            return self
        import ctypes as C
        lib = native.lib()
        nr = C.c_uint32(0)
        tid = np.ascontiguousarray(self.tid, dtype=np.uint32)
        native.check(lib.mmlst_build_runs(native.ptr(tid), n, None, None, None, C.byref(nr)))
        if nr.value > max_fraction * n:
            return self
        run_tid = np.empty(nr.value, np.uint32); run_start = np.empty(nr.value + 1, np.uint32); chunk_run = np.empty((n + 255) // 256, np.uint32)
        native.check(lib.mmlst_build_runs(native.ptr(tid), n, native.ptr(run_tid), native.ptr(run_start), native.ptr(chunk_run), C.byref(nr)))
        self.run_tid, self.run_start, self.chunk_run = run_tid, run_start, chunk_run
        # len(SEQ) once per chunk when every 256-record chunk is uniform (mmlst_chunk_qlen)
        cq = np.empty((n + 255) // 256, np.uint16)
        uniform = C.c_int(0)
        ql = np.ascontiguousarray(self.qlen, dtype=np.uint16)
        native.check(lib.mmlst_chunk_qlen(native.ptr(ql), n, native.ptr(cq), C.byref(uniform)))
        self.chunk_qlen = cq if uniform.value else None
        return self

    def pin(self) -> "SoaHost":
        """Copy the streams into page-locked memory (torch pinned tensors) so uploads are asynchronous DMA."""
        import torch

        keep = []
        for name in ("tid", "as0", "xm3", "qlen", "orig_idx", "p_recs", "planes", "qhash", "run_tid", "run_start", "chunk_run", "chunk_qlen"):
            arr = getattr(self, name)
            if arr is None:
                continue
            t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).pin_memory()
            keep.append(t)
            setattr(self, name, t.numpy().view(arr.dtype).reshape(arr.shape))
        self._keep = tuple(keep)
        return self


def touched_words(pos: np.ndarray, reflen: np.ndarray) -> np.ndarray:
    """32-column contig words a record touches: ((pos & 31) + reflen + 31) >> 5, 0 for an empty span."""
    pos = pos.astype(np.int64)
    reflen = reflen.astype(np.int64)
    return np.where(reflen > 0, ((pos & 31) + reflen + 31) >> 5, 0)


def row_words(nw: np.ndarray) -> np.ndarray:
    """3 planes x nw words, padded to an odd word count (bank-conflict-free row stride)."""
    rw = 3 * nw.astype(np.int64)
    return rw + ((rw > 0) & ((rw & 1) == 0))


def project_cigars(cig_off: np.ndarray, cig_ops: np.ndarray, n: int):
    """CIGAR -> reference projection: per record reflen and, for every reference offset, the query index of the
    base aligned there (-1 for D/N).  M,=,X consume both; I,S query only; D,N reference only; H,P nothing (htslib
    resolve_cigar2 as restated in oracle/mlst_oracle.py).  Returns (reflen[n], rec_of_seg, ref_start, q_start, seg_len)."""
    ncig = (cig_off[1:] - cig_off[:-1]).astype(np.int64)
    rec = np.repeat(np.arange(n, dtype=np.int64), ncig)
    op = (cig_ops & 0xF).astype(np.int64)
    ln = (cig_ops >> 4).astype(np.int64)
    cons_ref = np.isin(op, (0, 2, 3, 7, 8))
    cons_q = np.isin(op, (0, 1, 4, 7, 8))
    # exclusive prefix sums inside each record
    def seg_excl_cumsum(v):
        c = np.cumsum(v)
        start = c - v
        base = np.zeros(n + 1, dtype=np.int64)
        np.add.at(base, rec + 1, v)  # totals per record
        tot = base[1:]
        first = np.cumsum(tot) - tot
        return start - first[rec], tot
    ref_start, reflen = seg_excl_cumsum(ln * cons_ref)
    q_start, _ = seg_excl_cumsum(ln * cons_q)
    is_m = np.isin(op, (0, 7, 8))
    return reflen, rec[is_m], ref_start[is_m], q_start[is_m], ln[is_m]


def _pack_bits_u32(bits: np.ndarray) -> np.ndarray:
    """bool [n, 32*k] -> uint32 [n, k], bit i of word j = column 32 j + i."""
    return np.packbits(bits, axis=1, bitorder="little").view(np.uint32)


def pack_table(tab, minqual: int = DEFAULT_MINQUAL, max_depth: Optional[int] = DEFAULT_MAX_DEPTH, sentinel_nodes: int = 1,
               chunk: int = 1 << 18, run_fraction: float = 0.125) -> SoaHost:
    """AlnTable (fixed read length, ASCII seq + phred qual) -> SoaHost.  max_depth=None disables the htslib cap.
    run_fraction: attach the run-length form of the score stream when runs <= run_fraction * records (0 = never,
    1 = always)."""
    n = tab.n
    if n and np.any(tab.flag & BAM_FPROPER_PAIR):
        raise native.MmlstError(-6, "proper-pair records: htslib overlap handling (H2) is not implemented -- refusing")
    if n and np.any(tab.tid < 0):
        raise native.MmlstError(-4, "unmapped record (RNAME '*'): the reference crashes at metamlst.py:107")
    if n and (tab.AS.min() < -32768 or tab.AS.max() > 32767):
        raise native.MmlstError(-7, "AS outside int16")
    order = tab.coord_order()
    presorted = bool(np.all(order == np.arange(n)))
    L = tab.read_len
    # positional aux fields (H4): field 0 is AS; field 3 is XM when XS:i is present, else XO
    aux3 = np.where(tab.has_xs, tab.XM, tab.XO)
    tid = tab.tid[order].astype(np.uint32)
    as0 = tab.AS[order].astype(np.int16)
    xm3 = np.clip(aux3[order], 0, 255).astype(np.uint8)
    if np.any(aux3 < 0):
        raise native.MmlstError(-7, "negative 4th aux field")
    qlen = np.full(n, max(L, 1), dtype=np.uint16)
    orig_idx = None if presorted else order.astype(np.uint32)

    # ---- pileup stream
    reflen_all, seg_rec, seg_ref, seg_q, seg_len = project_cigars(tab.cig_off, tab.cig_ops, n)
    if n and reflen_all.max() > 65535:
        raise native.MmlstError(-7, "reference span > 65535")
    mapped = (tab.flag[order] & BAM_FUNMAP) == 0
    s_idx = order[mapped]  # table indices of pileup candidates, coordinate-sorted
    p_tid = tab.tid[s_idx].astype(np.uint32)
    p_pos_all = tab.pos[s_idx].astype(np.int32)
    p_reflen_all = reflen_all[s_idx].astype(np.uint32)
    admitted = np.ones(s_idx.shape[0], dtype=np.uint8)
    if max_depth is not None and s_idx.shape[0]:
        native.check(native.lib().mmlst_depth_cap(native.ptr(p_tid), native.ptr(p_pos_all), native.ptr(p_reflen_all),
                                                   s_idx.shape[0], int(max_depth), int(sentinel_nodes), native.ptr(admitted)))
    adm = admitted.astype(bool)
    sel = s_idx[adm]
    P = sel.shape[0]
    p_pos = p_pos_all[adm]
    p_reflen = p_reflen_all[adm].astype(np.uint16)
    p_as = tab.AS[sel].astype(np.int16)  # by NAME
    p_xm = np.clip(tab.XM[sel], 0, 255).astype(np.uint8)
    p_nw = touched_words(p_pos, p_reflen)
    if P and p_nw.max() > 65535:
        raise native.MmlstError(-7, "record touches more than 65535 words")
    p_shift = p_pos.astype(np.int64) & 31  # rows are aligned to the contig's 32-column words
    rw = row_words(p_nw)
    p_row_off = np.zeros(P + 1, dtype=np.int64)
    p_row_off[1:] = np.cumsum(rw)
    if p_row_off[-1] + PLANE_SLACK_WORDS >= (1 << 32):
        raise native.MmlstError(-7, "plane array exceeds 2^32 words")
    planes = np.zeros(int(p_row_off[-1]) + PLANE_SLACK_WORDS, dtype=np.uint32)
    # segments grouped by table record for fast lookup
    seg_order = np.argsort(seg_rec, kind="stable")
    seg_rec, seg_ref, seg_q, seg_len = seg_rec[seg_order], seg_ref[seg_order], seg_q[seg_order], seg_len[seg_order]
    seg_first = np.searchsorted(seg_rec, np.arange(n + 1))
    code_lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        code_lut[ch] = i
        code_lut[ch + 32] = i  # query_sequence[..].upper() (cmseq/cmseq.py:537)
    for c0 in range(0, P, chunk):
        idx = sel[c0:c0 + chunk]
        m = idx.shape[0]
        wmax = int(p_nw[c0:c0 + m].max()) if m else 0
        if wmax == 0:
            continue
        qidx = np.full((m, wmax * 32), -1, dtype=np.int64)
        # expand the M segments of these records
        s0, s1 = seg_first[idx], seg_first[idx + 1]
        nseg = s1 - s0
        sid = np.repeat(s0, nseg) + (np.arange(int(nseg.sum())) - np.repeat(np.cumsum(nseg) - nseg, nseg))
        row = np.repeat(np.arange(m), nseg)
        sl = seg_len[sid]
        tot = int(sl.sum())
        within = np.arange(tot) - np.repeat(np.cumsum(sl) - sl, sl)
        rr = np.repeat(row, sl)
        qidx[rr, np.repeat(seg_ref[sid], sl) + within + p_shift[c0:c0 + m][rr]] = np.repeat(seg_q[sid], sl) + within
        has = qidx >= 0
        qi = np.where(has, qidx, 0)
        base = tab.seq[idx][np.arange(m)[:, None], qi]
        ql = tab.qual[idx][np.arange(m)[:, None], qi]
        code = code_lut[base]
        qok = has & (ql >= minqual)  # H3: bases under minqual are not in the column at all
        V = qok & (code != 255)
        Nn = qok & (code == 255)
        B1 = V & ((code & 2) != 0)
        B0 = (V & ((code & 1) != 0)) | Nn
        Vw, B1w, B0w = _pack_bits_u32(V), _pack_bits_u32(B1), _pack_bits_u32(B0)
        nw = p_nw[c0:c0 + m]
        off = p_row_off[c0:c0 + m]
        for j in range(wmax):
            msk = nw > j
            o = off[msk] + 3 * j
            planes[o] = Vw[msk, j]
            planes[o + 1] = B1w[msk, j]
            planes[o + 2] = B0w[msk, j]
    n_ref = len(tab.ref_names)
    contig_start = np.searchsorted(tab.tid[sel], np.arange(n_ref + 1)).astype(np.uint64)
    recs = np.zeros(P, dtype=PREC_DTYPE)
    recs["pos"], recs["row_off"], recs["reflen"], recs["as_named"], recs["xm_named"] = p_pos, p_row_off[:-1], p_reflen, p_as, p_xm
    recs["nw"] = p_nw
    soa = SoaHost(list(tab.ref_names), np.asarray(tab.ref_lens, dtype=np.int32), tid, as0, xm3, qlen, orig_idx,
                  recs, planes, int(rw.max()) if P else 0, contig_start,
                  minqual, max_depth if max_depth is not None else 0, int((~adm).sum()))
    soa.qhash = qname_key(tab.qname_id[order])
    if run_fraction > 0:
        soa.build_runs(run_fraction)
    return soa


def qname_key(qname_id: np.ndarray) -> np.ndarray:
    """128-bit name key [n][2] of a synthetic table whose QNAMEs are "r<id>": the id itself (injective, so exact -- the
    BAM unpacker hashes the name bytes instead; only equality of keys matters to the coverage kernel)."""
    k = np.zeros((qname_id.shape[0], 2), dtype=np.uint64)
    k[:, 0] = qname_id.astype(np.uint64)
    k[:, 1] = np.uint64(0x51ED270B0B1F2A37)
    return k


# ----------------------------------------------------------------------------------------------------------------
# Allele database for the Hamming search
# ----------------------------------------------------------------------------------------------------------------

def _w_for(max_len: int) -> int:
    w = (max_len + 31) // 32
    for cand in (8, 16, 24, 32):
        if w <= cand:
            return cand
    if max_len < 0x8000:  # lengths are 15 bits (bit 15 flags non-ACGT sequences); widths beyond 32 words take the any-width kernel
        return (w + 7) // 8 * 8
    raise native.MmlstError(-7, "sequence of %d bases exceeds 32767" % max_len)


def encode_2bit_x(seqs: Sequence[bytes], W: int):
    """ASCII sequences -> (hi[n, W], lo[n, W], len[n] u16, xids[m] u32, xx[m, W] u32, xbytes[m, W*32] u8).
    Bit-planes of the code A=0 C=1 G=2 T=3, upper case only: stringDiff compares characters
    (metaMLST_functions.py:230-234).  A sequence holding any other character (IUPAC, 'N', lower case, H9) is FLAGGED
    with bit 15 of its length and listed in xids with the bit-plane of its exceptional columns (xx) and its ASCII bytes
    (xbytes): the exact kernel compares those (include/mmlst.h)."""
    n = len(seqs)
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
    if n and lens.max() > W * 32:
        raise native.MmlstError(-7, "sequence longer than W*32")
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    mat = np.zeros((n, W * 32), dtype=np.uint8)
    exc = np.zeros((n, W * 32), dtype=bool)
    flat = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    rows = np.repeat(np.arange(n), lens)
    cols = np.arange(flat.shape[0]) - np.repeat(np.cumsum(lens) - lens, lens)
    codes = lut[flat]
    bad = codes == 255
    mat[rows, cols] = np.where(bad, 0, codes)
    exc[rows[bad], cols[bad]] = True
    hi = _pack_bits_u32((mat & 2) != 0)
    lo = _pack_bits_u32((mat & 1) != 0)
    ln = lens.astype(np.uint16)
    xids = np.nonzero(exc.any(axis=1))[0].astype(np.uint32)
    ln[xids] |= np.uint16(0x8000)
    xx = np.ascontiguousarray(_pack_bits_u32(exc[xids])) if xids.size else np.zeros((0, W), np.uint32)
    xbytes = np.zeros((xids.size, W * 32), dtype=np.uint8)
    for k, i in enumerate(xids):
        xbytes[k, : lens[i]] = np.frombuffer(seqs[int(i)], dtype=np.uint8)
    return np.ascontiguousarray(hi), np.ascontiguousarray(lo), ln, xids, xx, xbytes


def encode_2bit(seqs: Sequence[bytes], W: int):
    """Strict form for clean (upper-case ACGT) sequences: (hi, lo, len); refuses anything the 2-bit planes cannot hold."""
    hi, lo, ln, xids, _xx, _xb = encode_2bit_x(seqs, W)
    if xids.size:
        raise native.MmlstError(-7, "non-ACGT letter in sequence %d: use encode_2bit_x (exact-character path, H9)" % int(xids[0]))
    return hi, lo, ln


def tile_db(hi: np.ndarray, lo: np.ndarray):
    """[n, W] row-major planes -> 32-row word-major tiles: out[(tile*W + w)*32 + r]."""
    n, W = hi.shape
    nt = (n + 31) // 32
    def t(x):
        p = np.zeros((nt * 32, W), dtype=np.uint32)
        p[:n] = x
        return np.ascontiguousarray(p.reshape(nt, 32, W).transpose(0, 2, 1)).reshape(-1)
    return t(hi), t(lo)
