"""BAM -> SoaHost through the native unpacker (`mmlst_bam_unpack`, csrc/bam_unpack.cpp).

Stands where the reference shells out to `samtools view -h -` (metamlst.py:96), opens the file with pysam
(cmseq/cmseq.py:54) and rewrites it with `samtools sort` / `samtools index` (metaMLST_functions.py:237-247).  The
arrays are views of memory owned by the native handle (page-locked when a CUDA device is present), kept alive by the
returned SoaHost.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import native, packing


class UnpackOpts(C.Structure):
    _fields_ = [("minqual", C.c_int), ("max_depth", C.c_uint32), ("sentinel_nodes", C.c_uint32), ("n_threads", C.c_int),
                ("pinned", C.c_int), ("assume_sorted", C.c_int), ("want_qhash", C.c_int), ("check_crc", C.c_int)]


class BamInfo(C.Structure):
    _fields_ = [("soa", native.Soa), ("qhash", C.c_void_p), ("ref_len", C.c_void_p), ("ref_names", C.c_char_p),
                ("header_text", C.c_char_p), ("n_dropped_by_cap", C.c_uint64), ("n_unmapped_flag", C.c_uint64),
                ("presorted", C.c_int), ("minqual", C.c_int), ("max_depth", C.c_uint32), ("seconds", C.c_double * 5)]


class _Handle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            if self.h:
                native.lib().mmlst_bam_free(self.h)
                self.h = None
        except Exception:  # noqa: BLE001
            pass


def _view(addr: Optional[int], n: int, dtype) -> np.ndarray:
    dt = np.dtype(dtype)
    if not addr or n == 0:
        return np.zeros(0, dtype=dt)
    buf = (C.c_uint8 * (n * dt.itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dt, count=n)


def unpack_bam(path: str, minqual: int = packing.DEFAULT_MINQUAL, max_depth: Optional[int] = packing.DEFAULT_MAX_DEPTH,
               presorted: bool = False, threads: int = 0, pinned: bool = True, want_qhash: bool = True,
               sentinel_nodes: int = 1) -> packing.SoaHost:
    """Unpack a BAM once into the score stream + pileup stream (include/mmlst.h).  `presorted` = metamlst.py --presorted
    (file order is kept and must be coordinate order); otherwise records are put in `samtools sort` order and
    `orig_idx` carries the file order stage 1 saw (H5).  max_depth=None disables the htslib depth cap."""
    lib = native.lib()
    o = UnpackOpts(int(minqual), int(max_depth or 0), int(sentinel_nodes), int(threads), 1 if pinned else 0,
                   1 if presorted else 0, 1 if want_qhash else 0, 1)
    h = C.c_void_p()
    native.check(lib.mmlst_bam_unpack(path.encode(), C.byref(o), C.byref(h)))
    keep = _Handle(h)
    info = BamInfo()
    native.check(lib.mmlst_bam_info(h, C.byref(info)))
    s = info.soa
    n, P, n_ref = int(s.n_rec), int(s.n_prec), int(s.n_ref)
    names = info.ref_names.decode("latin-1").split("\n") if n_ref else []
    soa = packing.SoaHost(
        names, _view(info.ref_len, n_ref, np.uint32).astype(np.int32),
        _view(s.tid, n, np.uint32), _view(s.as0, n, np.int16), _view(s.xm3, n, np.uint8), _view(s.qlen, n, np.uint16),
        _view(s.orig_idx, n, np.uint32) if s.orig_idx else None,
        _view(s.p_recs, P, packing.PREC_DTYPE), _view(s.planes, int(s.n_plane_words), np.uint32), int(s.max_row_words),
        _view(s.contig_start, n_ref + 1, np.uint64), int(info.minqual), int(info.max_depth), int(info.n_dropped_by_cap))
    if s.run_tid:
        nr = int(s.n_runs)
        soa.run_tid, soa.run_start = _view(s.run_tid, nr, np.uint32), _view(s.run_start, nr + 1, np.uint32)
        soa.chunk_run = _view(s.chunk_run, (n + 255) // 256, np.uint32)
        soa.chunk_qlen = _view(s.chunk_qlen, (n + 255) // 256, np.uint16) if s.chunk_qlen else None
        if nr > 0.125 * n:  # name-grouped input: the explicit tid form is the smaller one
            soa.run_tid = soa.run_start = soa.chunk_run = soa.chunk_qlen = None
    soa.qhash = _view(info.qhash, 2 * n, np.uint64).reshape(n, 2) if info.qhash else None
    soa.header_text = (info.header_text or b"").decode("latin-1")
    soa.unpack_seconds = dict(zip(("read", "inflate", "parse", "sort", "pack"), [float(x) for x in info.seconds]))
    soa._keep = (keep,)
    return soa
