"""Device-side (torch) packing of synthetic reads straight into the HBM layouts -- bench/test plumbing only.

`synth.gen_core` makes reads as torch tensors; `pack_cores` turns them into the coordinate-sorted score stream and
pileup stream of include/mmlst.h without ever materialising K copies of the read bases on the host.  The generic
route (BAM -> C++ unpacker, or AlnTable -> packing.pack_table) is validated against this one in
tests/test_devpack.py; here the CIGAR projection is specialised to the generator's four read shapes
(LM, 5S(L-10)M5S, aM1IbM, aM1DbM).  Nothing in this file is on a timed path.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import native, packing
from .streams import DeviceStreams  # noqa: F401  (re-exported: the class lives with the product code)


def _pack_words(bits: torch.Tensor) -> torch.Tensor:
    """bool [n, 32*k] -> int32 [n, k] (bit i of word j = column 32 j + i)."""
    n, c = bits.shape
    w = (bits.view(n, c // 32, 32).to(torch.int64) << torch.arange(32, device=bits.device, dtype=torch.int64)).sum(-1)
    return w.to(torch.int32)  # wraps: same bit pattern as uint32


def read_planes(core: dict, minqual: int = 20):
    """Per READ: (reflen int64 [n], nw int64 [n], rw int64 [n], rows int32 [n, RWmax]) -- 3 interleaved planes aligned to
    the contig's 32-column words (bit i of word j = contig column 32 ((start >> 5) + j) + i), odd-padded."""
    L = core["L"]
    bases, qual, rtype, a, start = core["bases"], core["qual"], core["rtype"], core["a_split"], core["start"]
    dev = bases.device
    n = bases.shape[0]
    reflen = torch.full((n,), L, dtype=torch.int64, device=dev)
    reflen[rtype == 1] = L - 10
    reflen[rtype == 2] = L - 1
    reflen[rtype == 3] = L + 1
    shift = (start.to(torch.int64) & 31)[:, None]
    nwmax = (31 + L + 1 + 31) // 32
    c = torch.arange(nwmax * 32, device=dev)[None, :]
    r = c - shift  # reference offset of row column c
    t, a_ = rtype[:, None], a[:, None]
    q = r.clone()
    q = torch.where(t == 1, r + 5, q)
    q = torch.where(t == 2, torch.where(r < a_, r, r + 1), q)
    q = torch.where(t == 3, torch.where(r < a_, r, torch.where(r == a_, torch.full_like(q, -1), r - 1)), q)
    valid = (r >= 0) & (r < reflen[:, None]) & (q >= 0)
    qc = q.clamp(0, L - 1)
    b = torch.gather(bases, 1, qc)
    ql = torch.gather(qual, 1, qc)
    code = torch.full_like(b, 255)
    for i, ch in enumerate(b"ACGT"):
        code[b == ch] = i
    qok = valid & (ql >= minqual)
    V = qok & (code != 255)
    Nn = qok & (code == 255)
    B1 = V & ((code & 2) != 0)
    B0 = (V & ((code & 1) != 0)) | Nn
    planes = torch.stack([_pack_words(V), _pack_words(B1), _pack_words(B0)], dim=2).reshape(n, 3 * nwmax)
    nw = (shift[:, 0] + reflen + 31) // 32
    rw = 3 * nw
    rw = rw + ((rw & 1) == 0).long()
    rwmax = 3 * nwmax + (1 if (3 * nwmax) % 2 == 0 else 0)
    rows = torch.zeros((n, rwmax), dtype=torch.int32, device=dev)
    rows[:, :3 * nwmax] = planes
    # words beyond 3*nw are zero (columns beyond the span are invalid); the pad word is zero
    return reflen, nw, rw, rows


def pack_cores(db, cores: List[dict], minqual: int = 20, max_depth: Optional[int] = 8000, sentinel_nodes: int = 1,
               want_qhash: bool = False) -> DeviceStreams:
    """Streams of a coordinate-sorted ("--presorted") BAM holding the records of all `cores` (chunks of one sample)."""
    dev = cores[0]["bases"].device
    K = cores[0]["K"]
    L = cores[0]["L"]
    tids, poss, revs, ASs, xms, reflens, nws, rws, rowss = [], [], [], [], [], [], [], [], []
    for c in cores:
        reflen, nw, rw, rows = read_planes(c, minqual)
        tids.append(c["rows"].reshape(-1))
        poss.append(c["start"][:, None].expand(-1, K).reshape(-1))
        revs.append(((c["flag"] >> 4) & 1).reshape(-1))
        ASs.append(c["AS"].reshape(-1))
        xms.append(c["xm"].reshape(-1))
        reflens.append(reflen)
        nws.append(nw)
        rws.append(rw)
        rowss.append(rows)
    tid = torch.cat(tids); pos = torch.cat(poss); rev = torch.cat(revs); AS = torch.cat(ASs); xm = torch.cat(xms)
    reflen_r = torch.cat(reflens); nw_r = torch.cat(nws); rw_r = torch.cat(rws); rows_r = torch.cat(rowss)
    n = tid.shape[0]
    read_of = torch.arange(n, device=dev) // K
    key = (tid << 33) | ((pos + 1) << 1) | rev
    order = torch.sort(key, stable=True).indices
    s = DeviceStreams()
    s.ref_names = db.ref_names()
    s.ref_lens = db.row_len().astype(np.int32)
    s.minqual = minqual
    s.max_depth = max_depth if max_depth is not None else 0
    s.tid = tid[order].to(torch.int32)
    s.as0 = AS[order].to(torch.int16)
    s.xm3 = xm[order].clamp(0, 255).to(torch.uint8)  # synthetic records always carry XS:i => 4th aux field is XM
    s.qlen = torch.full((n,), L, dtype=torch.int16, device=dev)
    rd = read_of[order]
    if want_qhash:  # 128-bit name key per record [n][2]: the read index itself (QNAME "r<id>": injective, so exact)
        s.qhash = torch.stack([rd, torch.full_like(rd, 0x51ED270B0B1F2A37)], dim=1).contiguous()
    # depth cap on the host (sequential), then compaction on the device
    p_reflen_all = reflen_r[rd]
    admitted = np.ones(n, dtype=np.uint8)
    if max_depth is not None:
        t_h = s.tid.cpu().numpy().view(np.uint32)
        p_h = pos[order].to(torch.int32).cpu().numpy()
        r_h = p_reflen_all.to(torch.int32).cpu().numpy().view(np.uint32)
        native.check(native.lib().mmlst_depth_cap(native.ptr(t_h), native.ptr(p_h), native.ptr(r_h), n, int(max_depth),
                                                   int(sentinel_nodes), native.ptr(admitted)))
    adm = torch.from_numpy(admitted).to(dev).bool()
    s.n_dropped = int(n - int(adm.sum()))
    sel = order[adm]
    rd = read_of[sel]
    rw = rw_r[rd]
    off = torch.zeros(sel.shape[0] + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(rw, 0)
    assert int(off[-1]) + packing.PLANE_SLACK_WORDS < (1 << 32)
    # 16-byte mmlst_prec records as int32 [P, 4]: pos | row_off | reflen + (as_named << 16) | xm_named + (nw << 16)
    s.p_recs = torch.stack([pos[sel], off[:-1], reflen_r[rd] | ((AS[sel] & 0xffff) << 16), xm[sel].clamp(0, 255) | (nw_r[rd] << 16)], dim=1).to(torch.int32).contiguous()
    s.n_prec = int(sel.shape[0])
    rwmax = rows_r.shape[1]
    if bool((rw == rwmax).all()):
        body = rows_r[rd].reshape(-1)
    else:
        g = rows_r[rd]
        body = g[torch.arange(rwmax, device=dev)[None, :] < rw[:, None]]
    s.planes = torch.cat([body, torch.zeros(packing.PLANE_SLACK_WORDS, dtype=torch.int32, device=dev)])
    s.max_row_words = int(rw.max()) if sel.shape[0] else 0
    tsel = tid[sel]
    s.contig_start = torch.searchsorted(tsel.contiguous(), torch.arange(len(s.ref_names) + 1, device=dev)).cpu().numpy().astype(np.uint64)
    s.build_runs()
    return s
