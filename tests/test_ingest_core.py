"""The per-record text of the DEVICE ingest (metamlst_b200/csrc/ingest_core.cuh: record chain, plausibility guess, field
extraction, plane rows) compiled for the host by g++ (tests/ingest_emul) and driven like csrc/ingest.cu drives it, against
the C++ unpacker (mmlst_bam_unpack) on the committed golden BAMs and on BAMs whose BGZF blocks cut records at arbitrary
bytes -- before a GPU ever runs it.  The GPU tests (tests/test_ingest_gpu.py) then check the kernels end to end."""
import ctypes as C
import glob
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from metamlst_b200 import bam, native, packing

EMUL = os.path.join(ROOT, "tests", "ingest_emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL, "libingest_emul.so")
    srcs = [os.path.join(EMUL, "ingest_emul.cpp"), os.path.join(ROOT, "metamlst_b200", "csrc", "ingest_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "metamlst_b200", "csrc"), "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emul_chain.restype = C.c_longlong
    lib.emul_parse.restype = C.c_uint64
    lib.emul_pack.restype = C.c_uint64
    return lib


def bgzf_blocks(raw: bytes):
    """[(payload, isize)] of a BGZF file."""
    out, p = [], 0
    while p < len(raw):
        xlen = struct.unpack_from("<H", raw, p + 10)[0]
        bsize = None
        q = p + 12
        while q + 4 <= p + 12 + xlen:
            si1, si2, slen = raw[q], raw[q + 1], struct.unpack_from("<H", raw, q + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", raw, q + 4)[0]
            q += 4 + slen
        total = bsize + 1
        isize = struct.unpack_from("<I", raw, p + total - 4)[0]
        if isize:
            out.append((raw[p + 12 + xlen:p + total - 8], isize))
        p += total
    return out


def reblock(raw: bytes, sizes) -> bytes:
    """The same BAM with its inflated stream cut into BGZF blocks of the given sizes (cycled): records straddle the cuts."""
    u = b"".join(zlib.decompress(pl, -15) for pl, _ in bgzf_blocks(raw))
    out, p, i = [], 0, 0
    while p < len(u):
        n = sizes[i % len(sizes)]
        chunk = u[p:p + n]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        pl = co.compress(chunk) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(pl) + 25) + pl + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        p += n
        i += 1
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def emulate(lib, raw: bytes, minqual=20, max_depth=8000):
    """What csrc/ingest.cu computes, with the kernels' per-record functions run on the host.  Returns a dict of numpy arrays in the
    layout of a SoaHost."""
    blocks = bgzf_blocks(raw)
    u = np.frombuffer(b"".join(zlib.decompress(pl, -15) for pl, _ in blocks) + b"\0" * 8, np.uint8)
    usize = u.size - 8
    uoff = np.zeros(len(blocks) + 1, np.uint64)
    uoff[1:] = np.cumsum([i for _, i in blocks])
    l_text = int(u[4:8].view("<i4")[0])
    p = 8 + l_text
    n_ref = int(u[p:p + 4].view("<i4")[0]); p += 4
    ref_len = np.zeros(max(n_ref, 1), np.uint32)
    names = []
    for i in range(n_ref):
        ln = int(u[p:p + 4].view("<i4")[0]); p += 4
        names.append(bytes(u[p:p + ln - 1]).decode()); p += ln
        ref_len[i] = u[p:p + 4].view("<u4")[0]; p += 4
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    cap = usize // 36 + 1
    roff = np.zeros(cap, np.uint64)
    rep = C.c_uint32(0)
    n = lib.emul_chain(ptr(u), C.c_uint64(usize), ptr(uoff), C.c_uint32(len(blocks)), C.c_uint64(p), C.c_int32(n_ref), ptr(ref_len), ptr(roff),
                       C.c_uint64(cap), C.byref(rep))
    assert n >= 0, "malformed chain"
    roff = roff[:n].copy()
    key = np.zeros(n, np.uint64); reflen = np.zeros(n, np.uint16); as0 = np.zeros(n, np.int16); asn = np.zeros(n, np.int16)
    qlen = np.zeros(n, np.uint16); xm3 = np.zeros(n, np.uint8); xmn = np.zeros(n, np.uint8); bits = np.zeros(n, np.uint8); qh = np.zeros(2 * n, np.uint64)
    e = lib.emul_parse(ptr(u), ptr(roff), C.c_uint64(n), C.c_int32(n_ref), ptr(key), ptr(reflen), ptr(as0), ptr(asn), ptr(qlen), ptr(xm3), ptr(xmn),
                       ptr(bits), ptr(qh))
    assert e == 0, ("refused", e >> 8, e & 0xff)
    order = np.argsort(key, kind="stable")
    sorted_already = bool(np.all(order == np.arange(n)))
    ks = key[order]
    tid = (ks >> np.uint64(33)).astype(np.uint32)
    pos = (((ks >> np.uint64(1)) & np.uint64(0xffffffff)).astype(np.int64) - 1).astype(np.int32)
    s_reflen, s_bits, s_roff = reflen[order], bits[order], roff[order]
    cand = np.nonzero(s_bits & 1)[0]
    adm = np.ones(cand.size, np.uint8)
    if max_depth and cand.size:
        ct, cp, cr = np.ascontiguousarray(tid[cand]), np.ascontiguousarray(pos[cand]), np.ascontiguousarray(s_reflen[cand].astype(np.uint32))
        native.check(native.lib().mmlst_depth_cap(native.ptr(ct), native.ptr(cp), native.ptr(cr), cand.size, max_depth, 1, native.ptr(adm)))
    lst = cand[adm != 0].astype(np.uint32)
    nw = packing.touched_words(pos[lst], s_reflen[lst].astype(np.int64))
    rw = packing.row_words(nw)
    rowoff = np.zeros(lst.size + 1, np.uint64)
    rowoff[1:] = np.cumsum(rw)
    planes = np.zeros(int(rowoff[-1]) + 8, np.uint32)
    e = lib.emul_pack(ptr(u), ptr(np.ascontiguousarray(s_roff)), ptr(lst), C.c_uint64(lst.size), ptr(np.ascontiguousarray(pos)), ptr(np.ascontiguousarray(s_reflen)),
                      ptr(np.ascontiguousarray(s_bits)), minqual, ptr(rowoff), ptr(planes))
    assert e == 0, ("pack refused", e >> 8, e & 0xff)
    planes2 = np.full_like(planes, 0xdeadbeef); planes2[int(rowoff[-1]):] = 0
    lib.emul_pack_columns(ptr(u), ptr(np.ascontiguousarray(s_roff)), ptr(lst), C.c_uint64(lst.size), ptr(np.ascontiguousarray(pos)), ptr(np.ascontiguousarray(s_reflen)),
                          minqual, ptr(rowoff), ptr(planes2))
    assert np.array_equal(planes, planes2), "the column-wise (warp) form of the packer differs from the serial walk"
    return dict(names=names, n=n, repairs=rep.value, sorted=sorted_already, tid=tid, as0=as0[order], xm3=xm3[order], qlen=qlen[order], orig_idx=order.astype(np.uint32),
                qhash=qh.reshape(n, 2)[order], p_pos=pos[lst], p_reflen=s_reflen[lst], p_as=asn[order][lst], p_xm=xmn[order][lst], p_nw=nw, p_row_off=rowoff[:-1],
                planes=planes, contig_start=np.searchsorted(tid[lst], np.arange(n_ref + 1)))


def check(got, soa):
    assert got["names"] == list(soa.ref_names) and got["n"] == soa.n_rec
    for k in ("tid", "as0", "xm3", "qlen"):
        assert np.array_equal(got[k], getattr(soa, k)), k
    if soa.orig_idx is not None:
        assert np.array_equal(got["orig_idx"], soa.orig_idx)
    else:
        assert got["sorted"]
    assert np.array_equal(got["qhash"], soa.qhash)
    assert np.array_equal(got["p_pos"], soa.p_recs["pos"]) and np.array_equal(got["p_reflen"], soa.p_recs["reflen"])
    assert np.array_equal(got["p_as"], soa.p_recs["as_named"]) and np.array_equal(got["p_xm"], soa.p_recs["xm_named"])
    assert np.array_equal(got["p_nw"], soa.p_recs["nw"]) and np.array_equal(got["p_row_off"].astype(np.uint32), soa.p_recs["row_off"])
    assert np.array_equal(got["planes"], soa.planes)
    assert np.array_equal(got["contig_start"].astype(np.uint64), soa.contig_start)


BAMS = sorted(glob.glob(os.path.join(GOLDEN, "*", "sample.bam")))


@pytest.mark.parametrize("path", BAMS, ids=[os.path.basename(os.path.dirname(p)) for p in BAMS])
def test_golden_bams_match_the_host_unpacker(emul, path):
    raw = open(path, "rb").read()
    soa = bam.unpack_bam(path, pinned=False)
    check(emulate(emul, raw), soa)


@pytest.mark.parametrize("sizes", [(977,), (4096, 313, 65280), (131,), (60000, 7)])
def test_records_cut_at_arbitrary_bytes(emul, sizes, tmp_path):
    """BGZF blocks that split records anywhere (inside block_size, inside the name, ...): the boundary guesses must be found or
    repaired, and nothing else may change."""
    src = os.path.join(GOLDEN, "basic", "sample.bam")
    raw = reblock(open(src, "rb").read(), sizes)
    p = str(tmp_path / "cut.bam")
    open(p, "wb").write(raw)
    soa = bam.unpack_bam(p, pinned=False)
    ref = bam.unpack_bam(src, pinned=False)
    assert np.array_equal(soa.planes, ref.planes)  # the host unpacker does not care about the cuts
    got = emulate(emul, raw)
    check(got, soa)


def test_depth_cap_and_minqual_variants(emul):
    src = os.path.join(GOLDEN, "basic", "sample.bam")
    raw = open(src, "rb").read()
    for md, mq in ((40, 20), (3, 30), (0, 0)):
        soa = bam.unpack_bam(src, pinned=False, max_depth=md or None, minqual=mq)
        check(emulate(emul, raw, minqual=mq, max_depth=md), soa)
