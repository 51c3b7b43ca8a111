"""Host packer (CIGAR projection, 3-plane rows, closed-form depth cap) against the oracles -- no GPU needed."""
import numpy as np
import pytest

from helpers import planes_to_counts, small_case
from metamlst_b200 import native, packing
from oracle import corc


def test_library_exports_every_declared_symbol():
    import re, os
    hdr = open(os.path.join(os.path.dirname(native.__file__), "..", "include", "mmlst.h")).read()
    declared = set(re.findall(r"\b(mmlst_[a-z0-9_]+)\s*\(", hdr))
    l = native.lib()
    for name in declared:
        assert hasattr(l, name), name
    assert declared <= set(native.EXPORTS) | {"mmlst_hamming_min_dev2"}
    assert l.mmlst_version() == 200


def test_ctypes_soa_layout_matches_the_header(tmp_path):
    # the ctypes mirror of mmlst_soa (native.Soa) must have the C compiler's layout of include/mmlst.h
    import ctypes as C, os, subprocess
    inc = os.path.join(os.path.dirname(native.__file__), "..", "include")
    fields = [n for n, _t in native.Soa._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mmlst.h"\nint main(void){printf("%zu", sizeof(mmlst_soa));' +
                   "".join('printf(" %%zu", offsetof(mmlst_soa, %s));' % f for f in fields) + 'printf(" %zu %zu", sizeof(mmlst_prec), sizeof(mmlst_chunk));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", inc, str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got[0] == C.sizeof(native.Soa)
    assert got[1:1 + len(fields)] == [getattr(native.Soa, f).offset for f in fields]
    assert got[-2:] == [packing.PREC_DTYPE.itemsize, C.sizeof(native.Chunk)]


def test_ctypes_mirrors_of_the_other_structs_match_the_header(tmp_path):
    """Every struct that crosses the C-ABI by pointer and is mirrored with ctypes: size and the offset of every field as the C compiler lays them out."""
    import ctypes as C, os, subprocess
    from metamlst_b200 import bam
    inc = os.path.join(os.path.dirname(native.__file__), "..", "include")
    pairs = [("mmlst_zstream", native.ZStream), ("mmlst_zpileup", native.ZPileup), ("mmlst_score_params", native.ScoreParams), ("mmlst_index", native.Index),
             ("mmlst_sample_params", native.SampleParams), ("mmlst_sample_result", native.SampleResult), ("mmlst_unpack_opts", bam.UnpackOpts),
             ("mmlst_bam_info_t", bam.BamInfo), ("mmlst_dev_bam_info_t", bam.DevBamInfo)]
    body = ""
    for cname, mirror in pairs:
        body += 'printf(" %%zu", sizeof(%s));' % cname
        for f, _t in mirror._fields_:
            body += 'printf(" %%zu", offsetof(%s, %s));' % (cname, f)
    src = tmp_path / "layout2.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mmlst.h"\nint main(void){' + body + 'return 0;}\n')
    exe = tmp_path / "layout2"
    subprocess.check_call(["gcc", "-I", inc, str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    i = 0
    for cname, mirror in pairs:
        assert got[i] == C.sizeof(mirror), cname
        offs = [getattr(mirror, f).offset for f, _t in mirror._fields_]
        assert got[i + 1:i + 1 + len(offs)] == offs, cname
        i += 1 + len(offs)


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 5000, 70001])
def test_build_runs_is_the_run_length_form_of_tid(n):
    import ctypes as C
    rng = np.random.default_rng(n)
    for style in ("sorted", "single", "alternating"):
        if style == "sorted":
            tid = np.sort(rng.integers(0, 40, n)).astype(np.uint32)
        elif style == "single":
            tid = np.full(n, 7, np.uint32)
        else:
            tid = (np.arange(n) % 3).astype(np.uint32)
        soa = packing.SoaHost([], np.zeros(0, np.int32), tid, np.zeros(n, np.int16), np.zeros(n, np.uint8), np.zeros(n, np.uint16), None,
                              np.zeros(0, packing.PREC_DTYPE), np.zeros(0, np.uint32), 0, np.zeros(1, np.uint64))
        soa.build_runs(max_fraction=1.0)
        if n == 0:
            assert soa.run_tid is None
            continue
        starts = np.concatenate([[0], np.nonzero(tid[1:] != tid[:-1])[0] + 1])
        assert np.array_equal(soa.run_tid, tid[starts])
        assert np.array_equal(soa.run_start, np.concatenate([starts, [n]]))
        assert np.array_equal(np.repeat(soa.run_tid, np.diff(soa.run_start.astype(np.int64))), tid)  # lossless
        first = np.arange(0, n, 256)
        assert np.array_equal(soa.chunk_run, np.searchsorted(soa.run_start, first, side="right") - 1)
        cs = soa.c_struct()
        assert cs.n_runs == len(starts) and cs.run_tid == soa.run_tid.ctypes.data
        # len(SEQ) per chunk: all-zero qlen is uniform; one odd record anywhere drops the form
        assert soa.chunk_qlen is not None and soa.chunk_qlen.shape[0] == (n + 255) // 256 and not soa.chunk_qlen.any()
        assert cs.chunk_qlen == soa.chunk_qlen.ctypes.data
        if n > 1:
            soa.qlen = np.repeat(np.arange((n + 255) // 256, dtype=np.uint16) + 90, 256)[:n].copy()
            soa.build_runs(max_fraction=1.0)
            assert np.array_equal(soa.chunk_qlen, np.arange((n + 255) // 256, dtype=np.uint16) + 90)
            soa.qlen[0] += 1  # chunk 0 holds at least two records
            soa.build_runs(max_fraction=1.0)
            assert soa.chunk_qlen is None and soa.run_tid is not None and not soa.c_struct().chunk_qlen
            soa.qlen[:] = 0
            soa.build_runs(max_fraction=1.0)
        # capacity is checked, not overrun
        nr = C.c_uint32(len(starts) - 1)
        rt = np.zeros(len(starts) + 1, np.uint32); rs = np.zeros(len(starts) + 2, np.uint32); cr = np.zeros(len(first), np.uint32)
        if len(starts) > 1:
            assert native.lib().mmlst_build_runs(native.ptr(tid), n, native.ptr(rt), native.ptr(rs), native.ptr(cr), C.byref(nr)) == -1
    # the threshold keeps the explicit form for run-poor streams
    if n >= 256:
        soa.build_runs(max_fraction=0.125)
        assert soa.run_tid is None


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.MmlstError) as e:
        native.Context(0)
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("maxcnt", [1, 2, 7, 33, 200, 8000])
def test_closed_form_depth_cap_equals_htslib_simulation(maxcnt):
    db, tab = small_case(seed=21, n_reads=1500, L=50, K=2, schemes={"ecoli": [("adk", 90), ("fumC", 70)]}, apl=3,
                         frac_clip=0.2, frac_indel=0.3)
    tab = tab.sorted_by_coord()
    rl = corc.ref_lengths(tab).astype(np.uint32)
    tid = tab.tid.astype(np.uint32)
    pos = tab.pos.astype(np.int32)
    adm = np.zeros(tab.n, np.uint8)
    native.check(native.lib().mmlst_depth_cap(native.ptr(tid), native.ptr(pos), native.ptr(rl), tab.n, maxcnt, 1, native.ptr(adm)))
    for t in sorted(set(int(x) for x in tab.tid)):
        _, sim = corc.contig_counts(tab, t, max_depth=maxcnt)
        assert np.array_equal(adm[tab.tid == t], sim), t


def test_depth_cap_rejects_unsorted():
    db, tab = small_case(seed=22, n_reads=200, L=50, K=1)
    rl = corc.ref_lengths(tab).astype(np.uint32)
    tid = tab.tid.astype(np.uint32); pos = tab.pos.astype(np.int32); adm = np.zeros(tab.n, np.uint8)
    with pytest.raises(native.MmlstError) as e:
        native.check(native.lib().mmlst_depth_cap(native.ptr(tid), native.ptr(pos), native.ptr(rl), tab.n, 8000, 1, native.ptr(adm)))
    assert e.value.code == -5


@pytest.mark.parametrize("order,maxd", [("name", 8000), ("coord", 8000), ("coord", 25), ("name", None)])
def test_packed_planes_reproduce_oracle_counts(order, maxd):
    db, tab = small_case(seed=23, n_reads=500, L=70, K=2, schemes={"ecoli": [("adk", 200), ("fumC", 131)]}, apl=3,
                         frac_clip=0.2, frac_indel=0.3, sub_err=0.03, n_frac=0.05)
    if order == "coord":
        tab = tab.sorted_by_coord()
    soa = packing.pack_table(tab, minqual=20, max_depth=maxd)
    st = tab.sorted_by_coord()
    assert (soa.orig_idx is None) == (order == "coord")
    assert soa.max_row_words % 2 == 1
    for t in sorted(set(int(x) for x in tab.tid)):
        want, adm = corc.contig_counts(st, t, 20, 110, 3, maxd)
        got = planes_to_counts(soa, t, int(tab.ref_lens[t]), 110, 3)
        assert np.array_equal(got, want.astype(np.int64)), t
        assert int(soa.contig_start[t + 1] - soa.contig_start[t]) == int(adm.sum())


def test_packer_refuses_proper_pairs_and_unmapped():
    db, tab = small_case(seed=24, n_reads=50, L=60, K=1)
    tab.flag = tab.flag | 0x3
    with pytest.raises(native.MmlstError) as e:
        packing.pack_table(tab)
    assert e.value.code == -6


def test_hamming_encoding_roundtrip_and_refusal():
    seqs = [b"ACGTACGTAA", b"TTTT", b"G" * 70]
    hi, lo, ln = packing.encode_2bit(seqs, 8)
    assert list(ln) == [10, 4, 70]
    for r, s in enumerate(seqs):
        for i, ch in enumerate(s):
            code = ((int(hi[r, i // 32]) >> (i % 32)) & 1) * 2 + ((int(lo[r, i // 32]) >> (i % 32)) & 1)
            assert b"ACGT"[code] == ch
    th, tl = packing.tile_db(hi, lo)
    assert th.shape[0] == 32 * 8 and int(th[(0 * 8 + 2) * 32 + 2]) == int(hi[2, 2])
    with pytest.raises(native.MmlstError):
        packing.encode_2bit([b"ACGN"], 8)
    # H9: the tolerant form flags the sequence (bit 15 of the length) and keeps the exceptional columns + bytes
    hi, lo, ln, xids, xx, xb = packing.encode_2bit_x([b"ACGT", b"ACNTa" + b"G" * 40, b"TTTT"], 8)
    assert list(ln) == [4, 45 | 0x8000, 4] and list(xids) == [1]
    assert int(xx[0, 0]) == 0b10100 and int(xx[0, 1]) == 0 and bytes(xb[0, :6]) == b"ACNTaG" and xb.shape == (1, 256)
    assert ((int(hi[1, 0]) >> 3) & 1, (int(lo[1, 0]) >> 3) & 1) == (1, 1)  # clean columns keep their code (T)
    assert ((int(hi[1, 0]) >> 2) & 1, (int(lo[1, 0]) >> 2) & 1) == (0, 0)  # exceptional columns are 0 in the planes


def test_deflated_score_stream_inflates_back_to_the_arrays():
    """SoaHost.deflate() (include/mmlst.h, mmlst_zstream): independent raw-DEFLATE blocks + a table ordered by source offset; zlib on the
    host gives back as0[] / xm3[] byte for byte (on the device the hardware decompression engine does: tests/test_gpu_parity.py)."""
    import zlib
    from helpers import small_case
    db, tab = small_case(seed=3, n_reads=3000)
    soa = packing.pack_table(tab.sorted_by_coord(), run_fraction=1.0)
    soa.deflate(block=4096, pinned=False)
    t = soa.z_table
    assert t.shape[1] == 4 and np.all(np.diff(t[:, 2].astype(np.int64)) > 0) and int(t[0, 2]) == 0
    out = {0: bytearray(soa.n_rec * 2), 1: bytearray(soa.n_rec)}
    for kind, dst, src, packed in t.tolist():
        clen, ulen = packed >> 32, packed & 0xffffffff
        raw = zlib.decompress(bytes(soa.z_bytes[src:src + clen]), -15)
        assert len(raw) == ulen <= 4096
        out[kind][dst:dst + ulen] = raw
    # the as0 blocks hold as0 + coeff * xm3 (mmlst_zstream.as_xm_coeff, picked per sample: the synthetic aligner charges 6 per mismatch)
    assert soa.z_as_xm_coeff == 6 and soa.c_struct()._zs.as_xm_coeff == 6
    got_as = np.frombuffer(bytes(out[0]), np.int16).astype(np.int32) - soa.z_as_xm_coeff * np.asarray(soa.xm3).astype(np.int32)
    assert np.array_equal(got_as, np.asarray(soa.as0).astype(np.int32)) and bytes(out[1]) == np.ascontiguousarray(soa.xm3).tobytes()
    assert int(t[-1, 2]) + (int(t[-1, 3]) >> 32) == soa.z_bytes.shape[0]
    smaller = soa.z_bytes.shape[0]
    soa.deflate(block=4096, pinned=False, as_xm_coeff=0)   # off: the blocks are the arrays themselves, and larger
    assert soa.z_as_xm_coeff == 0 and soa.z_bytes.shape[0] > smaller
    raw0 = b"".join(zlib.decompress(bytes(soa.z_bytes[src:src + (packed >> 32)]), -15) for kind, dst, src, packed in soa.z_table.tolist() if kind == 0)
    assert raw0 == np.ascontiguousarray(soa.as0).tobytes()
    # a coefficient that would leave int16 is dropped; only the covered prefix is transformed
    big = packing.pack_table(tab.sorted_by_coord(), run_fraction=1.0)
    big.as0 = np.where(np.arange(big.n_rec) == 5, 32767, np.asarray(big.as0)).astype(np.int16)
    big.xm3 = np.where(np.arange(big.n_rec) == 5, 3, np.asarray(big.xm3)).astype(np.uint8)
    assert big.deflate(block=4096, pinned=False, as_xm_coeff=6).z_as_xm_coeff == 0
    half = packing.pack_table(tab.sorted_by_coord(), run_fraction=1.0).deflate(block=1024, pinned=False, cover=0.5)
    covered = int((half.z_table[half.z_table[:, 0] == 0][:, 3] & np.uint64(0xffffffff)).sum())
    assert half.z_as_xm_coeff == 6 and 0 < covered < 2 * half.n_rec and covered % 1024 == 0
    soa.deflate(block=4096, pinned=False)
    cs = soa.c_struct()
    assert cs.z  # the C struct points at the zstream
    assert packing.pack_table(tab, run_fraction=0.0).deflate().z_bytes is None   # needs the run-length form


def test_deflated_pileup_stream_round_trips_contig_by_contig():
    """SoaHost.deflate(pileup=True) (include/mmlst.h, mmlst_zpileup): the blocks of every contig, inflated with zlib in table order, give back exactly
    its 16-byte records and its plane rows; contigs without records own no blocks; the blocks lie back to back in table order."""
    import zlib
    from metamlst_b200 import synth
    db = synth.make_db(("ecoli", "saureus"), alleles_per_locus=5, n_profiles=8, seed=21)
    soa = packing.pack_table(synth.make_sample(db, 3000, 100, seed=21, K=3)).build_runs()
    for block in (1 << 16, 777):
        soa.deflate(pinned=False, block=block, pileup=True)
        cs, tab, cb = soa.contig_start, soa.zp_table, soa.zp_contig_block
        assert cb.shape[0] == len(soa.ref_names) + 1 and cb[0] == 0 and int(cb[-1]) == tab.shape[0]
        pos = 0
        for t in range(len(soa.ref_names)):
            r0, r1 = int(cs[t]), int(cs[t + 1])
            out = [b"", b""]
            for b in range(int(cb[t]), int(cb[t + 1])):
                off, w = int(tab[b, 0]), int(tab[b, 1])
                kind, clen, ulen = w >> 63, (w >> 32) & 0x7fffffff, w & 0xffffffff
                assert off == pos and ulen <= block
                pos += clen
                d = zlib.decompress(soa.zp_bytes[off:off + clen].tobytes(), -15)
                assert len(d) == ulen and (kind == 1 or not out[1])   # plane blocks first, then record blocks
                out[kind] += d
            if r1 == r0:
                assert cb[t] == cb[t + 1]
                continue
            w0 = int(soa.p_recs["row_off"][r0])
            w1 = int(soa.p_recs["row_off"][r1 - 1]) + int(packing.row_words(soa.p_recs["nw"][r1 - 1:r1])[0])
            assert out[1] == soa.p_recs[r0:r1].tobytes() and out[0] == soa.planes[w0:w1].tobytes()
        assert soa.c_struct().zp
    soa.deflate(pinned=False)
    assert soa.zp_bytes is None and not soa.c_struct().zp
