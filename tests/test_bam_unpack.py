"""Native BAM unpacker (csrc/bam_unpack.cpp) vs the numpy packer fed by the oracle's pure-Python BAM reader: the two
share no code, so agreement pins BGZF/BAM decoding, aux handling (H4), `samtools sort` order, the depth cap (H1) and
the bit-plane projection of CIGAR/SEQ/QUAL (H3).  CPU only."""
import glob
import os

import numpy as np
import pytest

from metamlst_b200 import bam, native, packing, synth
from oracle import bamio

import helpers

GOLD = os.path.join(os.path.dirname(__file__), "golden")
STREAM_FIELDS = ("tid", "as0", "xm3", "qlen", "contig_start", "ref_lens", "p_recs", "planes")


def assert_same(want: packing.SoaHost, got: packing.SoaHost):
    for k in STREAM_FIELDS:
        assert np.array_equal(getattr(want, k), getattr(got, k)), k
    assert want.max_row_words == got.max_row_words
    assert want.n_dropped_by_cap == got.n_dropped_by_cap
    assert (want.orig_idx is None) == (got.orig_idx is None)
    if want.orig_idx is not None:
        assert np.array_equal(want.orig_idx, got.orig_idx)
    assert list(want.ref_names) == list(got.ref_names)
    # run-length form of the score stream: the unpacker's arrays are those of mmlst_build_runs on the same tid
    chk = packing.SoaHost([], np.zeros(0, np.int32), got.tid, got.as0, got.xm3, got.qlen, None, np.zeros(0, packing.PREC_DTYPE),
                          np.zeros(0, np.uint32), 0, np.zeros(1, np.uint64)).build_runs(0.125)
    assert (chk.run_tid is None) == (got.run_tid is None)
    if got.run_tid is not None:
        for k in ("run_tid", "run_start", "chunk_run"):
            assert np.array_equal(getattr(chk, k), getattr(got, k)), k
        assert (chk.chunk_qlen is None) == (got.chunk_qlen is None)  # len(SEQ) per chunk (3 B / record form)
        if got.chunk_qlen is not None:
            assert np.array_equal(chk.chunk_qlen, got.chunk_qlen) and np.array_equal(np.repeat(got.chunk_qlen, 256)[: got.n_rec], got.qlen)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*", "sample.bam"))), ids=lambda p: os.path.basename(os.path.dirname(p)))
def test_unpack_equals_numpy_packer_on_golden_bams(path):
    h, recs = bamio.read_bam(path)
    want = packing.pack_table(helpers.records_to_table(h, recs))
    assert_same(want, bam.unpack_bam(path, pinned=False))
    assert_same(want, bam.unpack_bam(path, pinned=False, threads=1))


@pytest.mark.parametrize("maxd", [None, 60, 8000])
@pytest.mark.parametrize("order", ["name", "coord"])
def test_unpack_synthetic_bam(tmp_path, order, maxd):
    db, tab = helpers.small_case(seed=11, n_reads=1500, L=100, K=4, orgs=("ecoli", "saureus"), apl=5)
    if order == "coord":
        tab = tab.sorted_by_coord()
    p = str(tmp_path / "s.bam")
    bamio.write_table_bam(p, tab)
    got = bam.unpack_bam(p, max_depth=maxd, pinned=False, presorted=(order == "coord"))
    assert_same(packing.pack_table(tab, max_depth=maxd), got)
    assert got.qhash is not None and got.qhash.shape[0] == tab.n
    # same QNAME <=> same hash on this sample (K records per read share a name)
    qn = tab.qname_id if got.orig_idx is None else tab.qname_id[got.orig_idx]
    _, inv = np.unique(qn, return_inverse=True)
    assert got.qhash.shape == (tab.n, 2)
    _, inv2 = np.unique(got.qhash, axis=0, return_inverse=True)
    inv2 = inv2.reshape(-1)
    assert len(set(zip(inv.tolist(), inv2.tolist()))) == len(set(inv.tolist()))


def _write(tmp_path, recs, names=("ecoli_adk_1",), lens=(500,), name="x.bam"):
    p = str(tmp_path / name)
    bamio.write_bam(p, list(names), list(lens), recs)
    return p


def _rec(**kw):
    aux = kw.pop("aux", (bamio.int_aux("AS", 180), bamio.int_aux("XS", 100), bamio.int_aux("XN", 0), bamio.int_aux("XM", 1)))
    d = dict(qname="r1", flag=0, tid=0, pos=10, mapq=42, cigar=((0, 20),), seq="ACGTACGTACGTACGTACGT", qual=bytes([30] * 20), aux=tuple(aux))
    d.update(kw)
    return bamio.BamRecord(d["qname"], d["flag"], d["tid"], d["pos"], d["mapq"], d["cigar"], d["seq"], d["qual"], d["aux"])


def test_refusals_name_the_reference_line(tmp_path):
    cases = [
        (_rec(flag=0x2 | 0x1), -6, "proper-pair"),
        (_rec(tid=-1, pos=-1), -4, "metamlst.py:107"),
        (_rec(aux=(bamio.int_aux("AS", 180),)), -4, "metamlst.py:109-110"),
        (_rec(aux=(("YT", "Z", "UU"), bamio.int_aux("XS", 1), bamio.int_aux("XN", 0), bamio.int_aux("XM", 1))), -4, "metamlst.py:109-110"),
        (_rec(aux=(bamio.int_aux("AS", 40000), bamio.int_aux("XS", 1), bamio.int_aux("XN", 0), bamio.int_aux("XM", 1))), -7, "int16"),
        (_rec(aux=(bamio.int_aux("ZZ", 180), bamio.int_aux("XS", 1), bamio.int_aux("XN", 0), bamio.int_aux("XO", 1))), -4, "get_tag"),
        (_rec(qual=bytes([0xFF] * 20)), -4, "query_qualities"),
    ]
    for i, (r, code, needle) in enumerate(cases):
        with pytest.raises(native.MmlstError) as e:
            bam.unpack_bam(_write(tmp_path, [r], name="c%d.bam" % i), pinned=False)
        assert e.value.code == code and needle in str(e.value), (i, str(e.value))


def test_presorted_flag_needs_coordinate_order_and_bad_files_fail(tmp_path):
    p = _write(tmp_path, [_rec(pos=50), _rec(pos=10, qname="r2")])
    with pytest.raises(native.MmlstError) as e:
        bam.unpack_bam(p, pinned=False, presorted=True)
    assert e.value.code == -5
    s = bam.unpack_bam(p, pinned=False)  # sorted internally, file order kept for H5
    assert s.orig_idx.tolist() == [1, 0] and s.p_pos.tolist() == [10, 50]
    raw = open(p, "rb").read()
    open(str(tmp_path / "trunc.bam"), "wb").write(raw[: len(raw) // 2])
    with pytest.raises(native.MmlstError):
        bam.unpack_bam(str(tmp_path / "trunc.bam"), pinned=False)
    open(str(tmp_path / "plain.bam"), "wb").write(b"not a bam file at all")
    with pytest.raises(native.MmlstError) as e:
        bam.unpack_bam(str(tmp_path / "plain.bam"), pinned=False)
    assert e.value.code == -4
    with pytest.raises(native.MmlstError) as e:
        bam.unpack_bam(str(tmp_path / "missing.bam"), pinned=False)
    assert e.value.code == -3


def test_unmapped_flag_scores_but_never_piles_up_and_seq_star_has_len_one(tmp_path):
    recs = [_rec(flag=0x4), _rec(qname="r2", seq="", qual=b"", cigar=()), _rec(qname="r3", pos=12)]
    s = bam.unpack_bam(_write(tmp_path, recs), pinned=False)
    assert s.n_rec == 3 and s.qlen.tolist() == [20, 1, 20]  # SEQ '*' -> len('*') == 1 (metamlst.py:111,115)
    assert s.n_prec == 2  # the 0x4 record is skipped by bam_plp_push; the empty one is admitted with reflen 0
    assert sorted(s.p_reflen.tolist()) == [0, 20]


def test_lenient_tags_keep_what_pysam_keeps_and_count_the_untagged(tmp_path):
    """`cmseq_api` unpacks with lenient_tags -- records MetaMLST itself crashes on (no integer 1st / 4th aux field, no AS:i / XM:i) stay in
    the pileup with exactly the plane rows they have in the fully tagged file; the strict unpacker still refuses them with the reference line named;
    the lenient stream cannot be scored; proper pairs stay refused either way."""
    from metamlst_b200 import api
    db, tab = helpers.small_case(seed=23, n_reads=300, L=100, K=2, orgs=("ecoli",), apl=4)
    recs = list(bamio.table_records(tab.sorted_by_coord()))
    names, lens = list(tab.ref_names), [int(x) for x in tab.ref_lens]
    full = str(tmp_path / "full.bam"); bamio.write_bam(full, names, lens, recs, sort_order="coordinate")
    want = bam.unpack_bam(full, pinned=False, presorted=True)
    # BWA-like: AS:i and NM:i only (2 aux fields, no XM); every 5th record without any aux field at all
    stripped = []
    for i, r in enumerate(recs):
        aux = [] if i % 5 == 0 else [a for a in r.aux if a[0] == "AS"] + [bamio.int_aux("NM", 1)]
        stripped.append(r._replace(aux=aux))
    lean = str(tmp_path / "lean.bam"); bamio.write_bam(lean, names, lens, stripped, sort_order="coordinate")
    with pytest.raises(native.MmlstError, match="metamlst.py:109-110"):
        bam.unpack_bam(lean, pinned=False, presorted=True)
    got = bam.unpack_bam(lean, pinned=False, presorted=True, lenient_tags=True)
    assert got.lenient and got.n_untagged == got.n_prec == want.n_prec > 0
    assert np.array_equal(got.planes, want.planes) and np.array_equal(got.contig_start, want.contig_start)
    for k in ("pos", "row_off", "reflen", "nw"):
        assert np.array_equal(got.p_recs[k], want.p_recs[k]), k
    assert not got.p_recs["as_named"].any() and not got.p_recs["xm_named"].any()
    assert (got.as0 == -32768).all() and (got.xm3 == 255).all() and np.array_equal(got.qlen, want.qlen) and np.array_equal(got.tid, want.tid)
    with pytest.raises(ValueError, match="lenient_tags"):
        api.score_soa_raw(None, got, api.AlleleIndex(names))
    # the fully tagged file through the lenient unpacker: nothing changes, nothing is counted
    same = bam.unpack_bam(full, pinned=False, presorted=True, lenient_tags=True)
    assert same.n_untagged == 0
    assert_same(want, same)
    # positional fields present but the NAMED tags missing (XS-less bowtie2 order is fine; here the 4 fields are NM, MD-like integers)
    named = [r._replace(aux=[bamio.int_aux("NM", 3), bamio.int_aux("X0", 1), bamio.int_aux("X1", 0), bamio.int_aux("XO", 2)]) for r in recs]
    nm = str(tmp_path / "named.bam"); bamio.write_bam(nm, names, lens, named, sort_order="coordinate")
    with pytest.raises(native.MmlstError, match="cmseq/cmseq.py:545"):
        bam.unpack_bam(nm, pinned=False, presorted=True)
    got2 = bam.unpack_bam(nm, pinned=False, presorted=True, lenient_tags=True)
    assert got2.n_untagged == got2.n_prec and (got2.as0 == 3).all() and (got2.xm3 == 2).all() and np.array_equal(got2.planes, want.planes)
    # H2 stays refused
    paired = [r._replace(flag=r.flag | 0x3) for r in recs[:10]]
    pp = str(tmp_path / "paired.bam"); bamio.write_bam(pp, names, lens, paired, sort_order="coordinate")
    for lenient in (False, True):
        with pytest.raises(native.MmlstError, match="proper-pair"):
            bam.unpack_bam(pp, pinned=False, presorted=True, lenient_tags=lenient)
