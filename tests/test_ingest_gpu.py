"""BAM ingest ON THE DEVICE (csrc/ingest.cu: hardware DEFLATE of the BGZF blocks, record chain, parse, sort, depth cap, plane rows
-- all in HBM) against the C++ host unpacker (mmlst_bam_unpack, itself checked against the Python oracle's reader and pileup in
test_bam_unpack.py / test_ragged_reads.py): every array of both streams must be identical, for every committed golden BAM, for
BAMs whose BGZF blocks cut records at arbitrary bytes, for ragged reads with every CIGAR operator, with the depth cap active,
and the refusals must be the same refusals.  Then the whole sample driver with ingest="device" against the reference's files."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from metamlst_b200 import bam, native, packing, sample
from oracle import bamio

pytestmark = pytest.mark.gpu
BAMS = sorted(glob.glob(os.path.join(GOLDEN, "*", "sample.bam")))


def same_streams(st, soa):
    h = lambda t: t.cpu().numpy()
    assert list(st.ref_names) == list(soa.ref_names) and np.array_equal(st.ref_lens, soa.ref_lens)
    assert int(st.tid.shape[0]) == soa.n_rec
    assert np.array_equal(h(st.tid).view(np.uint32), soa.tid) and np.array_equal(h(st.as0), soa.as0)
    assert np.array_equal(h(st.xm3), soa.xm3) and np.array_equal(h(st.qlen).view(np.uint16), soa.qlen)
    if soa.orig_idx is None:
        assert st.orig_idx is None
    else:
        assert np.array_equal(h(st.orig_idx).view(np.uint32), soa.orig_idx)
    if soa.qhash is not None:
        assert np.array_equal(h(st.qhash).view(np.uint64), soa.qhash)
    assert st.n_prec == soa.n_prec and st.max_row_words == soa.max_row_words and st.n_dropped == soa.n_dropped_by_cap
    assert np.array_equal(h(st.p_recs).reshape(-1).view(packing.PREC_DTYPE), soa.p_recs)
    assert np.array_equal(h(st.planes).view(np.uint32), soa.planes)
    assert np.array_equal(st.contig_start, soa.contig_start)
    if soa.run_tid is None:
        assert st.run_tid is None
    else:
        assert np.array_equal(h(st.run_tid).view(np.uint32), soa.run_tid) and np.array_equal(h(st.run_start).view(np.uint32), soa.run_start)
        assert np.array_equal(h(st.chunk_run).view(np.uint32), soa.chunk_run)
        assert (st.chunk_qlen is None) == (soa.chunk_qlen is None)
        if soa.chunk_qlen is not None:
            assert np.array_equal(h(st.chunk_qlen).view(np.uint16), soa.chunk_qlen)


@pytest.mark.parametrize("path", BAMS, ids=[os.path.basename(os.path.dirname(p)) for p in BAMS])
@pytest.mark.parametrize("presorted", [False, True])
def test_golden_bams(path, presorted):
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    if presorted and "--presorted" not in man[os.path.basename(os.path.dirname(path))]["args"]:
        pytest.skip("file is not coordinate-sorted")
    soa = bam.unpack_bam(path, presorted=presorted, pinned=False)
    st = bam.ingest_bam(path, 0, presorted=presorted)
    same_streams(st, soa)
    assert st.ingest_stats["inflated_bytes"] > st.ingest_stats["compressed_bytes"] > 0


@pytest.mark.parametrize("sizes", [(977,), (4096, 313, 65280), (131,), (60000, 7)])
def test_records_cut_at_arbitrary_bytes(sizes, tmp_path):
    from test_ingest_core import reblock
    src = os.path.join(GOLDEN, "basic", "sample.bam")
    p = str(tmp_path / "cut.bam")
    open(p, "wb").write(reblock(open(src, "rb").read(), sizes))
    st = bam.ingest_bam(p, 0)
    same_streams(st, bam.unpack_bam(p, pinned=False))
    same_streams(st, bam.unpack_bam(src, pinned=False))


@pytest.mark.parametrize("seed,max_depth,minqual", [(1, 8000, 20), (3, 40, 20), (2, 7, 30), (4, None, 0)])
def test_ragged_reads_every_cigar_op_and_the_depth_cap(tmp_path, seed, max_depth, minqual):
    from test_ragged_reads import _ragged_bam
    p, _recs, _rng = _ragged_bam(tmp_path, seed)
    st = bam.ingest_bam(p, 0, minqual=minqual, max_depth=max_depth)
    same_streams(st, bam.unpack_bam(p, minqual=minqual, max_depth=max_depth, pinned=False))


def test_bytes_in_pinned_memory_and_a_second_stream(tmp_path):
    path = os.path.join(GOLDEN, "basic", "sample.bam")
    data = bam.read_pinned(path)
    assert data.is_pinned()
    s2 = torch.cuda.Stream()
    with torch.cuda.stream(s2):
        st = bam.ingest_bam(data, 0)
    same_streams(st, bam.unpack_bam(path, pinned=False))
    st2 = bam.ingest_bam(np.frombuffer(open(path, "rb").read(), np.uint8), 0, want_qhash=False)
    assert st2.qhash is None and torch.equal(st2.as0, st.as0)


def _mutated(tmp_path, mutate, name="m.bam"):
    h, recs = bamio.read_bam(os.path.join(GOLDEN, "basic", "sample.bam"))
    recs = mutate(list(recs))
    p = str(tmp_path / name)
    bamio.write_bam(p, h.ref_names, h.ref_lens, recs)
    return p


@pytest.mark.parametrize("case,code", [("paired", -6), ("noref", -4), ("no_aux3", -4), ("as_range", -7)])
def test_refusals_are_the_host_unpackers_refusals(tmp_path, case, code):
    def mutate(recs):
        r = recs[17]
        if case == "paired":
            recs[17] = r._replace(flag=r.flag | 0x3)
        elif case == "noref":
            recs[17] = r._replace(tid=-1, pos=-1)
        elif case == "no_aux3":
            recs[17] = r._replace(aux=r.aux[:3])
        elif case == "as_range":
            recs[17] = r._replace(aux=(bamio.int_aux("AS", 70000),) + tuple(r.aux[1:]))
        return recs
    p = _mutated(tmp_path, mutate)
    with pytest.raises(native.MmlstError) as e_host:
        bam.unpack_bam(p, pinned=False)
    with pytest.raises(native.MmlstError) as e_dev:
        bam.ingest_bam(p, 0)
    assert e_host.value.code == e_dev.value.code == code
    assert "record 17" in str(e_dev.value)


def test_not_a_bam(tmp_path):
    p = str(tmp_path / "x.bam")
    open(p, "wb").write(b"not a bam file at all, not even gzip" * 10)
    with pytest.raises(native.MmlstError):
        bam.ingest_bam(p, 0)
    bamio.bgzf_write(p, b"BAX\1" + b"\0" * 100)
    with pytest.raises(native.MmlstError):
        bam.ingest_bam(p, 0)


@pytest.mark.parametrize("name", sorted(os.path.basename(os.path.dirname(p)) for p in BAMS))
def test_sample_typer_with_device_ingest_reproduces_reference_files(name, tmp_path):
    from test_sample_driver import _check_against_golden, _params
    d = os.path.join(GOLDEN, name)
    typer = sample.SampleTyper(os.path.join(d, "db.sqlite"), device=0, ingest="device", **_params(name))
    res = typer.type_bam(os.path.join(d, "sample.bam"), str(tmp_path / "out"), want_stdout=True, timestamp=7)
    typer.close()
    _check_against_golden(name, res, str(tmp_path / "out"))
