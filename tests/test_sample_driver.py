"""Whole-sample / cohort driver (metamlst_b200/sample.py) against the reference's own files.

The golden `.nfo`, `.out` and stdout under tests/golden/<case>/ were written by the UNMODIFIED metamlst.py
(metamlst.py:85-299) run over the shims (oracle/make_golden.py).

* CPU tests: the three compute seams the driver calls (`api.score_soa`, `api.coverage_sums`, `api.build_consensus`) are
  replaced by the oracle on the records of the same BAM, so what is checked is the driver's HOST logic: locus gates,
  dict order, number formatting, file names and append mode, the screen text with its colour escapes.
* GPU tests (`-m gpu`): the same comparison with nothing replaced -- BAM -> C++ unpacker -> libmmlst -> files.
"""
import glob
import json
import os
import sqlite3

import pytest

from metamlst_b200 import api, bam, sample
from oracle import bamio
from oracle import mlst_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
CASES = sorted(k for k in MAN if os.path.exists(os.path.join(GOLDEN, k, "sample.bam")))


def _params(name):
    a = MAN[name]["args"]

    def opt(flag, cast, default):
        return cast(a[a.index(flag) + 1]) if flag in a else default
    return dict(minscore=opt("--minscore", int, 80), max_xM=opt("--max_xM", int, 5), min_read_len=opt("--min_read_len", int, 50),
                min_accuracy=opt("--min_accuracy", float, 0.90), penalty=opt("--penalty", int, 100), nloci=opt("--nloci", int, 100),
                species_filter=opt("--filter", str, None), write_known="-a" in a, log="--log" in a, presorted="--presorted" in a)


def _golden_text(d, fname):
    p = os.path.join(d, fname)
    return open(p, newline="").read() if os.path.exists(p) else ""


def _check_against_golden(name, res, out_dir):
    d = os.path.join(GOLDEN, name)
    nfo = os.path.join(out_dir, "sample.nfo")
    assert (open(nfo, newline="").read() if os.path.exists(nfo) else "") == _golden_text(d, "sample.nfo"), name
    if MAN[name]["args"].count("--log"):
        outs = glob.glob(os.path.join(out_dir, "sample_*.out"))
        assert len(outs) == 1
        got = open(outs[0], newline="").read()
        gold = _golden_text(d, "sample.out")
        # the SAMPLE: line carries the path the script was given; everything else byte for byte
        assert got.split("\r\n", 1)[1] == gold.split("\r\n", 1)[1], name
        assert got.startswith("SAMPLE:\t\t\t\t\t")
    gold_stdout = _golden_text(d, "metamlst.stdout")
    if gold_stdout:
        # line 1 is os.path.realpath(sample name) of the directory the golden run happened in
        assert res.stdout.split("\n", 1)[1] == gold_stdout.split("\n", 1)[1], name
        assert res.stdout.split("\n", 1)[0].startswith(sample.OKBLUE + "Sample file: " + sample.ENDC)


class _NoDevice:
    """Stands where native.Context would: the CPU tests never reach the library."""
    handle = None

    def close(self):
        pass


@pytest.fixture
def oracle_seams(monkeypatch):
    """api.score_soa / coverage_sums / build_consensus answered by the oracle from the BAM the soa came from."""
    state = {}

    def load(path):
        h, recs = bamio.read_bam(path)
        state["h"], state["recs"] = h, recs
        state["sorted"] = sorted(recs, key=lambda r: (r.tid, r.pos, (r.flag >> 4) & 1))

    def score_soa(ctx, soa, index, minscore=80, max_xM=5, min_read_len=50, species_filter=None, penalty=100):
        cel, bank, total, ignored = orc.stage1(state["h"], state["recs"], minscore, max_xM, min_read_len, species_filter, penalty)
        state["bank"] = bank
        return cel, total, ignored, None

    def coverage_sums(ctx, soa, index, *a, **k):
        return {key: sum(v.values()) for key, v in state["bank"].items()}

    def build_consensus(ctx, soa, chromosomeList, filterScore, max_xM, debugMode=False, impl=0):
        cons = orc.build_consensus(state["h"], state["sorted"], dict(chromosomeList), filterScore, max_xM, orc.HTS_MAX_DEPTH_DEFAULT)
        return [api.ConsRecord(c.seq, c.id, c.description) for c in cons]

    monkeypatch.setattr(api, "score_soa", score_soa)
    monkeypatch.setattr(api, "coverage_sums", coverage_sums)
    monkeypatch.setattr(api, "build_consensus", build_consensus)
    return load


@pytest.mark.parametrize("name", CASES)
def test_driver_host_logic_reproduces_reference_files(name, oracle_seams, tmp_path):
    d = os.path.join(GOLDEN, name)
    path = os.path.join(d, "sample.bam")
    oracle_seams(path)
    p = _params(name)
    typer = sample.SampleTyper(os.path.join(d, "db.sqlite"), ctx=_NoDevice(), engine="host", **p)
    soa = bam.unpack_bam(path, presorted=p["presorted"])  # host C++ unpacker: supplies ref_names to the driver
    res = typer.type_unpacked(soa, path, want_stdout=True)
    typer.write(res, path, str(tmp_path / "out"), timestamp=7)
    typer.close()
    assert os.path.basename(glob.glob(str(tmp_path / "out" / "*.out"))[0]) == "sample_7.out" if p["log"] else True
    _check_against_golden(name, res, str(tmp_path / "out"))
    assert res.total_reads >= res.ignored_reads >= 0


def test_nfo_is_appended_one_line_per_typing_run(oracle_seams, tmp_path):
    """metamlst.py:284 opens the .nfo in append mode: typing the same sample twice doubles the file."""
    d = os.path.join(GOLDEN, "basic")
    path = os.path.join(d, "sample.bam")
    oracle_seams(path)
    typer = sample.SampleTyper(os.path.join(d, "db.sqlite"), ctx=_NoDevice(), engine="host", **_params("basic"))
    soa = bam.unpack_bam(path)
    for _ in range(2):
        res = typer.type_unpacked(soa, path)
        typer.write(res, path, str(tmp_path / "o"), timestamp=1)
    gold = _golden_text(d, "sample.nfo")
    assert open(tmp_path / "o" / "sample.nfo", newline="").read() == gold + gold
    assert res.stdout == ""  # quiet unless asked


def test_missing_database_is_refused_like_the_reference(tmp_path):
    with pytest.raises(IOError):  # metamlst.py:72-73
        sample.SampleTyper(str(tmp_path / "nope.sqlite"), ctx=_NoDevice())


def test_broken_database_ends_the_sample(oracle_seams, tmp_path):
    """metamlst.py:188-190: fewer genes in the DB than loci seen in the BAM -> message and exit; nothing is written."""
    d = os.path.join(GOLDEN, "basic")
    path = os.path.join(d, "sample.bam")
    db2 = str(tmp_path / "db.sqlite")
    src = sqlite3.connect(os.path.join(d, "db.sqlite"))
    dst = sqlite3.connect(db2)
    src.backup(dst)
    dst.execute("DELETE FROM genes WHERE geneName = 'adk'")
    dst.commit(); dst.close(); src.close()
    oracle_seams(path)
    typer = sample.SampleTyper(db2, ctx=_NoDevice(), engine="host", **_params("basic"))
    res = typer.type_unpacked(bam.unpack_bam(path), path, want_stdout=True)
    assert res.broken_db and res.nfo_lines == [] and "Database is brokenecoli" in res.stdout.replace(" for", "")


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["device", "host", "onecall"])
@pytest.mark.parametrize("name", CASES)
def test_sample_typer_on_gpu_reproduces_reference_files(name, engine, tmp_path):
    """engine="device": BAM -> C++ unpacker -> DeviceStreams.from_soa -> DevicePipeline (one kernel chain, device-side selection);
    engine="host": the four seams through the host-buffer C-ABI; engine="onecall": mmlst_sample.  All must leave the reference's bytes."""
    d = os.path.join(GOLDEN, name)
    path = os.path.join(d, "sample.bam")
    typer = sample.SampleTyper(os.path.join(d, "db.sqlite"), device=0, engine=engine, **_params(name))
    soa = bam.unpack_bam(path, presorted=_params(name)["presorted"], want_qhash=True)
    res = typer.type_unpacked(soa, path, want_stdout=True)
    typer.write(res, path, str(tmp_path / "out"), timestamp=7)
    _check_against_golden(name, res, str(tmp_path / "out"))
    # the quiet form (no log, no screen text: nothing but the result block leaves the device) writes the same .nfo
    typer.log = False
    res2 = typer.type_unpacked(soa, path)
    typer.close()
    assert res2.nfo_lines == res.nfo_lines and res2.stdout == "" and res2.out_log is None
    assert (res2.total_reads, res2.ignored_reads) == (res.total_reads, res.ignored_reads)


@pytest.mark.gpu
def test_broken_database_on_the_device_engine(tmp_path):
    d = os.path.join(GOLDEN, "basic")
    path = os.path.join(d, "sample.bam")
    db2 = str(tmp_path / "db.sqlite")
    src = sqlite3.connect(os.path.join(d, "db.sqlite"))
    dst = sqlite3.connect(db2)
    src.backup(dst)
    dst.execute("DELETE FROM genes WHERE geneName = 'adk'")
    dst.commit(); dst.close(); src.close()
    typer = sample.SampleTyper(db2, device=0, **_params("basic"))
    res = typer.type_unpacked(bam.unpack_bam(path, want_qhash=True), path, want_stdout=True)
    typer.close()
    assert res.broken_db and res.nfo_lines == [] and "Database is brokenecoli" in res.stdout.replace(" for", "")


@pytest.mark.gpu
def test_cohort_driver_types_every_sample_in_order(tmp_path):
    """type_cohort == `for bam in cohort: metamlst.py bam -o out` (overlapped unpack / GPU / write)."""
    names = [n for n in CASES if MAN[n]["args"] in (["--log"], [])][:3] or CASES[:1]
    # same DB is required by one cohort call: type each case's BAM against its own DB in separate calls
    for n in names:
        d = os.path.join(GOLDEN, n)
        out = str(tmp_path / n)
        res = sample.type_cohort([os.path.join(d, "sample.bam")] * 3, os.path.join(d, "db.sqlite"), out, devices=(0,), **_params(n))
        assert [r.sample for r in res] == ["sample"] * 3
        gold = _golden_text(d, "sample.nfo")
        got = open(os.path.join(out, "sample.nfo"), newline="").read() if os.path.exists(os.path.join(out, "sample.nfo")) else ""
        assert got == gold * 3, n


@pytest.mark.gpu
def test_cohort_pool_one_process_per_gpu(tmp_path):
    """sample.CohortPool: persistent worker processes (spawn), one per GPU, device ingest; the files are the reference's."""
    name = "basic"
    d = os.path.join(GOLDEN, name)
    pool = sample.CohortPool(os.path.join(d, "db.sqlite"), [0], ingest="device", **{k: v for k, v in _params(name).items() if k != "log"})
    try:
        for rep in range(2):
            out = str(tmp_path / ("o%d" % rep))
            res = pool.type([os.path.join(d, "sample.bam")] * 3, out)
            assert [r.sample for r in res] == ["sample"] * 3 and all(r.records > 0 for r in res)
            gold = _golden_text(d, "sample.nfo")
            assert open(os.path.join(out, "sample.nfo"), newline="").read() == gold * 3
    finally:
        pool.close()
