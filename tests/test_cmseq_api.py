"""Seam S2' (metamlst_b200/cmseq_api.py) against what the reference's own cmseq returns (tests/golden/cmseq_api.json.gz, made by
oracle/make_golden_cmseq.py: the unmodified cmseq/cmseq.py over the pysam shim).  CPU: the per-contig count seam is answered
by the C oracle, everything above it is the product's host code.  GPU (`-m gpu`): the real path, BAM -> native unpacker ->
pileup kernel -> the same host code."""
import gzip
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import records_to_table
from metamlst_b200 import cmseq_api
from oracle import bamio, corc

GOLD = json.loads(gzip.open(os.path.join(GOLDEN, "cmseq_api.json.gz")).read())


def _same(a, b, path=""):
    """Exact equality of the JSON-able structures, NaN equal to NaN, ints and floats distinguished only by value."""
    if isinstance(a, dict) and isinstance(b, dict):
        assert list(a.keys()) == list(b.keys()), (path, list(a.keys())[:5], list(b.keys())[:5])
        for k in a:
            _same(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, list) and isinstance(b, list):
        assert len(a) == len(b), (path, len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, path + "[%d]" % i)
    elif isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b):
        pass
    else:
        assert a == b, (path, a, b)


def _plain(x):
    if isinstance(x, dict):
        return {str(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)) or type(x).__name__ in ("dict_values", "dict_keys"):
        return [_plain(v) for v in x]
    if isinstance(x, np.generic):
        return x.item()
    return x


def _call(contig, case):
    kw = dict(case["kwargs"])
    if "BAM_tagFilter" in kw:
        kw["BAM_tagFilter"] = [tuple(e) for e in kw["BAM_tagFilter"]]
    if "consensus_rule" in kw:
        kw["consensus_rule"] = getattr(cmseq_api.BamContig, kw["consensus_rule"])
    if "stats_value" in kw:
        return contig.get_all_base_values(kw.pop("stats_value"), **kw)
    return getattr(contig, case["method"])(**kw)


def _check_scenario(scen, bf):
    g = GOLD[scen]
    assert list(bf.get_contigs()) == [c for c in bf.references if c in set(g["contigs"])]  # header order, cmseq/cmseq.py:76
    for case in g["cases"]:
        got = _plain(_call(bf.get_contig_by_label(case["contig"]), case))
        _same(got, case["result"], "%s:%s:%s:%r" % (scen, case["method"], case["contig"], case["kwargs"]))


class _OracleCounts:
    """Answers cmseq_api._contig_counts from the committed BAM with the C oracle (test side only)."""

    def __init__(self, bam):
        h, recs = bamio.read_bam(bam)
        self.tab = records_to_table(h, recs).sorted_by_coord()

    def __call__(self, ctx, soa, tid, minscore, max_xm):
        counts, _ = corc.contig_counts(self.tab, tid, soa.minqual, minscore, max_xm, 8000)
        return counts


@pytest.mark.parametrize("scen", sorted(GOLD))
def test_cmseq_seam_host_logic_matches_the_reference(scen, monkeypatch):
    bam = os.path.join(GOLDEN, scen, "sample.bam")
    monkeypatch.setattr(cmseq_api, "_contig_counts", _OracleCounts(bam))
    g = GOLD[scen]
    bf = cmseq_api.BamFile(bam, filterInputList=list(g["contigs"]), ctx=object())
    assert len(bf.references) == g["n_references"]
    _check_scenario(scen, bf)
    # the other spellings of filterInputList (cmseq/cmseq.py:60-73) and the read-count / length gates
    assert list(cmseq_api.BamFile(bam, filterInputList=",".join(g["contigs"]), ctx=object()).contigs) == list(bf.contigs)
    assert sorted(cmseq_api.BamFile(bam, minimumReadsAligning=50, minlen=100, ctx=object()).contigs) == g["kept_minreads_50"]
    assert bf.get_contig_by_label("no_such_contig") is None


def test_cmseq_seam_refuses_what_it_does_not_implement(monkeypatch):
    bam = os.path.join(GOLDEN, "basic", "sample.bam")
    monkeypatch.setattr(cmseq_api, "_contig_counts", _OracleCounts(bam))
    bf = cmseq_api.BamFile(bam, ctx=object())
    c = next(bf.get_contigs_obj())
    with pytest.raises(NotImplementedError):
        c.get_base_stats(trimReads=(5, 5))
    with pytest.raises(NotImplementedError):
        c.get_base_stats(BAM_tagFilter=[("NM", "loc_lte", 3)])
    with pytest.raises(ValueError):
        c.get_base_stats(min_read_depth=0)
    c.set_stepper("all")
    with pytest.raises(NotImplementedError):
        c.get_base_stats()
    with pytest.raises(Exception, match="is not accessible"):
        cmseq_api.BamFile(bam + ".missing", ctx=object())


def test_cmseq_seam_accepts_untagged_bams_and_raises_only_under_a_tag_filter(tmp_path, monkeypatch):
    """A BAM without XM:i (BWA-like): reference cmseq works on it as long as no BAM_tagFilter is given and dies with pysam's KeyError otherwise
    (cmseq/cmseq.py:545).  Same here: the unfiltered statistics equal those of the fully tagged file, a tag filter raises KeyError."""
    src = os.path.join(GOLDEN, "basic", "sample.bam")
    h, recs = bamio.read_bam(src)
    lean = str(tmp_path / "lean.bam")
    bamio.write_bam(lean, h.ref_names, h.ref_lens, [r._replace(aux=[a for a in r.aux if a[0] == "AS"]) for r in recs])
    monkeypatch.setattr(cmseq_api, "_contig_counts", _OracleCounts(src))   # the counts do not depend on the tags without a filter
    full, bf = cmseq_api.BamFile(src, ctx=object()), cmseq_api.BamFile(lean, ctx=object())
    assert list(bf.contigs) == list(full.contigs)
    name = next(iter(bf.contigs))
    assert _plain(bf.get_contig_by_label(name).get_base_stats()) == _plain(full.get_contig_by_label(name).get_base_stats())
    assert bf.get_contig_by_label(name).reference_free_consensus() == full.get_contig_by_label(name).reference_free_consensus()
    with pytest.raises(KeyError, match="cmseq/cmseq.py:545"):
        bf.get_contig_by_label(name).get_base_stats(BAM_tagFilter=[("AS", "loc_gte", 50), ("XM", "loc_lte", 5)])
    full.get_contig_by_label(name).get_base_stats(BAM_tagFilter=[("AS", "loc_gte", 50), ("XM", "loc_lte", 5)])   # tagged file: fine


@pytest.mark.gpu
@pytest.mark.parametrize("scen", sorted(GOLD))
def test_cmseq_seam_on_the_gpu_matches_the_reference(scen):
    bam = os.path.join(GOLDEN, scen, "sample.bam")
    bf = cmseq_api.BamFile(bam, filterInputList=list(GOLD[scen]["contigs"]))
    try:
        _check_scenario(scen, bf)
    finally:
        bf.close()
