"""Pins oracle/mlst_oracle.py ("Leg A") against the outputs of the reference's UNMODIFIED scripts run over
import shims ("Leg B", oracle/make_golden.py; outputs committed under tests/golden/)."""
import json
import os
import re

import pytest

from conftest import GOLDEN
from oracle import bamio, mlst_oracle as orc

MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
SCEN = [k for k in MAN if k != "cohort"]
ANSI = re.compile(r"\x1b\[[0-9;]*m")


def _args(extra):
    kw = dict(minscore=80, max_xM=5, min_read_len=50, min_accuracy=0.90, species_filter=None, write_known=False)
    i = 0
    while i < len(extra):
        a = extra[i]
        if a == "--minscore":
            kw["minscore"] = int(extra[i + 1]); i += 1
        elif a == "--max_xM":
            kw["max_xM"] = int(extra[i + 1]); i += 1
        elif a == "--min_accuracy":
            kw["min_accuracy"] = float(extra[i + 1]); i += 1
        elif a == "--filter":
            kw["species_filter"] = extra[i + 1]; i += 1
        elif a == "-a":
            kw["write_known"] = True
        i += 1
    return kw


@pytest.mark.parametrize("name", SCEN)
def test_type_sample_matches_reference(name):
    d = os.path.join(GOLDEN, name)
    h, recs = bamio.read_bam(os.path.join(d, "sample.bam"))
    db = orc.OracleDB(os.path.join(d, "db.sqlite"))
    res = orc.type_sample(h, recs, db, "sample", **_args(MAN[name]["args"]))
    # .nfo: byte-exact, organism order = dict order (H5)
    nfo_path = os.path.join(d, "sample.nfo")
    gold_nfo = open(nfo_path, newline="").read() if os.path.exists(nfo_path) else ""
    assert "".join(res["nfo"]) == gold_nfo
    # .out log: counters + per-allele rows (ints + rounded float, H6)
    out = open(os.path.join(d, "sample.out"), newline="").read()
    assert "TOTAL ALIGNED READS:\t\t\t\t%d\r\n" % res["total"] in out
    assert " - OF WHICH IGNORED:\t\t\t\t%d BAM READS" % res["ignored"] in out
    rows = out.split("RESULTS ------------------------------\r\n")[1]
    assert "".join(orc.out_log_rows(res["cel"])) == rows
    # stdout tables: coverage (H7), score, hits, allele list; Ns / SNPs / confidence / notes
    text = ANSI.sub("", open(os.path.join(d, "metamlst.stdout")).read())
    for sp, e in res["species"].items():
        for gene, cov, best, hits, close in e.get("loci", []):
            pat = r"^  %s\s+%s\s+%s\s+%d %s\s*$" % (re.escape(gene), re.escape(str(cov)), re.escape(str(best)), hits, re.escape(close))
            assert re.search(pat, text, re.M), (gene, cov, best, hits, close)
        for rid, leng, holes, snps, conf, note in e.get("table", []):
            g, a = rid.split("_")[1], rid.split("_")[2]
            line = "  " + g.ljust(7) + a.ljust(7) + leng.rjust(7) + holes.rjust(7) + str(snps).rjust(7) + conf.rjust(15) + str(note).rjust(10)
            assert line in text.split("\n"), line


def test_deep_fixture_exercises_depth_cap():
    d = os.path.join(GOLDEN, "deep")
    h, recs = bamio.read_bam(os.path.join(d, "sample.bam"))
    tid = recs[0].tid
    contig = [r for r in recs if r.tid == tid]
    eng = orc.PileupEngine(tid, 8000)
    ncol = sum(1 for _ in eng.columns(contig))
    assert ncol == h.ref_lens[tid]
    assert len(eng.dropped) > 0 and len(eng.admitted) + len(eng.dropped) == len(contig)


def test_cohort_merge_matches_reference():
    d = os.path.join(GOLDEN, "cohort")
    db = orc.OracleDB(os.path.join(d, "db.sqlite"))
    cel = orc.parse_nfo_folder(os.path.join(d, "nfo"))
    assert list(cel) == ["ecoli"]
    st = orc.merge_bacterium(db, "ecoli", cel["ecoli"], z=5)
    assert orc.st_table_text(st) == open(os.path.join(d, "nfo", "merged", "ecoli_ST.txt"), newline="").read()
    assert orc.report_text(st) == open(os.path.join(d, "nfo", "merged", "ecoli_report.txt"), newline="").read()
    codes = sorted(v[2] for v in st["encounteredProfiles"].values())
    assert codes == [1, 1, 3]  # accepted new, accepted new, rejected (> z SNPs)


def test_semantics_probes():
    # H6 / H8 / H9 probes from SURVEY.md 8
    assert round(2705 / 20, 1) == 135.2 and round(12345 / 100, 1) == 123.5
    assert orc.majority_rule({"base_freq": {"A": 2, "T": 2, "C": 0, "G": 0, "N": 2}}) == "A"
    assert orc.majority_rule({"base_freq": {"A": 0, "T": 2, "C": 0, "G": 0, "N": 2}}) == "N"
    assert orc.majority_rule({"base_freq": {"A": 0, "T": 2, "C": 0, "G": 2, "N": 0}}) == "G"
    assert orc.string_diff("ACGTAC", "ACGA") == 1
