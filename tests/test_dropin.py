"""metamlst_b200/dropin.py -- the reference's own names and signatures -- on the committed golden BAMs, against the
files the UNMODIFIED reference wrote for them (tests/golden, oracle/make_golden.py): stage 1 through score_bam, stage 2 through
buildConsensus(bamFile, chromosomeList, filterScore, max_xM, debugMode), the lines of metamlst.py in between restated by the oracle."""
import inspect
import json
import os
import sqlite3

import pytest

from conftest import GOLDEN
from metamlst_b200 import dropin
from oracle import mlst_oracle as orc

MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
CASES = sorted(k for k in MAN if os.path.exists(os.path.join(GOLDEN, k, "sample.bam")))


def test_signatures_are_the_references():
    assert list(inspect.signature(dropin.buildConsensus).parameters) == ["bamFile", "chromosomeList", "filterScore", "max_xM", "debugMode"]  # metaMLST_functions.py:249
    assert list(inspect.signature(dropin.sort_index).parameters) == ["bamFile"]                                                            # :237
    assert list(inspect.signature(dropin.stringDiff).parameters) == ["s1", "s2"]                                                           # :230
    assert list(inspect.signature(dropin.defineProfile).parameters) == ["conn", "geneList"]                                                # :205
    assert list(inspect.signature(dropin.score_bam).parameters)[:5] == ["bam_path", "minscore", "max_xM", "min_read_len", "species_filter"]  # SURVEY 8b, seam S1


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_metamlst_flow_through_the_rebound_names(name):
    from test_sample_driver import _params
    d = os.path.join(GOLDEN, name)
    p = _params(name)
    path = os.path.join(d, "sample.bam")
    cel, bank, total, ignored = dropin.score_bam(path, p["minscore"], p["max_xM"], p["min_read_len"], p["species_filter"], p["penalty"], p["presorted"])
    gold_out = os.path.join(d, "sample.out")
    if os.path.exists(gold_out):   # the --log rows are cel, row by row, with the reference's own formatting
        rows = open(gold_out, newline="").read().split("RESULTS ------------------------------\r\n", 1)[1]
        assert "".join(orc.out_log_rows(cel)) == rows
        head = open(gold_out, newline="").read()
        assert "TOTAL ALIGNED READS:\t\t\t\t%d\r\n" % total in head and " - OF WHICH IGNORED:\t\t\t\t%d BAM READS" % ignored in head
    db = orc.OracleDB(os.path.join(d, "db.sqlite"))
    lines = []
    for speciesKey, species in cel.items():   # metamlst.py:181-287 with the two seams rebound
        genes = db.genes(speciesKey)
        if len(genes) < len(species):
            break
        if int((float(len(species)) / float(len(set(genes) | set(species)))) * 100) < p["nloci"]:
            continue
        chrom = dict((speciesKey + "_" + g + "_" + a, db.unal_sequence(speciesKey, g, a)) for g, a in orc.select_alleles(species))
        for k in species:
            assert speciesKey + "_" + k in bank   # the coverage column has every detected locus (metamlst.py:228)
        consenSeq = dropin.buildConsensus(path, chrom, p["minscore"], p["max_xM"], False)
        assert [r.id for r in consenSeq] == list(chrom)
        line, _rows = orc.nfo_line(speciesKey, "sample", [orc.ConsRecord(r.seq, r.id, r.description) for r in consenSeq], p["min_accuracy"], p["write_known"],
                                   db.sequence_find)
        if line:
            lines.append(line)
    gold = os.path.join(d, "sample.nfo")
    assert "".join(lines) == (open(gold, newline="").read() if os.path.exists(gold) else "")


@pytest.mark.gpu
def test_merge_names():
    d = os.path.join(GOLDEN, "basic")
    conn = sqlite3.connect(os.path.join(d, "db.sqlite"))
    conn.row_factory = sqlite3.Row
    row = conn.execute("SELECT bacterium, gene, alleleVariant, sequence FROM alleles LIMIT 1 OFFSET 3").fetchone()
    seq = row["sequence"]
    mutated = seq[:10] + ("A" if seq[10] != "A" else "C") + seq[11:]
    dist, allele = dropin.closest_allele(conn, row["bacterium"], row["gene"], mutated)
    want = min((orc.string_diff(mutated, r["sequence"]), i) for i, r in enumerate(conn.execute(
        "SELECT sequence FROM alleles WHERE bacterium = ? AND gene = ? ORDER BY recID", (row["bacterium"], row["gene"]))))
    assert dist == want[0] <= 1
    assert dropin.stringDiff("ACGTAC", "ACGA") == 1   # zip truncation (H9)
    labels = ["%s_%s_%s" % (r[0], r[1], r[2]) for r in conn.execute("SELECT bacterium, gene, alleleVariant FROM alleles GROUP BY gene")]
    from metamlst_b200 import api
    assert dropin.defineProfile(conn, labels) == api.define_profile(conn, labels)
