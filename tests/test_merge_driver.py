"""Cohort merge driver (metamlst_b200/merge.py) against the files the UNMODIFIED metamlst-merge.py
wrote for the golden cohort (tests/golden/cohort, oracle/make_golden.py): merged/ecoli_ST.txt, merged/ecoli_report.txt and
the screen text.  CPU tests answer the closest-allele search from the oracle's per-character loop (host logic only);
the GPU tests use the real batched Hamming search."""
import os
import shutil
import sqlite3

import numpy as np
import pytest

from metamlst_b200 import merge
from oracle import mlst_oracle as orc

COHORT = os.path.join(os.path.dirname(__file__), "golden", "cohort")
DB = os.path.join(COHORT, "db.sqlite")


def _oracle_closest(db_path):
    odb = orc.OracleDB(db_path)

    def closest(bacterium, items):
        return [orc.closest_allele(odb, bacterium, g, s)[0] for g, s in items]
    return closest


def _copy_cohort(tmp_path):
    dst = tmp_path / "nfo"
    dst.mkdir()
    for f in sorted(os.listdir(os.path.join(COHORT, "nfo"))):
        if f.endswith(".nfo"):
            shutil.copy(os.path.join(COHORT, "nfo", f), dst / f)
    return str(dst)


def _gold(name):
    return open(os.path.join(COHORT, "nfo", "merged", name), newline="").read()


def _check(folder, screen):
    assert open(os.path.join(folder, "merged", "ecoli_ST.txt"), newline="").read() == _gold("ecoli_ST.txt")
    assert open(os.path.join(folder, "merged", "ecoli_report.txt"), newline="").read() == _gold("ecoli_report.txt")
    gold = open(os.path.join(COHORT, "merge.stdout"), newline="").read()
    assert screen == gold


def test_parse_folder_matches_oracle():
    a = merge.read_nfo_folder(os.path.join(COHORT, "nfo"))
    b = orc.parse_nfo_folder(os.path.join(COHORT, "nfo"))
    assert a == b and list(a) == list(b)
    assert merge.read_nfo_folder(os.path.join(COHORT, "nfo"), "saureus") == {}


def test_merge_host_logic_reproduces_reference_files(tmp_path):
    folder = _copy_cohort(tmp_path)
    states, screen = merge.merge_folder(folder, DB, closest=_oracle_closest(DB))
    _check(folder, screen)
    st = states["ecoli"]
    want = orc.merge_bacterium(orc.OracleDB(DB), "ecoli", orc.parse_nfo_folder(folder)["ecoli"], 5)
    assert st.isolates == want["isolates"]
    assert {k: v for k, v in st.new_profiles.items()} == want["encounteredProfiles"]
    assert st.old_profiles == want["oldProfiles"]
    assert st.n_searched >= 1


@pytest.mark.parametrize("z", [None, 0, 1, 3, 50])
def test_merge_thresholds_match_oracle(tmp_path, z):
    folder = _copy_cohort(tmp_path)
    states, _ = merge.merge_folder(folder, DB, z=z, closest=_oracle_closest(DB), write=False)
    want = orc.merge_bacterium(orc.OracleDB(DB), "ecoli", orc.parse_nfo_folder(folder)["ecoli"], z)
    st = states["ecoli"]
    assert merge.CohortMerger.st_table(st) == orc.st_table_text(want)
    assert merge.CohortMerger.report(st) == orc.report_text(want)
    assert st.new_profiles == want["encounteredProfiles"]


def test_exact_table_is_sequence_exists_and_locate():
    m = merge.CohortMerger(DB, closest=lambda b, items: [0] * len(items))
    odb = orc.OracleDB(DB)
    t = m.exact_table("ecoli")
    rows = sqlite3.connect(DB).execute("SELECT sequence FROM alleles WHERE bacterium='ecoli'").fetchall()
    for (s,) in rows[::7]:
        assert s in t and t[s] == odb.sequence_locate("ecoli", s)
        assert s.lower() not in t and not odb.sequence_exists("ecoli", s.lower())  # case-sensitive (H10)
    m.close()


def test_metadata_join_in_report(tmp_path):
    folder = _copy_cohort(tmp_path)
    meta = tmp_path / "meta.tsv"
    meta.write_text("sampleID\tdiet\ns0\tomnivore\ns2\tvegan\nbroken line\n")
    states, _ = merge.merge_folder(folder, DB, closest=_oracle_closest(DB), meta_path=str(meta))
    rep = open(os.path.join(folder, "merged", "ecoli_report.txt")).read().splitlines()
    assert rep[0] == "ST\tConfidence\tsampleID\tdiet"
    by_sample = {l.split("\t")[2]: l for l in rep[1:]}
    assert by_sample["s0"].endswith("\ts0\tomnivore") and by_sample["s2"].endswith("\ts2\tvegan")
    assert any(l.split("\t")[2] not in ("s0", "s2") and len(l.split("\t")) == 3 for l in rep[1:])


def test_no_device_no_search():
    m = merge.CohortMerger(DB)  # no ctx, no stub
    with pytest.raises(RuntimeError):
        m.closest_distances("ecoli", [("adk", "ACGT")])
    m.close()
    with pytest.raises(IOError):
        merge.CohortMerger("/nonexistent/db.sqlite")


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_merge_on_gpu_reproduces_reference_files(tmp_path):
    from metamlst_b200 import native
    folder = _copy_cohort(tmp_path)
    ctx = native.Context(0)
    try:
        states, screen = merge.merge_folder(folder, DB, ctx=ctx)
        _check(folder, screen)
        for z in (None, 0, 2, 50):
            st = merge.merge_folder(folder, DB, ctx=ctx, z=z, write=False)[0]["ecoli"]
            want = orc.merge_bacterium(orc.OracleDB(DB), "ecoli", orc.parse_nfo_folder(folder)["ecoli"], z)
            assert merge.CohortMerger.st_table(st) == orc.st_table_text(want), z
    finally:
        ctx.close()


@pytest.mark.gpu
def test_batched_search_equals_per_character_loop_on_mutated_cohort(tmp_path):
    """A larger synthetic cohort: every sample line carries loci mutated 0..12 times (some with N / IUPAC letters)."""
    from metamlst_b200 import native
    rng = np.random.default_rng(77)
    conn = sqlite3.connect(DB)
    rows = conn.execute("SELECT gene, alleleVariant, sequence FROM alleles WHERE bacterium='ecoli' ORDER BY recID").fetchall()
    by_gene = {}
    for g, v, s in rows:
        by_gene.setdefault(g, []).append((v, s))
    folder = tmp_path / "nfo"
    folder.mkdir()
    for smp in range(24):
        items = []
        for g, alle in sorted(by_gene.items()):
            v, s = alle[int(rng.integers(len(alle)))]
            k = int(rng.choice([0, 0, 1, 2, 4, 6, 12]))
            b = list(s)
            for p in rng.choice(len(b), size=k, replace=False):
                b[p] = "ACGTNR"[int(rng.integers(6))] if b[p] != "A" else "C"
            seq = "".join(b)
            items.append("ecoli_%s_%s::%s::100.0::%.2f" % (g, v, "" if k == 0 and rng.random() < 0.5 else seq, 0.0))
        (folder / ("m%02d.nfo" % smp)).write_text("ecoli\tm%02d\t" % smp + "\t".join(items) + "\r\n")
    ctx = native.Context(0)
    try:
        for z in (5, 1):
            st = merge.merge_folder(str(folder), DB, ctx=ctx, z=z, write=False)[0]["ecoli"]
            want = orc.merge_bacterium(orc.OracleDB(DB), "ecoli", orc.parse_nfo_folder(str(folder))["ecoli"], z)
            assert st.new_profiles == want["encounteredProfiles"], z
            assert st.isolates == want["isolates"], z
            assert merge.CohortMerger.st_table(st) == orc.st_table_text(want), z
    finally:
        ctx.close()
