"""Rows a10 / a11 on the device (csrc/st_match.cu) against the reference's own SQL statements run by sqlite3 on the same DB:
exact-sequence lookup (sequenceExists / sequenceLocate, metaMLST_functions.py:168-172,218-222) and ST assignment
(defineProfile, :205-216, hazard H11), through the host-buffer C-ABI (mmlst_exact_match, mmlst_st_match)."""
import random
import sqlite3

import numpy as np
import pytest

from metamlst_b200 import api, native, synth


def _db(tmp_path, seed=41):
    db = synth.make_db(("ecoli", "saureus"), alleles_per_locus=40, n_profiles=300, seed=seed)
    path = str(tmp_path / "db.sqlite")
    db.write_sqlite(path)
    conn = sqlite3.connect(path)
    rng = random.Random(seed)
    cur = conn.cursor()
    seqs = [r[0] for r in cur.execute("SELECT sequence FROM alleles WHERE bacterium = 'ecoli'")]
    # hazards: the same sequence under TWO genes (fetchone -> lowest rowid), a row that is a strict prefix / extension of another,
    # IUPAC / lower-case rows (H9/H10), a duplicate of an earlier row inside one gene, an organism-crossing duplicate
    extra = [("recA", "ecoli", 900, seqs[3]), ("adk", "ecoli", 901, seqs[5][:-7]), ("adk", "ecoli", 902, seqs[6] + "ACGT"),
             ("fumC", "ecoli", 903, seqs[50][:100] + "N" + seqs[50][101:]), ("fumC", "ecoli", 904, seqs[51].lower()),
             ("fumC", "ecoli", 905, seqs[52][:30] + "R" + seqs[52][31:]), ("gyrB", "ecoli", 906, seqs[100]), ("arcC", "saureus", 907, seqs[7])]
    cur.executemany("INSERT INTO alleles (gene, bacterium, alleleVariant, sequence) VALUES (?,?,?,?)", extra)
    conn.commit()
    return db, conn, seqs, extra, rng


def _sql_locate(conn, bacterium, seq):
    row = conn.execute("SELECT alleleVariant FROM alleles WHERE sequence = ? AND bacterium = ?", (str(seq), bacterium)).fetchone()
    return None if row is None else str(row[0])


@pytest.mark.gpu
def test_exact_lookup_matches_sql(tmp_path):
    db, conn, seqs, extra, rng = _db(tmp_path)
    ctx = native.Context(0)
    idx = api.HammingIndex.from_sqlite(ctx, conn, "ecoli")
    qs = [seqs[i] for i in rng.sample(range(len(seqs)), 60)] + [e[3] for e in extra]
    qs += [seqs[5], seqs[6], seqs[5][:-7] + "A", seqs[6] + "ACG"]            # prefix / extension neighbours that must NOT match their row
    qs += [seqs[9][:200] + ("A" if seqs[9][200] != "A" else "C") + seqs[9][201:]]  # one substitution: no match
    qs += [seqs[51], seqs[51].lower(), seqs[50], seqs[52][:30] + "Y" + seqs[52][31:], "ACGT" * 300, "A"]
    rng.shuffle(qs)
    rows = idx.exact_first(qs, [idx.organism_range("ecoli")] * len(qs))
    n_hit = 0
    for q, r in zip(qs, rows):
        want = _sql_locate(conn, "ecoli", q)
        got = None if int(r) == api.NO_IDX else str(idx.rows[int(r)][2])
        assert got == want, (q[:20], len(q), got, want)
        n_hit += want is not None
    assert n_hit >= 60
    # a range restricted to one locus answers for that locus only
    blk = idx.block[("ecoli", "recA")]
    r = idx.exact_first([seqs[3]], [blk])
    assert str(idx.rows[int(r[0])][2]) == "900"
    ctx.close()


@pytest.mark.gpu
def test_define_profile_matches_sql(tmp_path):
    db, conn, seqs, extra, rng = _db(tmp_path)
    conn.row_factory = sqlite3.Row
    ctx = native.Context(0)
    pidx = api.ProfileIndex(ctx, conn)
    genes = [g for g, _ in db.loci["ecoli"]]
    prof = db.profiles["ecoli"]
    lists = []
    for _ in range(200):
        st = prof[rng.randrange(prof.shape[0])]
        labels = ["ecoli_%s_%d" % (g, int(v)) for g, v in zip(genes, st)]
        k = rng.random()
        if k < 0.3:
            labels[rng.randrange(7)] = "ecoli_%s_%d" % (genes[rng.randrange(7)], rng.randrange(1, 41))   # near miss: best < 7
        elif k < 0.4:
            labels[rng.randrange(6)] = "ecoli_nogene_1"        # unknown label in the middle: dropped, denominator shrinks (H11)
        elif k < 0.5:
            labels[-1] = "ecoli_adk_100001"                    # unknown LAST label: [(0, 0)] (H11)
        elif k < 0.55:
            labels = labels[:3]                                # partial profile: many STs tie
        elif k < 0.6:
            labels.append(labels[0])                           # a label twice: IN is a set test, len(recs) counts it twice
        lists.append(labels)
    lists.append(["saureus_%s_%d" % (g, int(v)) for (g, _l), v in zip(db.loci["saureus"], db.profiles["saureus"][0])])
    lists.append(["ecoli_adk_%d" % int(prof[0][0])])   # one locus only: every ST carrying that allele ties at 100 %
    got = pidx.define_profiles(lists)
    for labels, g in zip(lists, got):
        assert g == api.define_profile(conn, labels), labels
    assert any(len(g) > 1 for g in got) and any(g == [(0, 0)] for g in got) and any(g and g[0][1] == 100 for g in got)
    ctx.close()
