"""The C restatement (oracle/c) against the Python oracle (itself pinned to the reference by test_oracle_golden)."""
import numpy as np
import pytest

from helpers import lut_from_db, small_case, table_to_records
from oracle import corc, mlst_oracle as orc


@pytest.mark.parametrize("order", ["name", "coord"])
def test_score_matches_python_oracle(order):
    db, tab = small_case(seed=3, n_reads=500, sub_err=0.02)
    if order == "coord":
        tab = tab.sorted_by_coord()
    tab.has_xs = (np.arange(tab.n) % 5) != 0  # H4: 4th aux field becomes XO
    h, recs = table_to_records(tab)
    cel, bank, total, ignored = orc.stage1(h, recs, minscore=170, max_xM=4, min_read_len=50)
    allow, locus_of, n_loci = lut_from_db(db)
    s, c, first, counters = corc.score(tab, allow, locus_of, n_loci, 170, 4, 50)
    assert (int(counters[0]), int(counters[1])) == (total, ignored)
    names = tab.ref_names
    seen = 0
    for sp, genes in cel.items():
        for g, alleles in genes.items():
            maxlen = max(n for (_s, n, _a) in alleles.values())
            for a, (score, n, _avg) in alleles.items():
                t = names.index("%s_%s_%s" % (sp, g, a))
                raw = int(s[t]) - (maxlen - int(c[t])) * 100 if int(c[t]) != maxlen else int(s[t])
                assert (raw, int(c[t])) == (score, n)
                seen += 1
    assert seen == int((c > 0).sum())
    # H5: gene order inside a species == ascending first passing record index
    order_idx = [min(int(first[names.index("ecoli_%s_%s" % (g, a))]) for a in cel["ecoli"][g]) for g in cel["ecoli"]]
    assert order_idx == sorted(order_idx)
    for g, alleles in cel["ecoli"].items():  # allele insertion order too
        ai = [int(first[names.index("ecoli_%s_%s" % (g, a))]) for a in alleles]
        assert ai == sorted(ai)


@pytest.mark.parametrize("maxcnt", [1, 3, 17, 60, 8000])
def test_depth_cap_simulation_matches_engine(maxcnt):
    db, tab = small_case(seed=7, n_reads=900, L=50, K=1, schemes={"ecoli": [("adk", 90), ("fumC", 70)]}, apl=3,
                         frac_clip=0.2, frac_indel=0.2)
    tab = tab.sorted_by_coord()
    h, recs = table_to_records(tab)
    for tid in sorted(set(int(t) for t in tab.tid)):
        contig = [r for r in recs if r.tid == tid]
        eng = orc.PileupEngine(tid, maxcnt)
        for _ in eng.columns(contig):
            pass
        _, admitted = corc.contig_counts(tab, tid, max_depth=maxcnt)
        assert sorted(eng.admitted) == list(np.nonzero(admitted)[0])
        assert len(eng.dropped) == int((admitted == 0).sum())


@pytest.mark.parametrize("maxcnt", [5, 40, 8000])
def test_pileup_counts_and_consensus_match_python_oracle(maxcnt):
    db, tab = small_case(seed=9, n_reads=700, L=60, K=2, schemes={"ecoli": [("adk", 150), ("fumC", 97)]}, apl=3,
                         frac_clip=0.2, frac_indel=0.2, sub_err=0.03, n_frac=0.05)
    tab = tab.sorted_by_coord()
    h, recs = table_to_records(tab)
    tf = [("AS", "loc_gte", 100), ("XM", "loc_lte", 3)]
    for tid in sorted(set(int(t) for t in tab.tid))[:4]:
        contig = [r for r in recs if r.tid == tid]
        stats, _ = orc.get_base_stats(contig, tid, 1, 20, tf, maxcnt)
        counts, _ = corc.contig_counts(tab, tid, 20, 100, 3, maxcnt)
        ln = int(tab.ref_lens[tid])
        for col in range(ln):
            A, Cc, G, T, N = (int(x) for x in counts[col])
            if A + Cc + G + T >= 1:
                f = stats[col + 1]["base_freq"]
                assert (f["A"], f["C"], f["G"], f["T"], f["N"]) == (A, Cc, G, T, N)
            else:
                assert (col + 1) not in stats
        dbseq = db.row_seq(tid)
        cons_py = orc.build_consensus(h, recs, {tab.ref_names[tid]: dbseq}, 100, 3, maxcnt)[0]
        seq, holes, snps = corc.consensus(counts, dbseq.encode(), 1)
        assert (seq, "CI::%d_SP::%d" % (holes, snps)) == (cons_py.seq, cons_py.description)


def test_consensus_tie_break_and_holes():
    counts = np.array([[2, 0, 0, 2, 2], [0, 0, 0, 2, 2], [0, 0, 2, 2, 0], [0, 0, 0, 0, 5], [0, 0, 0, 0, 0], [0, 1, 0, 0, 3]], np.uint32)
    seq, holes, snps = corc.consensus(counts, b"ACGTAC", 1)
    assert seq == "AcGtac" and holes == 4 and snps == 0
    seq, holes, snps = corc.consensus(counts, b"CCTTAC", 1)
    assert seq == "AcGtac" and holes == 4 and snps == 2


def test_hamming_matches_string_diff():
    rng = np.random.default_rng(1)
    rows = ["".join(rng.choice(list("ACGT"), size=int(l))) for l in rng.integers(20, 60, size=40)]
    qs = [rows[3], rows[7][:25] + "A" * 10, "ACGT" * 20]
    off = np.zeros(len(rows) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in rows])
    flat = np.frombuffer("".join(rows).encode(), np.uint8)
    d, a = corc.hamming_min([q.encode() for q in qs], flat, off, [(0, 40), (0, 40), (5, 30)])
    for i, (q, (r0, r1)) in enumerate(zip(qs, [(0, 40), (0, 40), (5, 30)])):
        dd = [orc.string_diff(q, rows[r]) for r in range(r0, r1)]
        assert int(d[i]) == min(dd) and int(a[i]) == r0 + dd.index(min(dd))


@pytest.mark.parametrize("threads,max_depth", [(1, 8000), (3, 8000), (2, 40), (1, None)])
def test_cpu_path_whole_sample_matches_python_oracle(threads, max_depth):
    """oracle/cpu_path.py (what bench.py times as the CPU baseline and checks the GPU arm against at full size): the same
    sample as an AlnTable through the Python oracle's stage 1 / selection / build_consensus."""
    from metamlst_b200 import synth
    from oracle import cpu_path
    db = synth.make_db(("ecoli", "saureus"), alleles_per_locus=6, n_profiles=8, seed=11)
    kw = dict(read_len=100, seed=5, K=3, org_props=(0.6, 0.4))
    tab = synth.make_sample(db, 1500, **kw).sorted_by_coord()
    core = synth.gen_core(db, 1500, **kw)
    w = cpu_path.workload_from_cores(db, [core])
    assert np.array_equal(w.tid, tab.tid) and np.array_equal(w.pos, tab.pos) and np.array_equal(w.aux0, tab.AS)
    got = cpu_path.run(w, minscore=150, max_xM=4, min_read_len=50, penalty=100, max_depth=max_depth, threads=threads)
    h, recs = table_to_records(tab)
    cel, _bank, total, ignored = orc.stage1(h, recs, minscore=150, max_xM=4, min_read_len=50, penalty=100)
    assert (int(got["counters"][0]), int(got["counters"][1])) == (total, ignored)
    assert got["cel"] == cel and [list(v) for v in got["cel"].values()] == [list(v) for v in cel.values()]  # values and dict order
    want = {}
    for sp, genes in cel.items():
        chrom = {"%s_%s_%s" % (sp, g, a): db.row_seq(tab.ref_names.index("%s_%s_%s" % (sp, g, a))) for g, a in orc.select_alleles(genes)}
        for rec in orc.build_consensus(h, recs, chrom, 150, 4, max_depth=max_depth if max_depth else 1 << 30):
            holes, snps = rec.description.split("_")
            want.setdefault(sp, []).append((rec.id, str(rec.seq), int(holes.split("::")[1]), int(snps.split("::")[1])))
    assert got["result"] == want
