"""Parity of the CUDA path (through the C-ABI) against the oracles.  Needs a B200: `pytest -m gpu`."""
import ctypes as C
import json
import os
import sqlite3

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import lut_from_db, records_to_table, small_case, table_to_records
from metamlst_b200 import api, native, packing
from oracle import bamio, corc, mlst_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = native.Context(0)
    yield c
    c.close()


# forms of the run-length score kernel (mmlst_set_score_variant) the stage-1 tests run under
SCORE_VARIANTS = tuple(int(x) for x in os.environ.get("MMLST_TEST_SCORE_VARIANTS", "6,5,0,2").split(","))  # 6 = the library default


@pytest.fixture(params=SCORE_VARIANTS, ids=lambda v: "form%d" % v)
def kernel_form(request):
    prev = native.lib().mmlst_set_score_variant(request.param)
    assert native.lib().mmlst_set_score_variant(-1) == request.param
    yield request.param
    native.lib().mmlst_set_score_variant(prev)


# ------------------------------------------------------------------------------------------------ stage 1
@pytest.mark.parametrize("form", ["tid", "runs"])  # explicit allele id per record (9 B) / run-length form (5 B)
@pytest.mark.parametrize("order", ["name", "coord"])
@pytest.mark.parametrize("n_reads,species_filter", [(37, None), (3000, None), (3000, "saureus"), (60000, "ecoli,saureus")])
def test_score_bit_exact(ctx, order, n_reads, species_filter, form, kernel_form):
    db, tab = small_case(seed=31, n_reads=n_reads, orgs=("ecoli", "saureus"), apl=8, sub_err=0.02)
    if order == "coord":
        tab = tab.sorted_by_coord()
    tab.has_xs = (np.arange(tab.n) % 7) != 0  # H4
    soa = packing.pack_table(tab, run_fraction=1.0 if form == "runs" else 0.0)
    assert (soa.run_tid is not None) == (form == "runs")
    index = api.AlleleIndex(tab.ref_names)
    cel, total, ignored, (s, c, f) = api.score_soa(ctx, soa, index, 176, 3, 50, species_filter, 100)
    allow, locus_of, n_loci = lut_from_db(db, species_filter)
    ws, wc, wf, wcnt = corc.score(tab, allow, locus_of, n_loci, 176, 3, 50)
    assert np.array_equal(s, ws) and np.array_equal(c, wc) and np.array_equal(f, wf)
    assert (total, ignored) == (int(wcnt[0]), int(wcnt[1]))
    if n_reads <= 3000:  # the whole dict, including insertion order (H5) and rounded floats (H6)
        h, recs = table_to_records(tab)
        want, _bank, wt, wi = orc.stage1(h, recs, 176, 3, 50, species_filter, 100)
        assert json.dumps(cel) == json.dumps(want)
        assert "".join(orc.out_log_rows(cel)) == "".join(orc.out_log_rows(want))


@pytest.mark.parametrize("n_reads,block", [(40, 1 << 20), (3000, 4096), (60000, 1 << 16), (400000, 1 << 20)])
def test_score_from_the_deflated_stream(ctx, n_reads, block):
    """mmlst_soa.z: as0[] / xm3[] shipped as DEFLATE blocks and inflated by the hardware decompression engine -- same tables, bit for bit, as
    the plain arrays and as the C port of the oracle; a corrupt block table is refused."""
    db, tab = small_case(seed=33, n_reads=n_reads, orgs=("ecoli", "saureus"), apl=8, sub_err=0.02)
    tab = tab.sorted_by_coord()
    soa = packing.pack_table(tab, run_fraction=1.0)
    index = api.AlleleIndex(tab.ref_names)
    plain = api.score_soa_raw(ctx, soa, index, 176, 3, 50)
    soa.deflate(block=block, pinned=(n_reads > 3000))
    assert soa.z_bytes is not None and soa.z_table.shape[0] >= 2 and soa.z_bytes.shape[0] < 3 * soa.n_rec
    keep_as0, keep_xm3 = soa.as0, soa.xm3
    soa.as0, soa.xm3 = np.zeros_like(keep_as0), np.zeros_like(keep_xm3)   # the call must not read the plain arrays
    got = api.score_soa_raw(ctx, soa, index, 176, 3, 50)
    for a, b in zip(plain[:3], got[:3]):
        assert np.array_equal(a, b)
    assert plain[3:] == got[3:]
    allow, locus_of, n_loci = lut_from_db(db)
    ws, wc, wf, wcnt = corc.score(tab, allow, locus_of, n_loci, 176, 3, 50)
    assert np.array_equal(got[0], ws) and np.array_equal(got[1], wc) and np.array_equal(got[2], wf) and got[3:] == (int(wcnt[0]), int(wcnt[1]))
    bad = soa.z_table.copy()
    bad[0, 3] = (bad[0, 3] >> np.uint64(32) << np.uint64(32)) | np.uint64((int(bad[0, 3]) & 0xffffffff) - 1)   # wrong inflated size
    soa.z_table = bad
    with pytest.raises(native.MmlstError):
        api.score_soa_raw(ctx, soa, index, 176, 3, 50)


def _np_score(tid, as0, xm3, qlen, idx, allow, n_ref, minscore, max_xm, min_len):
    al = allow[tid] != 0
    ok = al & (as0 >= minscore) & (qlen >= min_len) & (xm3 <= max_xm)
    s = np.zeros(n_ref, np.int64); c = np.zeros(n_ref, np.int64); f = np.full(n_ref, 0xFFFFFFFF, np.int64)
    np.add.at(s, tid[ok], as0[ok].astype(np.int64)); np.add.at(c, tid[ok], 1); np.minimum.at(f, tid[ok], idx[ok].astype(np.int64))
    return s, c.astype(np.uint32), f.astype(np.uint32), int(al.sum()), int((al & ~ok).sum())


@pytest.mark.parametrize("n", [1, 255, 256, 257, 511, 513, 4096, 100003, 5_000_000])
def test_score_run_length_form_on_raw_streams(ctx, n, kernel_form):
    # run boundaries everywhere: single-record runs, runs ending exactly on 256-record chunk edges, one run spanning many
    # chunks, filtered alleles in the middle, negative scores, with and without a file-order index, tails of every size
    rng = np.random.default_rng(n)
    n_ref = 300
    names = ["o%d_g%d_%d" % (t % 3, t % 7, t) for t in range(n_ref)]
    index = api.AlleleIndex(names)
    for style in ("mixed", "long", "short"):
        if style == "long":
            lens = rng.integers(1, max(2, n // 3), 64)
        elif style == "short":
            lens = rng.integers(1, 4, min(n, 200000))
        else:
            lens = np.concatenate([rng.integers(1, 900, 4000), [256, 256, 512, 1, 1, 255, 257, 768]])
            rng.shuffle(lens)
        lens = lens[np.cumsum(lens) <= n]
        if lens.sum() < n:
            lens = np.concatenate([lens, [n - lens.sum()]])
        rt = rng.integers(0, n_ref, lens.shape[0])
        rt[1:][rt[1:] == rt[:-1]] += 1  # adjacent runs differ
        rt %= n_ref
        tid = np.repeat(rt, lens).astype(np.uint32)
        assert tid.shape[0] == n
        as0 = rng.integers(-50, 301, n).astype(np.int16)
        as0[rng.integers(0, n, max(1, n // 1000))] = 32767  # large scores: the warp sums must not overflow a packed field
        xm3 = rng.integers(0, 8, n).astype(np.uint8)
        qlen_rec = rng.integers(30, 160, n).astype(np.uint16)
        qlen_chunk = np.repeat(rng.integers(30, 160, (n + 255) // 256), 256)[:n].astype(np.uint16)  # one len(SEQ) per chunk: QC form (3 B / record)
        allow = (rng.random(n_ref) < 0.8).astype(np.uint8)
        for oidx, qlen in [(o, q) for o in (None, rng.permutation(n).astype(np.uint32)) for q in (qlen_rec, qlen_chunk)]:
            soa = packing.SoaHost(names, np.full(n_ref, 500, np.int32), tid, as0, xm3, qlen, oidx, np.zeros(0, packing.PREC_DTYPE),
                                  np.zeros(0, np.uint32), 0, np.zeros(n_ref + 1, np.uint64)).build_runs(max_fraction=1.0)
            assert soa.run_tid is not None
            assert (soa.chunk_qlen is not None) == (qlen is qlen_chunk or n == 1)
            sum_as = np.zeros(n_ref, np.int64); n_hit = np.zeros(n_ref, np.uint32); first = np.full(n_ref, 0xFFFFFFFF, np.uint32)
            counters = np.zeros(2, np.uint64)
            cs = soa.c_struct()
            prm = native.ScoreParams(100, 5, 50)
            native.check(native.lib().mmlst_score(ctx.handle, C.byref(cs), native.ptr(allow), native.ptr(index.locus_of), index.n_loci, C.byref(prm),
                                                  native.ptr(sum_as), native.ptr(n_hit), native.ptr(first), native.ptr(counters)))
            idx = np.arange(n) if oidx is None else oidx
            ws, wc, wf, wt, wi = _np_score(tid, as0.astype(np.int64), xm3.astype(int), qlen.astype(int), idx, allow, n_ref, 100, 5, 50)
            assert np.array_equal(sum_as, ws) and np.array_equal(n_hit, wc) and np.array_equal(first, wf), (n, style, oidx is None)
            assert (int(counters[0]), int(counters[1])) == (wt, wi)
            # the explicit-tid kernel on the same stream (no run arrays) agrees
            soa.run_tid = None
            s2 = np.zeros(n_ref, np.int64); c2 = np.zeros(n_ref, np.uint32); f2 = np.full(n_ref, 0xFFFFFFFF, np.uint32); k2 = np.zeros(2, np.uint64)
            cs = soa.c_struct()
            native.check(native.lib().mmlst_score(ctx.handle, C.byref(cs), native.ptr(allow), native.ptr(index.locus_of), index.n_loci, C.byref(prm),
                                                  native.ptr(s2), native.ptr(c2), native.ptr(f2), native.ptr(k2)))
            assert np.array_equal(s2, ws) and np.array_equal(c2, wc) and np.array_equal(f2, wf) and np.array_equal(k2, counters)
            if soa.chunk_qlen is not None:  # the same stream through the 5 B / record run-length kernel (explicit qlen[])
                soa.build_runs(max_fraction=1.0)
                soa.chunk_qlen = None
                s3 = np.zeros(n_ref, np.int64); c3 = np.zeros(n_ref, np.uint32); f3 = np.full(n_ref, 0xFFFFFFFF, np.uint32); k3 = np.zeros(2, np.uint64)
                cs = soa.c_struct()
                native.check(native.lib().mmlst_score(ctx.handle, C.byref(cs), native.ptr(allow), native.ptr(index.locus_of), index.n_loci, C.byref(prm),
                                                      native.ptr(s3), native.ptr(c3), native.ptr(f3), native.ptr(k3)))
                assert np.array_equal(s3, ws) and np.array_equal(c3, wc) and np.array_equal(f3, wf) and np.array_equal(k3, counters)


def test_expand_runs_gives_back_tid(ctx):
    import torch
    rng = np.random.default_rng(9)
    for n in (1, 256, 1000, 300001):
        tid = np.sort(rng.integers(0, 5000, n)).astype(np.uint32)
        soa = packing.SoaHost([], np.zeros(0, np.int32), tid, np.zeros(n, np.int16), np.zeros(n, np.uint8), np.zeros(n, np.uint16), None,
                              np.zeros(0, packing.PREC_DTYPE), np.zeros(0, np.uint32), 0, np.zeros(1, np.uint64)).build_runs(1.0)
        d = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
        rt, rs, cr = d(soa.run_tid), d(soa.run_start), d(soa.chunk_run)
        out = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        native.check(native.lib().mmlst_expand_runs_dev(native.ptr(rt), native.ptr(rs), int(rt.shape[0]), native.ptr(cr), n, native.ptr(out),
                                                        torch.cuda.current_stream().cuda_stream))
        assert np.array_equal(out.cpu().numpy().view(np.uint32), tid)


def test_score_empty_and_thresholds(ctx):
    db, tab = small_case(seed=32, n_reads=300)
    soa = packing.pack_table(tab)
    index = api.AlleleIndex(tab.ref_names)
    cel, total, ignored, _ = api.score_soa(ctx, soa, index, 10000, 5, 50)
    assert cel == {} and total == tab.n and ignored == tab.n
    cel, total, ignored, _ = api.score_soa(ctx, soa, index, 80, 5, 101)  # min_read_len above the read length
    assert cel == {} and ignored == tab.n
    cel, total, ignored, _ = api.score_soa(ctx, soa, index, 80, 5, 50, "nosuch")
    assert cel == {} and total == 0


# ------------------------------------------------------------------------------------------------ coverage column (H7)
def _np_coverage(tid, as0, xm3, qlen, idx, key, allow, locus_of, minscore, max_xm, min_len):
    """numpy statement of metamlst.py:127,228 on raw arrays: last passing record per (locus, name) wins."""
    ok = (allow[tid] != 0) & (as0 >= minscore) & (qlen >= min_len) & (xm3 <= max_xm)
    want = {}
    last = {}
    for i in np.nonzero(ok)[0][np.argsort(idx[ok], kind="stable")]:
        last[(int(locus_of[tid[i]]), int(key[i, 0]), int(key[i, 1]))] = int(qlen[i])
    for (l, _a, _b), v in last.items():
        want[l] = want.get(l, 0) + v
    return want


def test_coverage_matches_oracle_sequence_bank(ctx):
    # the reference's sequenceBank on real records: K alignments of a read share its QNAME (unique per locus: 1 of K counts)
    for order in ("name", "coord"):
        db, tab = small_case(seed=33, n_reads=2500, orgs=("ecoli", "saureus"), apl=6, sub_err=0.02)
        if order == "coord":
            tab = tab.sorted_by_coord()
        soa = packing.pack_table(tab)
        index = api.AlleleIndex(tab.ref_names)
        for filt in (None, "saureus"):
            h, recs = table_to_records(tab)
            _cel, bank, _t, _i = orc.stage1(h, recs, 176, 3, 50, filt, 100)
            want = {k: sum(v.values()) for k, v in bank.items()}
            assert api.coverage_sums(ctx, soa, index, 176, 3, 50, filt) == want
            # resident form: right after the scoring call on the same context
            api.score_soa(ctx, soa, index, 176, 3, 50, filt, 100)
            assert api.coverage_sums(ctx, soa, index, 176, 3, 50, filt, stream_resident=True) == want


def test_coverage_last_record_wins_and_names_shared_across_loci(ctx):
    # raw streams: few distinct names (heavy duplication, the same name on several loci), random lengths, shuffled
    # file order through orig_idx => the LAST record in FILE order decides (dict overwrite, metamlst.py:127)
    rng = np.random.default_rng(5)
    names = ["o_g%d_%d" % (g, a) for g in range(5) for a in range(1, 4)]
    index = api.AlleleIndex(names)
    for n in (1, 31, 1000, 70001):
        tid = np.sort(rng.integers(0, len(names), n)).astype(np.uint32)
        as0 = rng.integers(150, 201, n).astype(np.int16)
        xm3 = rng.integers(0, 6, n).astype(np.uint8)
        qlen = rng.integers(40, 152, n).astype(np.uint16)
        key = np.zeros((n, 2), np.uint64)
        key[:, 0] = rng.integers(0, max(2, n // 5), n).astype(np.uint64)
        key[:, 1] = rng.integers(0, 2, n).astype(np.uint64)  # {0,0} occurs: the empty-slot sentinel must not swallow it
        for oidx in (None, rng.permutation(n).astype(np.uint32)):
            soa = packing.SoaHost(names, np.full(len(names), 500, np.int32), tid, as0, xm3, qlen, oidx,
                                  np.zeros(0, packing.PREC_DTYPE), np.zeros(0, np.uint32), 0, np.zeros(len(names) + 1, np.uint64))
            soa.qhash = key
            idx = np.arange(n) if oidx is None else oidx.astype(np.int64)
            want = _np_coverage(tid, as0.astype(int), xm3.astype(int), qlen.astype(int), idx, key, np.ones(len(names), np.uint8), index.locus_of, 170, 3, 50)
            got = api.coverage_sums(ctx, soa, index, 170, 3, 50)
            assert got == {"o_g%d" % l: v for l, v in want.items()}, n


def test_coverage_from_bam_names(ctx, tmp_path):
    # real QNAME bytes through the unpacker's 128-bit hash, unsorted file (orig_idx path), vs the oracle on the same BAM
    from metamlst_b200 import bam
    for name in ("basic", "two_org", "no_xs"):
        d = os.path.join(GOLDEN, name)
        h, recs = bamio.read_bam(os.path.join(d, "sample.bam"))
        _cel, bank, _t, _i = orc.stage1(h, recs, 80, 5, 50, None, 100)
        soa = bam.unpack_bam(os.path.join(d, "sample.bam"))
        index = api.AlleleIndex(soa.ref_names)
        assert api.coverage_sums(ctx, soa, index, 80, 5, 50) == {k: sum(v.values()) for k, v in bank.items()}, name


# ------------------------------------------------------------------------------------------------ stage 2
CASES = [
    dict(seed=41, n_reads=400, L=100, K=4, maxd=8000),
    dict(seed=42, n_reads=5000, L=150, K=4, maxd=8000),
    dict(seed=43, n_reads=5000, L=70, K=2, maxd=8000, frac_clip=0.3, frac_indel=0.3, n_frac=0.05, sub_err=0.03),
    dict(seed=44, n_reads=40000, L=100, K=1, maxd=300),   # cap active
    dict(seed=45, n_reads=40000, L=150, K=2, maxd=None),  # uncapped, deep
    dict(seed=46, n_reads=3, L=100, K=1, maxd=8000),
    dict(seed=47, n_reads=20000, L=50, K=1, maxd=8000, schemes={"ecoli": [("adk", 80), ("fumC", 70)]}),  # H1 at the real cap
    dict(seed=48, n_reads=2000, L=100, K=4, maxd=8000, orgs=("ecoli", "saureus", "kpneumoniae")),
]


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("case", CASES, ids=[str(c["seed"]) for c in CASES])
def test_pileup_counts_and_consensus_bit_exact(ctx, case, impl):
    kw = dict(case)
    maxd = kw.pop("maxd")
    db, tab = small_case(apl=4, **kw)
    soa = packing.pack_table(tab, 20, maxd)
    st = tab.sorted_by_coord()
    minscore, max_xm = 2 * tab.read_len - 24, 3
    tids = sorted(set(int(t) for t in tab.tid))
    dbs = [db.row_seq(t) for t in tids]
    seqs, holes, snps, counts, col_off = api.pileup_consensus(ctx, soa, tids, dbs, minscore, max_xm, 1, impl, want_counts=True)
    for i, t in enumerate(tids):
        want, _ = corc.contig_counts(st, t, 20, minscore, max_xm, maxd)
        got = counts[col_off[i]:col_off[i + 1]]
        assert np.array_equal(got, want), (t, np.argwhere(got != want)[:5])
        wseq, wh, ws = corc.consensus(want, dbs[i].encode(), 1)
        assert (seqs[i], int(holes[i]), int(snps[i])) == (wseq, wh, ws)


def test_pileup_linearity_shards_sum_to_whole(ctx):
    """counts(A) + counts(B) == counts(A u B): what the multi-GPU allreduce relies on."""
    db, tab = small_case(seed=51, n_reads=20000, L=100, K=1, apl=3)
    st = tab.sorted_by_coord()
    half = st.n // 2
    tids = sorted(set(int(t) for t in tab.tid))
    dbs = [db.row_seq(t) for t in tids]
    tot = None
    for part in (st.take(np.arange(half)), st.take(np.arange(half, st.n))):
        soa = packing.pack_table(part, 20, None)
        _, _, _, counts, _ = api.pileup_consensus(ctx, soa, tids, dbs, 170, 3, 1, 2, want_counts=True)
        tot = counts.astype(np.int64) if tot is None else tot + counts
    soa = packing.pack_table(st, 20, None)
    _, _, _, whole, _ = api.pileup_consensus(ctx, soa, tids, dbs, 170, 3, 1, 2, want_counts=True)
    assert np.array_equal(tot, whole.astype(np.int64))
    assert int(whole.sum()) == int(((st.qual >= 20) & (st.seq != 0)).sum()) - _unaligned_qok(st)


def _unaligned_qok(st):
    """quality-passing bases that are soft-clipped or inserted (not in any column)."""
    n = 0
    for i in range(st.n):
        q = 0
        for c in st.cig_ops[st.cig_off[i]:st.cig_off[i + 1]]:
            op, ln = int(c) & 15, int(c) >> 4
            if op in (1, 4):
                n += int((st.qual[i, q:q + ln] >= 20).sum())
            if op in (0, 1, 4, 7, 8):
                q += ln
    return n


def test_build_consensus_seam_on_golden_bams(ctx):
    """Seam S2 against the reference's own .nfo (committed golden, produced by the unmodified scripts over shims)."""
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    for name in ("basic", "two_org", "no_xs", "lowcov", "deep", "strict"):
        d = os.path.join(GOLDEN, name)
        h, recs = bamio.read_bam(os.path.join(d, "sample.bam"))
        tab = records_to_table(h, recs)
        args = man[name]["args"]
        minscore = int(args[args.index("--minscore") + 1]) if "--minscore" in args else 80
        max_xm = int(args[args.index("--max_xM") + 1]) if "--max_xM" in args else 5
        min_acc = float(args[args.index("--min_accuracy") + 1]) if "--min_accuracy" in args else 0.9
        soa = packing.pack_table(tab)
        index = api.AlleleIndex(tab.ref_names)
        cel, total, ignored, _ = api.score_soa(ctx, soa, index, minscore, max_xm, 50, None, 100)
        odb = orc.OracleDB(os.path.join(d, "db.sqlite"))
        lines = []
        for sp, species in cel.items():
            chosen = dict((sp + "_" + g + "_" + a, odb.unal_sequence(sp, g, a)) for g, a in api.select_alleles(species))
            cons = api.build_consensus(ctx, soa, chosen, minscore, max_xm)
            line, _rows = orc.nfo_line(sp, "sample", cons, min_acc, "-a" in args, odb.sequence_find)
            if line:
                lines.append(line)
        gold = open(os.path.join(d, "sample.nfo"), newline="").read() if os.path.exists(os.path.join(d, "sample.nfo")) else ""
        assert "".join(lines) == gold, name
        out = open(os.path.join(d, "sample.out"), newline="").read()
        assert "".join(orc.out_log_rows(cel)) == out.split("RESULTS ------------------------------\r\n")[1], name


def test_consensus_len_mismatch_raises_like_reference(ctx):
    db, tab = small_case(seed=52, n_reads=200)
    soa = packing.pack_table(tab)
    t = int(tab.tid[0])
    with pytest.raises(IndexError):
        api.build_consensus(ctx, soa, {tab.ref_names[t]: db.row_seq(t)[:-3]}, 80, 5)
    with pytest.raises(AttributeError):
        api.build_consensus(ctx, soa, {"ecoli_nosuch_1": "ACGT"}, 80, 5)


# ------------------------------------------------------------------------------------------------ stage 3
def _rand_rows(rng, n, lo, hi):
    return ["".join(rng.choice(list("ACGT"), size=int(l))) for l in rng.integers(lo, hi + 1, size=n)]


@pytest.mark.parametrize("n_rows,lo,hi", [(1, 30, 30), (33, 400, 540), (700, 470, 520), (3000, 60, 250), (1000, 600, 1000)])
def test_hamming_min_bit_exact(ctx, n_rows, lo, hi):
    rng = np.random.default_rng(n_rows)
    base = _rand_rows(rng, 1, hi, hi)[0]
    rows = []
    for l in rng.integers(lo, hi + 1, size=n_rows):  # near-identical alleles: realistic small distances + ties
        s = list(base[:int(l)])
        for p in rng.integers(0, len(s), size=int(rng.integers(0, 8))):
            s[int(p)] = "ACGT"[int(rng.integers(0, 4))]
        rows.append("".join(s))
    rows[min(5, n_rows - 1)] = rows[0]  # exact duplicate -> tie must resolve to the lowest row
    loci = [("org", "g%d" % (i * 3 // max(n_rows, 1))) for i in range(n_rows)]
    idx = api.HammingIndex(ctx, [(o, g, i + 1, s) for i, ((o, g), s) in enumerate(zip(loci, rows))])
    qs, ranges = [], []
    for k in range(40):
        r = int(rng.integers(0, n_rows))
        s = list(rows[r])
        for p in rng.integers(0, len(s), size=int(rng.integers(0, 6))):
            s[int(p)] = "ACGT"[int(rng.integers(0, 4))]
        if k % 5 == 0:
            s = s[:max(1, len(s) - int(rng.integers(0, 40)))]  # shorter query: zip truncation (H9)
        if k % 7 == 0:
            s = s + list("ACGT" * 5)  # longer than its rows
        qs.append("".join(s)[:idx.W * 32])
        ranges.append(idx.block[("org", loci[r][1])] if k % 2 else (0, n_rows))
    flat = np.frombuffer("".join(r[3] for r in idx.rows).encode(), np.uint8)
    off = np.zeros(n_rows + 1, np.int64)
    off[1:] = np.cumsum([len(r[3]) for r in idx.rows])
    wd, wa = corc.hamming_min([q.encode() for q in qs], flat, off, ranges)
    d, a = idx.search(qs, ranges)
    assert np.array_equal(d, wd) and np.array_equal(a, wa)


@pytest.mark.parametrize("n_rows,frac_rows,frac_q", [(40, 0.2, 0.2), (900, 0.01, 0.1), (900, 0.3, 0.0), (300, 0.0, 0.5), (64, 1.0, 1.0)])
def test_hamming_non_acgt_characters_exact_path(ctx, n_rows, frac_rows, frac_q):
    """H9: stringDiff compares characters -- IUPAC codes, N and lower case in DB rows and/or queries; every pair with a
    flagged side goes through the exact kernel, the rest through the 2-bit kernel, and both merge into one min/argmin."""
    rng = np.random.default_rng(1000 + n_rows)
    base = _rand_rows(rng, 1, 500, 500)[0]
    odd = list("NRYKMSWnacgt-")

    def mutate(s, n_sub, p_odd):
        s = list(s)
        for p in rng.integers(0, len(s), size=n_sub):
            s[int(p)] = "ACGT"[int(rng.integers(0, 4))]
        if rng.random() < p_odd:
            for p in rng.integers(0, len(s), size=int(rng.integers(1, 6))):
                s[int(p)] = odd[int(rng.integers(0, len(odd)))]
        return "".join(s)

    rows = [mutate(base[: int(l)], int(rng.integers(0, 6)), frac_rows) for l in rng.integers(420, 501, size=n_rows)]
    rows[min(7, n_rows - 1)] = rows[2]  # duplicate (possibly flagged) row: the lowest index wins
    loci = [("org", "g%d" % (i * 2 // n_rows)) for i in range(n_rows)]
    idx = api.HammingIndex(ctx, [(o, g, i + 1, s) for i, ((o, g), s) in enumerate(zip(loci, rows))])
    assert (idx.n_flagged_rows > 0) == (frac_rows > 0)
    qs, ranges = [], []
    for k in range(60):
        r = int(rng.integers(0, n_rows))
        q = mutate(rows[r], int(rng.integers(0, 4)), frac_q)
        if k % 6 == 0:
            q = q[: max(1, len(q) - int(rng.integers(0, 60)))]
        if k % 9 == 0:
            q = rows[r]  # identical to a (maybe flagged) row: distance 0 even through odd characters
        qs.append(q)
        ranges.append(idx.block[("org", loci[r][1])] if k % 2 else (0, n_rows))
    flat = np.frombuffer("".join(r[3] for r in idx.rows).encode(), np.uint8)
    off = np.zeros(n_rows + 1, np.int64)
    off[1:] = np.cumsum([len(r[3]) for r in idx.rows])
    wd, wa = corc.hamming_min([q.encode() for q in qs], flat, off, ranges)
    d, a = idx.search(qs, ranges)
    assert np.array_equal(d, wd) and np.array_equal(a, wa)
    for k in range(0, 60, 7):  # and the reference's own per-character loop
        r0, r1 = ranges[k]
        dist = [orc.string_diff(qs[k], idx.rows[r][3]) for r in range(r0, r1)]
        assert (int(d[k]), int(a[k])) == (min(dist), r0 + dist.index(min(dist)))


def test_closest_allele_seam_matches_oracle_on_golden_cohort(ctx):
    d = os.path.join(GOLDEN, "cohort")
    odb = orc.OracleDB(os.path.join(d, "db.sqlite"))
    idx = api.HammingIndex.from_sqlite(ctx, sqlite3.connect(os.path.join(d, "db.sqlite")))
    cel = orc.parse_nfo_folder(os.path.join(d, "nfo"))
    n = 0
    for line, _sample in cel["ecoli"]:
        for label, (seq, _acc, _snp) in line.items():
            if not seq:
                continue
            _o, g, _a = label.split("_")
            assert idx.closest_allele("ecoli", g, seq) == orc.closest_allele(odb, "ecoli", g, seq)
            n += 1
    assert n >= 4
    # the merge classification driven by the GPU search reproduces the reference's ST table byte for byte
    st = orc.merge_bacterium(odb, "ecoli", cel["ecoli"], 5, closest=idx.closest_allele)
    assert orc.st_table_text(st) == open(os.path.join(d, "nfo", "merged", "ecoli_ST.txt"), newline="").read()
    assert api.define_profile(odb.conn, ["ecoli_adk_1", "ecoli_nosuch_1"]) == [(0, 0)]


# ------------------------------------------------------------------------------------------------ device-side selection
def test_device_selection_rounding_and_order_match_python(ctx):
    """H5/H6 on the device: exact Python round(x, 1) as integer tenths, lowest int(allele) among ties, dict order."""
    from metamlst_b200 import pipeline
    rng = np.random.default_rng(7)
    names = ["s%d_g%d_%d" % (sp, g, a) for sp in range(3) for g in range(4) for a in (7, 3, 12, 1, 25, 2)]
    grouped = api.AlleleIndex(names)                                  # allele rows grouped by locus: the kernel runs without a row list
    mixed = api.AlleleIndex([names[i] for i in rng.permutation(len(names))])   # loci interleaved: the row list is used
    assert np.all(np.diff(grouped.locus_of.astype(np.int64)) >= 0) and np.any(np.diff(mixed.locus_of.astype(np.int64)) < 0)
    prev_warp = native.lib().mmlst_set_select_warp_finalize(-1)
    for trial in range(60):
        index = grouped if trial % 2 == 0 else mixed
        native.lib().mmlst_set_select_warp_finalize(1 if trial % 4 < 2 else 0)   # one-warp and CTA-wide finalization, same answers
        n = rng.integers(0, 40, size=len(names)).astype(np.uint32)
        n[rng.random(len(names)) < 0.3] = 0
        # sums chosen so that many quotients land on / next to x.x5 ties
        s = (n.astype(np.int64) * rng.integers(100, 140, size=len(names))) + rng.integers(0, 3, size=len(names)) * (n.astype(np.int64) // 20 + 1)
        s = np.where(rng.random(len(names)) < 0.3, (n.astype(np.int64) * 27 + n.astype(np.int64) // 20), s)  # x.05-type ties
        f = rng.permutation(len(names)).astype(np.uint32)
        f[n == 0] = 0xFFFFFFFF
        nloci = int(rng.choice([100, 75, 50, 0]))
        want_all = api.fast_select(index, s, n, f, 100)
        # --nloci gate (metamlst.py:194-206) applied on the host side of the comparison
        want = [(sp, t) for sp, t in want_all if int((float(len(t)) / 4.0) * 100) >= nloci]
        got, err = pipeline.device_select(index, s, n, f, 100, nloci)
        assert err == 0 and got == want, (trial, got, want)
    native.lib().mmlst_set_select_warp_finalize(prev_warp)
    # hand-made ties: 2705/20 = 135.25 -> 135.2 and 2704/20 = 135.2 tie; lowest allele number (3) must win over 7
    idx2 = api.AlleleIndex(["o_g_7", "o_g_3", "o_g_5"])
    got, _ = pipeline.device_select(idx2, np.array([2705, 2704, 2690], np.int64), np.array([20, 20, 20], np.uint32), np.array([0, 1, 2], np.uint32))
    assert got == [("o", [1])]
    # penalty path: fewer hits than the locus maximum
    got, _ = pipeline.device_select(idx2, np.array([1000, 2704, 100], np.int64), np.array([7, 20, 1], np.uint32), np.array([0, 1, 2], np.uint32))
    assert got == api.fast_select(idx2, np.array([1000, 2704, 100], np.int64), np.array([7, 20, 1], np.uint32), np.array([0, 1, 2], np.uint32), 100)


@pytest.mark.parametrize("case", [dict(seed=81, n_reads=3000, L=100, K=4, orgs=("ecoli", "saureus")),
                                  dict(seed=82, n_reads=30000, L=150, K=4, orgs=("ecoli",)),
                                  dict(seed=83, n_reads=60, L=100, K=2, orgs=("ecoli", "saureus", "kpneumoniae"))])
def test_device_pipeline_equals_host_selected_path_and_oracle(ctx, case, kernel_form):
    from metamlst_b200 import devpack, pipeline, synth
    kw = dict(case)
    orgs = kw.pop("orgs")
    db = synth.make_db(orgs, alleles_per_locus=6, n_profiles=10, seed=kw["seed"])
    gk = dict(read_len=kw["L"], seed=kw["seed"], K=kw["K"], sub_err=0.02)
    core = synth.gen_core(db, kw["n_reads"], device="cuda:0", **gk)
    st = devpack.pack_cores(db, [core], 20, 8000, want_qhash=True)
    tab = synth.make_sample(db, kw["n_reads"], device="cuda:0", **gk)
    index = api.AlleleIndex(db.ref_names())
    # coverage column of the device-resident stream (H7) vs the numpy statement on the same reads
    ms = 2 * kw["L"] - 30
    ok = (tab.AS >= ms) & (np.where(tab.has_xs, tab.XM, tab.XO) <= 4) & (kw["L"] >= 50)
    pairs = np.unique(np.stack([index.locus_of[tab.tid[ok]].astype(np.int64), tab.qname_id[ok]], axis=1), axis=0)
    want_cov = {"%s_%s" % index.locus_names[int(l)]: int(c) * kw["L"] for l, c in zip(*np.unique(pairs[:, 0], return_counts=True))}
    assert pipeline.DevicePipeline(st, index, db.row_seq, minscore=ms, max_xM=4).run_coverage() == want_cov
    for nloci in (100, 50):
        pipe = pipeline.DevicePipeline(st, index, db.row_seq, minscore=2 * kw["L"] - 30, max_xM=4, nloci=nloci)
        a = pipe.step()
        b = pipe.step_host_select()
        if nloci == 100:
            b = {sp: v for sp, v in b.items() if len(v) == 7}
        else:
            b = {sp: v for sp, v in b.items() if int(len(v) / 7.0 * 100) >= nloci}
        assert a == b
        assert a == pipe.step()  # idempotent
        pipe.fused_tail = True    # pileup + consensus as ONE launch (the CTA finishing a locus's last chunk calls its consensus): same results,
        assert a == pipe.step() == pipe.step()
        pipe.capture()
        assert a == pipe.step_graph() == pipe.step_graph()
        pipe.fused_tail = False   # ... in either order of use (the ticket counters are left at zero)
        pipe.graph = None
        assert a == pipe.step()
    # oracle: consensus of every chosen contig
    stab = tab.sorted_by_coord()
    for sp, lst in a.items():
        for contig, seq, holes, snps in lst:
            t = index.name_to_tid[contig]
            want, _ = corc.contig_counts(stab, t, 20, 2 * kw["L"] - 30, 4, 8000)
            assert (seq, holes, snps) == corc.consensus(want, db.row_seq(t).encode(), 1)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 1003, 65536 + 5])
def test_as_untransform_is_exact_for_any_length(n):
    """mmlst_as_untransform_dev (as0[i] -= coeff * xm3[i], in place): eight records per thread plus a scalar tail, against numpy for lengths around the
    vector width, negative and extreme values that stay inside int16 after the subtraction."""
    import torch
    rng = np.random.default_rng(n)
    xm = rng.integers(0, 256, n, dtype=np.uint8)
    for coeff in (6, -3, 0):
        want = rng.integers(-20000, 20000, n).astype(np.int16)
        stored = (want.astype(np.int32) + coeff * xm.astype(np.int32)).astype(np.int16)
        a = torch.from_numpy(stored.copy()).cuda() if n else torch.zeros(0, dtype=torch.int16, device="cuda")
        x = torch.from_numpy(xm.copy()).cuda() if n else torch.zeros(0, dtype=torch.uint8, device="cuda")
        native.check(native.lib().mmlst_as_untransform_dev(native.ptr(a) if n else 0, native.ptr(x) if n else 0, n, coeff, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert np.array_equal(a.cpu().numpy(), want)


def test_sample_lanes_keep_several_samples_in_flight_with_type_soa_results(ctx):
    """api.SampleLanes (one context + host thread per lane, mmlst_sample inside): different samples interleaved over 1, 2 and 3 lanes come back in
    submission order and equal what `type_soa` returns for each of them one at a time (plain and deflated streams, a species filter, tables)."""
    from metamlst_b200 import devpack, synth
    db = synth.make_db(("ecoli", "saureus"), alleles_per_locus=6, n_profiles=10, seed=77)
    index = api.AlleleIndex(db.ref_names())
    soas, st0 = [], None
    for i, n_reads in enumerate((2500, 9000, 400, 6000, 1200)):
        st = devpack.pack_cores(db, [synth.gen_core(db, n_reads, device="cuda:0", read_len=100, seed=300 + i, K=4, sub_err=0.02)], 20, 8000)
        st0 = st0 or st
        soa = st.to_host(pinned=True)
        if i % 2 and soa.run_tid is not None:
            soa.deflate(block=4096, pinned=False, cover=0.7, pileup=(i == 3))
        soas.append(soa)
    sidx = api.SampleIndex(ctx, index, st0.ref_lens, db.row_seq)
    kw = dict(minscore=170, max_xM=4, min_read_len=50, penalty=100, nloci=0)
    want = [api.type_soa(sidx, soa, **kw) for soa in soas]
    assert len({str(w["species"]) for w in want}) > 1   # the samples really differ
    for lanes in (1, 2, 3):
        sl = api.SampleLanes(0, index, st0.ref_lens, db.row_seq, lanes=lanes)
        try:
            for _rep in range(3):
                got = sl.map(soas * 2, **kw)
                assert [g["species"] for g in got] == [w["species"] for w in want] * 2
                assert [(g["totalReads"], g["ignoredReads"], g["tids"]) for g in got] == [(w["totalReads"], w["ignoredReads"], w["tids"]) for w in want] * 2
            f = sl.submit(soas[1], species_filter="ecoli", want_tables=True, **kw).result()
            w = api.type_soa(sidx, soas[1], species_filter="ecoli", want_tables=True, **kw)
            assert f["species"] == w["species"] and all(np.array_equal(a, b) for a, b in zip(f["tables"], w["tables"]))
        finally:
            sl.close()


@pytest.mark.parametrize("case", [dict(seed=91, n_reads=3000, L=100, K=4, orgs=("ecoli", "saureus")),
                                  dict(seed=92, n_reads=40000, L=150, K=4, orgs=("ecoli",)),
                                  dict(seed=93, n_reads=60, L=100, K=2, orgs=("ecoli", "saureus", "kpneumoniae"))])
def test_one_call_sample_over_host_buffers_equals_the_device_pipeline_and_the_two_seams(ctx, case):
    """mmlst_sample (score -> device selection -> pileup of the chosen contigs -> consensus, host buffers in, results out) against the device-resident
    pipeline, against mmlst_score + host selection + mmlst_pileup_consensus, and against the oracle; plain, deflated and partly deflated streams."""
    from metamlst_b200 import devpack, pipeline, synth
    kw = dict(case)
    orgs = kw.pop("orgs")
    db = synth.make_db(orgs, alleles_per_locus=6, n_profiles=10, seed=kw["seed"])
    gk = dict(read_len=kw["L"], seed=kw["seed"], K=kw["K"], sub_err=0.02)
    core = synth.gen_core(db, kw["n_reads"], device="cuda:0", **gk)
    st = devpack.pack_cores(db, [core], 20, 8000)
    index = api.AlleleIndex(db.ref_names())
    ms = 2 * kw["L"] - 30
    soa = st.to_host(pinned=True)
    sidx = api.SampleIndex(ctx, index, st.ref_lens, db.row_seq)
    raw = api.score_soa_raw(ctx, soa, index, ms, 4, 50)
    for nloci in (100, 50):
        want = pipeline.DevicePipeline(st, index, db.row_seq, minscore=ms, max_xM=4, nloci=nloci).step()
        got = api.type_soa(sidx, soa, ms, 4, 50, 100, nloci, want_tables=True)
        assert got["species"] == list(want.items())
        assert (got["totalReads"], got["ignoredReads"]) == raw[3:]
        for a, b in zip(got["tables"], raw[:3]):
            assert np.array_equal(a, b)
        assert api.type_soa(sidx, soa, ms, 4, 50, 100, nloci)["species"] == got["species"]   # idempotent, tables not asked for
    # the two seams over the same buffers (host selection has no --nloci gate: compare at nloci=0)
    chosen = api.fast_select(index, raw[0], raw[1], raw[2], 100)
    ts = [t for _sp, tt in chosen for t in tt]
    seqs, holes, snps, _, _ = api.pileup_consensus(ctx, soa, ts, [db.row_seq(t) for t in ts], ms, 4, 1, 0)
    g0 = api.type_soa(sidx, soa, ms, 4, 50, 100, 0)
    assert g0["tids"] == ts
    assert [x for _sp, lst in g0["species"] for x in lst] == [(index.ref_names[t], seqs[i], int(holes[i]), int(snps[i])) for i, t in enumerate(ts)]
    if soa.run_tid is not None:
        for cover in (1.0, 0.5, 0.0):
            for coeff in (None, 0, 3):   # as0 + coeff * xm3 in the blocks (picked per sample / off / a coefficient that does not fit the data)
                soa.deflate(block=4096, pinned=False, cover=cover, as_xm_coeff=coeff)
                assert soa.z_as_xm_coeff == ((6 if coeff is None else coeff) if cover > 0 and soa.z_bytes is not None and (soa.z_table[:, 0] == 0).any() else 0)
                r = api.type_soa(sidx, soa, ms, 4, 50, 100, 0, want_tables=True)
                assert r["species"] == g0["species"] and all(np.array_equal(a, b) for a, b in zip(r["tables"], raw[:3]))
                if cover == 1.0:   # the two-seam call (mmlst_score) undoes it too
                    assert all(np.array_equal(a, b) for a, b in zip(api.score_soa_raw(ctx, soa, index, ms, 4, 50)[:3], raw[:3]))
        soa.z_bytes = soa.z_table = None
    # the pileup stream as DEFLATE blocks per contig (mmlst_zpileup): only the chosen contigs' blocks cross the bus, the engine writes records and rows
    g100 = api.type_soa(sidx, soa, ms, 4, 50, 100, 100)
    for block in (1 << 16, 1000):
        soa.deflate(block=block, pinned=False, cover=0.5, pileup=True)
        assert soa.zp_bytes is not None and int(soa.zp_contig_block[-1]) == soa.zp_table.shape[0]
        assert api.type_soa(sidx, soa, ms, 4, 50, 100, 0)["species"] == g0["species"]
        assert api.type_soa(sidx, soa, ms, 4, 50, 100, 100)["species"] == g100["species"]
    t_hit = g0["tids"][0]
    b = int(soa.zp_contig_block[t_hit])
    good = int(soa.zp_table[b, 1])
    soa.zp_table[b, 1] = good + 1    # a block that claims one byte more than it inflates to
    with pytest.raises(RuntimeError, match="zp"):
        api.type_soa(sidx, soa, ms, 4, 50, 100, 0)
    soa.zp_table[b, 1] = good
    assert api.type_soa(sidx, soa, ms, 4, 50, 100, 0)["species"] == g0["species"]
    soa.z_bytes = soa.z_table = soa.zp_bytes = soa.zp_table = soa.zp_contig_block = None
    # species filter: what is not allowed never scores, so it is never chosen
    f = api.type_soa(sidx, soa, ms, 4, 50, 100, 0, species_filter=orgs[0])
    assert [sp for sp, _ in f["species"]] == [sp for sp, _ in g0["species"] if sp == orgs[0]]
    # oracle: consensus of every chosen contig
    tab = synth.make_sample(db, kw["n_reads"], device="cuda:0", **gk).sorted_by_coord()
    for sp, lst in g0["species"]:
        for contig, seq, h, s in lst:
            t = index.name_to_tid[contig]
            wc, _ = corc.contig_counts(tab, t, 20, ms, 4, 8000)
            assert (seq, h, s) == corc.consensus(wc, db.row_seq(t).encode(), 1)


def test_one_call_sample_refusals(ctx):
    from metamlst_b200 import devpack, synth
    db = synth.make_db(("ecoli",), alleles_per_locus=4, n_profiles=5, seed=5)
    core = synth.gen_core(db, 2000, device="cuda:0", read_len=100, seed=5, K=2, sub_err=0.01)
    st = devpack.pack_cores(db, [core], 20, 8000)
    index = api.AlleleIndex(db.ref_names())
    soa = st.to_host(pinned=True)
    other = native.Context(0)
    try:   # no index in this context
        bufs = [np.zeros(8192, np.uint32) for _ in range(5)]
        cons = np.zeros(8192, np.uint8)
        res = native.SampleResult()
        res.chosen_tid, res.chosen_species, res.col_off, res.holes, res.snps = (native.ptr(b) for b in bufs)
        res.cons, res.cons_capacity = native.ptr(cons), 8192
        prm = native.SampleParams(170, 4, 50, 100, 100, 1, 0)
        cs = soa.c_struct()
        allow = index.allow_mask(None)
        assert native.lib().mmlst_sample(other.handle, C.byref(cs), native.ptr(allow), C.byref(prm), C.byref(res)) == native.E_ARG
        assert b"mmlst_index_upload" in native.lib().mmlst_last_error()
    finally:
        other.close()
    # H10: a DB sequence shorter than the BAM LN of a chosen reference -> IndexError like metaMLST_functions.py:267
    sidx = api.SampleIndex(ctx, index, st.ref_lens, lambda t: db.row_seq(t)[:-3])
    with pytest.raises(IndexError):
        api.type_soa(sidx, soa, 170, 4)
    # "Database is broken": more detected loci than `genes` rows for the organism (metamlst.py:188-190)
    sidx = api.SampleIndex(ctx, index, st.ref_lens, db.row_seq, genes_in_db={"ecoli": 3})
    with pytest.raises(RuntimeError, match="Database is broken"):
        api.type_soa(sidx, soa, 170, 4)
    with pytest.raises(ValueError):
        api.type_soa(sidx, soa, 170, 255)
