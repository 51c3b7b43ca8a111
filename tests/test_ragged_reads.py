"""Ragged input through the native unpacker (CPU): reads of MANY lengths in one BAM (trimmed reads) with every CIGAR operator
BAM allows -- M I D N S H P = X -- lower-case / IUPAC bases, qualities around the threshold, several alignments per contig
start.  The unpacked streams are checked against the Python oracle working on the raw records: the score-stream fields
record by record, and the 3-bit plane rows decoded with numpy (helpers.planes_to_counts, test side) against the oracle's
pileup counts per contig, with and without the tag filter, with the depth cap active and inactive."""
import numpy as np
import pytest

import json

import helpers
from metamlst_b200 import bam
from oracle import bamio, mlst_oracle as orc

OPS = {"M": 0, "I": 1, "D": 2, "N": 3, "S": 4, "H": 5, "P": 6, "=": 7, "X": 8}


def _random_record(rng, i, tid, clen):
    ql = int(rng.integers(30, 151))
    # a CIGAR whose query-consuming ops add up to ql: optional clips, then M/=/X blocks separated by I / D / N / P
    left = int(rng.integers(0, 6)) if rng.random() < 0.3 else 0
    right = int(rng.integers(0, 6)) if rng.random() < 0.3 else 0
    body = ql - left - right
    cig = []
    if rng.random() < 0.1:
        cig.append((OPS["H"], int(rng.integers(1, 9))))
    if left:
        cig.append((OPS["S"], left))
    while body > 0:
        blk = int(min(body, rng.integers(5, 60)))
        cig.append((int(rng.choice([OPS["M"], OPS["M"], OPS["="], OPS["X"]])), blk))
        body -= blk
        if body > 0:
            k = rng.random()
            if k < 0.25:
                ins = int(min(body, rng.integers(1, 4)))
                cig.append((OPS["I"], ins)); body -= ins  # may end the read: an insertion right before the clip / the end
            elif k < 0.5:
                cig.append((OPS["D"], int(rng.integers(1, 5))))
            elif k < 0.6:
                cig.append((OPS["N"], int(rng.integers(1, 30))))
            elif k < 0.65:
                cig.append((OPS["P"], int(rng.integers(1, 3))))
    if right:
        cig.append((OPS["S"], right))
    if rng.random() < 0.1:
        cig.append((OPS["H"], int(rng.integers(1, 9))))
    qlen = sum(l for op, l in cig if op in (0, 1, 4, 7, 8))
    assert qlen == ql
    rlen = sum(l for op, l in cig if op in (0, 2, 3, 7, 8))
    pos = int(rng.integers(0, max(1, clen - rlen)))
    alphabet = np.array(list("ACGTACGTACGTACGTNacgtRY"))
    seq = "".join(rng.choice(alphabet, ql))
    qual = bytes(int(x) for x in rng.choice([2, 19, 20, 21, 35, 40], ql, p=[0.05, 0.1, 0.1, 0.1, 0.35, 0.3]))
    AS, XM = int(rng.integers(60, 300)), int(rng.integers(0, 9))
    aux = (bamio.int_aux("AS", AS), bamio.int_aux("XS", 10), bamio.int_aux("XN", 0), bamio.int_aux("XM", XM), bamio.int_aux("XO", 0))
    flag = int(rng.choice([0, 16, 256, 272]))
    return bamio.BamRecord("q%d" % (i // 3), flag, tid, pos, 42, tuple(cig), seq, qual, aux)


NAMES = ["ecoli_adk_1", "ecoli_adk_2", "ecoli_fumC_7", "saureus_arcC_3"]
LENS = [536, 536, 469, 456]


def _ragged_bam(tmp_path, seed):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(1800):
        tid = int(rng.integers(0, len(NAMES)))
        recs.append(_random_record(rng, i, tid, LENS[tid]))
    # pile many reads on one start so the cap (40) bites where it is asked to
    recs += [_random_record(rng, 5000 + i, 2, LENS[2])._replace(pos=100) for i in range(120)]
    p = str(tmp_path / "ragged.bam")
    bamio.write_bam(p, NAMES, LENS, recs)
    _h, recs = bamio.read_bam(p)  # what the file holds (BAM stores bases as 4-bit codes: case is gone, IUPAC letters stay)
    return p, recs, rng


def _oracle_counts(recs, tid, tf, max_depth):
    contig = [r for r in sorted(recs, key=lambda r: (r.tid, r.pos, (r.flag >> 4) & 1)) if r.tid == tid]
    want = np.zeros((LENS[tid], 5), np.int64)
    full, _ = orc.get_base_stats(contig, tid, 0, 20, tf, max_depth)  # depth 0: also the columns that hold only N
    for pos1, d in full.items():
        f = d["base_freq"]
        want[pos1 - 1] = [f["A"], f["C"], f["G"], f["T"], f["N"]]
    return want


@pytest.mark.parametrize("seed,max_depth", [(1, 8000), (2, 8000), (3, 40)])
def test_ragged_lengths_and_every_cigar_op(tmp_path, seed, max_depth):
    names, lens = NAMES, LENS
    p, recs, rng = _ragged_bam(tmp_path, seed)
    soa = bam.unpack_bam(p, minqual=20, max_depth=max_depth, pinned=False, threads=3)
    assert soa.chunk_qlen is None or len(recs) < 256  # ragged chunks: the per-chunk len(SEQ) form must not be offered
    # ---- score stream, record by record (file order through orig_idx)
    order = np.arange(len(recs)) if soa.orig_idx is None else np.asarray(soa.orig_idx)
    for k in rng.integers(0, len(recs), 400).tolist() + [0, len(recs) - 1]:
        r = recs[int(order[k])]
        assert int(soa.tid[k]) == r.tid and int(soa.qlen[k]) == len(r.seq)
        assert int(soa.as0[k]) == r.aux[0][2] and int(soa.xm3[k]) == r.aux[3][2]
    # the sort is `samtools sort` order: (tid, pos, reverse strand), stable
    key = [(recs[int(i)].tid, recs[int(i)].pos, (recs[int(i)].flag >> 4) & 1) for i in order]
    assert key == sorted(key)
    # ---- pileup stream: decoded planes vs the oracle's pileup of the raw records
    h = bamio.BamHeader("", names, lens)
    srt = sorted(recs, key=lambda r: (r.tid, r.pos, (r.flag >> 4) & 1))
    for tid in range(len(names)):
        contig = [r for r in srt if r.tid == tid]
        for minscore, max_xm in ((-32768, 255), (150, 4)):
            tf = None if minscore < 0 else [("AS", "loc_gte", minscore), ("XM", "loc_lte", max_xm)]
            stats, _eng = orc.get_base_stats(contig, tid, 1, 20, tf, max_depth)
            want = np.zeros((lens[tid], 5), np.int64)
            full, _ = orc.get_base_stats(contig, tid, 0, 20, tf, max_depth)  # depth 0: also the columns that hold only N
            for pos1, d in full.items():
                f = d["base_freq"]
                want[pos1 - 1] = [f["A"], f["C"], f["G"], f["T"], f["N"]]
            got = helpers.planes_to_counts(soa, tid, lens[tid], minscore, max_xm)
            assert np.array_equal(got, want), (seed, tid, minscore, np.nonzero((got != want).any(axis=1))[0][:10])
            assert sum(1 for d in stats.values()) == int((want[:, :4].sum(axis=1) >= 1).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("seed,max_depth", [(1, 8000), (3, 40)])
def test_ragged_lengths_on_the_gpu(tmp_path, seed, max_depth):
    """The same file through the kernels: stage-1 tables and dict order against metamlst.py's own loop (oracle stage1), pileup
    counts of both implementations against the oracle's pileup, consensus against the majority rule."""
    from metamlst_b200 import api, native
    p, recs, _rng = _ragged_bam(tmp_path, seed)
    soa = bam.unpack_bam(p, minqual=20, max_depth=max_depth)
    ctx = native.Context(0)
    try:
        index = api.AlleleIndex(soa.ref_names)
        h = bamio.BamHeader("", NAMES, LENS)
        for minscore, max_xm, min_len in ((150, 4, 50), (80, 5, 100), (0, 8, 0)):
            cel, total, ignored, _raw = api.score_soa(ctx, soa, index, minscore, max_xm, min_len, None, 100)
            want, _bank, wt, wi = orc.stage1(h, recs, minscore, max_xm, min_len, None, 100)
            assert json.dumps(cel) == json.dumps(want) and (total, ignored) == (wt, wi)
        tids = list(range(len(NAMES)))
        for minscore, max_xm in ((-32768, 255), (150, 4)):
            tf = None if minscore < 0 else [("AS", "loc_gte", minscore), ("XM", "loc_lte", max_xm)]
            want = [_oracle_counts(recs, t, tf, max_depth) for t in tids]
            for impl in (1, 2):
                seqs, holes, snps, counts, col_off = api.pileup_consensus(ctx, soa, tids, ["A" * n for n in LENS], minscore, max_xm, 1, impl, True)
                for i, t in enumerate(tids):
                    assert np.array_equal(counts[col_off[i]:col_off[i + 1]], want[t]), (impl, t, minscore)
                    contig = [r for r in sorted(recs, key=lambda r: (r.tid, r.pos, (r.flag >> 4) & 1)) if r.tid == t]
                    cons = orc.reference_free_consensus(contig, t, LENS[t], 1, 20, "N", tf, max_depth)
                    filled = "".join(("a" if c == "N" else c) for c in cons)  # DB base 'A', lower-cased, fills a hole (metaMLST_functions.py:267)
                    assert seqs[i] == filled and int(holes[i]) == cons.count("N") and int(snps[i]) == sum(1 for c in cons if c not in "AN")
    finally:
        ctx.close()
