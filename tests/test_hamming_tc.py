"""The tensor-core form of the all-pairs closest-allele sweep (csrc/hamming_tc.cu: tcgen05.mma kind::f8f6f4, accumulators in TMEM)
against the XOR+POPC kernel (itself checked against the oracle's stringDiff in test_gpu_parity.py) and against the C port of the
oracle: identical best[] = min over rows of (distance << 32 | row), ties to the lowest row, zip truncation, flagged sequences skipped."""
import numpy as np
import pytest
import torch

from metamlst_b200 import native, packing
from oracle import corc

pytestmark = pytest.mark.gpu


def _planes(seqs, W):
    hi, lo, ln, _xi, _xx, _xb = packing.encode_2bit_x(seqs, W)
    return hi, lo, ln


def _tc_search(q_hi, q_lo, q_len, d_hi_t, d_lo_t, d_len, n_q, n_rows, W, base=0):
    lib = native.lib()
    dev = "cuda:0"
    st = torch.cuda.current_stream().cuda_stream
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32 if a.dtype == np.uint32 else np.int16)).to(dev)
    qh, ql, qn, dh, dl, dn = t(q_hi), t(q_lo), t(q_len), t(d_hi_t), t(d_lo_t), t(d_len)
    qimg = torch.empty(int(lib.mmlst_hamming_tc_image_bytes(n_q, W, 128)), dtype=torch.uint8, device=dev)
    dimg = torch.empty(int(lib.mmlst_hamming_tc_image_bytes(n_rows, W, 256)), dtype=torch.uint8, device=dev)
    qmax = torch.zeros((n_q + 127) // 128, dtype=torch.int32, device=dev)
    dmax = torch.zeros((n_rows + 255) // 256, dtype=torch.int32, device=dev)
    native.check(lib.mmlst_hamming_tc_expand_dev(native.ptr(qh), native.ptr(ql), native.ptr(qn), n_q, W, 0, 128, native.ptr(qimg), native.ptr(qmax), st))
    native.check(lib.mmlst_hamming_tc_expand_dev(native.ptr(dh), native.ptr(dl), native.ptr(dn), n_rows, W, 1, 256, native.ptr(dimg), native.ptr(dmax), st))
    best = torch.full((n_q,), -1, dtype=torch.int64, device=dev)
    native.check(lib.mmlst_hamming_tc_search_dev(native.ptr(qimg), native.ptr(qmax), native.ptr(qn), n_q, native.ptr(dimg), native.ptr(dmax), native.ptr(dn),
                                                 n_rows, W, base, native.ptr(best), st))
    torch.cuda.synchronize()
    # the reference kernel on the same planes
    blocks = torch.tensor([0, n_q, 0, n_rows], dtype=torch.int32, device=dev)
    want = torch.full((n_q,), -1, dtype=torch.int64, device=dev)
    native.check(lib.mmlst_hamming_min_dev(native.ptr(dh), native.ptr(dl), native.ptr(dn), n_rows, W, native.ptr(qh), native.ptr(ql), native.ptr(qn), n_q,
                                           native.ptr(blocks), 1, base, native.ptr(want), st))
    torch.cuda.synchronize()
    return best.cpu().numpy().view(np.uint64), want.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("n_rows,n_q,W,seed", [(700, 150, 8, 1), (5000, 300, 24, 2), (256, 128, 8, 3), (1025, 129, 16, 4), (40, 3, 24, 5)])
def test_tensor_core_sweep_equals_the_popc_kernel_and_the_oracle(n_rows, n_q, W, seed):
    rng = np.random.default_rng(seed)
    L = 32 * W
    base_seq = rng.integers(0, 4, L)
    letters = np.frombuffer(b"ACGT", np.uint8)

    def mutate(k):
        s = base_seq.copy()
        p = rng.choice(L, size=k, replace=False)
        s[p] = (s[p] + rng.integers(1, 4, k)) % 4
        return s

    rows = []
    for r in range(n_rows):
        ln = int(rng.integers(max(1, L - 200), L + 1)) if r % 7 else int(rng.integers(1, L + 1))
        rows.append(letters[mutate(int(rng.integers(0, 12)))[:ln]].tobytes())
    rows[3] = rows[2]                      # a tie: the lower row must win
    qs = []
    for q in range(n_q):
        src = rows[int(rng.integers(0, n_rows))]
        s = np.frombuffer(src, np.uint8).copy()
        k = int(rng.integers(0, 6))
        if k and len(s) > k:
            p = rng.choice(len(s), size=k, replace=False)
            s[p] = letters[(np.searchsorted(letters, s[p]) + rng.integers(1, 4, k)) % 4]
        ln = int(rng.integers(1, L + 1)) if q % 5 == 0 else len(s)
        s = np.concatenate([s, letters[rng.integers(0, 4, max(0, ln - len(s)))]])[:ln]
        qs.append(s.tobytes())
    qs[0] = rows[2]
    d_hi, d_lo, d_len = _planes(rows, W)
    if n_rows > 10:
        d_len = d_len.copy(); d_len[5] |= 0x8000   # a flagged row: both kernels leave it to the exact path
    q_hi, q_lo, q_len = _planes(qs, W)
    th, tl = packing.tile_db(d_hi, d_lo)
    got, want = _tc_search(q_hi, q_lo, q_len, th, tl, d_len, n_q, n_rows, W, base=1000)
    assert np.array_equal(got, want)
    # and the oracle's stringDiff on the raw strings (flagged row excluded)
    flat = np.frombuffer(b"".join(rows), np.uint8)
    off = np.zeros(n_rows + 1, np.int64); off[1:] = np.cumsum([len(r) for r in rows])
    sub = list(range(min(n_q, 40)))
    for q in sub:
        best = None
        for r in range(n_rows):
            if n_rows > 10 and r == 5:
                continue
            a, b = qs[q], rows[r]
            m = min(len(a), len(b))
            d = int((np.frombuffer(a[:m], np.uint8) != np.frombuffer(b[:m], np.uint8)).sum())
            if best is None or d < best[0]:
                best = (d, r)
        assert (int(got[q] >> np.uint64(32)), int(got[q] & np.uint64(0xffffffff)) - 1000) == best, q
    assert int(got[0] >> np.uint64(32)) == 0 and int(got[0] & np.uint64(0xffffffff)) - 1000 == 2


def test_rows_longer_than_1024_bases_take_the_any_width_kernel():
    """W > 32 plane words (loci longer than 1024 bp) used to be refused: the any-width kernel answers them, same contract."""
    from metamlst_b200 import api
    rng = np.random.default_rng(9)
    letters = np.frombuffer(b"ACGT", np.uint8)
    base = rng.integers(0, 4, 1500)
    rows = []
    for r in range(300):
        s = base.copy()
        p = rng.choice(1500, size=int(rng.integers(0, 9)), replace=False)
        s[p] = (s[p] + 1) % 4
        rows.append(("sp", "g%d" % (r % 3), r + 1, letters[s[: int(rng.integers(1100, 1501))]].tobytes().decode()))
    rows[7] = rows[7][:3] + (rows[7][3][:600] + "N" + rows[7][3][601:],)   # a flagged row on the exact path
    ctx = native.Context(0)
    idx = api.HammingIndex(ctx, rows)
    assert idx.W > 32
    qs = [rows[i][3][: 1200 + i] for i in (1, 5, 7, 100)] + [letters[base].tobytes().decode()]
    genes = ["g1", "g2", "g1", "g1", "g0"]
    d, a = idx.search(qs, [idx.block[("sp", g)] for g in genes])
    flat = np.frombuffer("".join(r[3] for r in idx.rows).encode(), np.uint8)
    off = np.zeros(len(idx.rows) + 1, np.int64); off[1:] = np.cumsum([len(r[3]) for r in idx.rows])
    wd, wa = corc.hamming_min([q.encode() for q in qs], flat, off, [idx.block[("sp", g)] for g in genes])
    assert d.tolist() == wd.tolist() and a.tolist() == wa.tolist()
    ctx.close()
