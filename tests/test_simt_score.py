"""The warp program of every form of the run-length score kernel, run WITHOUT a GPU: tests/simt compiles the kernels' own text
(metamlst_b200/csrc/score_runs_kernels.cuh) for the host, one std::thread per lane (warp collectives, barriers, mbarriers and
TMA bulk copies emulated), and the result tables are compared with numpy.  Covers control flow and integer arithmetic
(run walking, chunk/run boundary cases, SWAR filters, tails, the ring's stage/phase bookkeeping); memory ordering and speed
are the GPU tests' business."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from metamlst_b200 import packing

SIMT = os.path.join(ROOT, "tests", "simt")


@pytest.fixture(scope="module")
def simt():
    so = os.path.join(SIMT, "libsimt_score.so")
    srcs = [os.path.join(SIMT, "simt_score.cpp"), os.path.join(SIMT, "simt_host_emul.h"),
            os.path.join(ROOT, "metamlst_b200", "csrc", "score_runs_kernels.cuh"), os.path.join(ROOT, "metamlst_b200", "csrc", "common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", SIMT,
                               "-I", os.path.join(ROOT, "metamlst_b200", "csrc"), "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.simt_score_runs.restype = C.c_int
    lib.simt_score_runs.argtypes = [C.c_int, C.c_uint] + [C.c_void_p] * 2 + [C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint64, C.c_uint64, C.c_void_p,
                                                                                                       C.c_uint32, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _np_score(tid, as0, xm3, qlen, idx, allow, n_ref, minscore, max_xm, min_len):
    al = allow[tid] != 0
    ok = al & (as0 >= minscore) & (qlen >= min_len) & (xm3 <= max_xm)
    s = np.zeros(n_ref, np.int64); c = np.zeros(n_ref, np.int64); f = np.full(n_ref, 0xFFFFFFFF, np.int64)
    np.add.at(s, tid[ok], as0[ok].astype(np.int64)); np.add.at(c, tid[ok], 1); np.minimum.at(f, tid[ok], idx[ok].astype(np.int64))
    return s, c.astype(np.uint32), f.astype(np.uint32), int(al.sum()), int((al & ~ok).sum())


def _streams(n, style, rng, n_ref):
    if style == "long":
        lens = rng.integers(1, max(2, n // 3), 64)
    elif style == "short":
        lens = rng.integers(1, 4, min(n, 200000))
    elif style == "config2":  # ~1900 records per run, the bench workload's shape
        lens = rng.integers(1500, 2300, max(1, n // 1500))
    else:
        lens = np.concatenate([rng.integers(1, 900, 4000), [256, 256, 512, 1, 1, 255, 257, 768]])
        rng.shuffle(lens)
    lens = lens[np.cumsum(lens) <= n]
    if lens.sum() < n:
        lens = np.concatenate([lens, [n - lens.sum()]])
    rt = rng.integers(0, n_ref + 3, lens.shape[0])  # ids >= n_ref: records of references the table does not know
    rt[1:][rt[1:] == rt[:-1]] += 1
    rt %= n_ref + 3
    rt[1:][rt[1:] == rt[:-1]] = (rt[1:][rt[1:] == rt[:-1]] + 1) % (n_ref + 3)
    return np.repeat(rt, lens).astype(np.uint32)


# forms 0 / 1: register-staged kernels; 2..5: ring configurations; 12..15: the ring builds with the L2 hints.  The library
# default (5) and the experimental form 6 see every size; the other forms the sizes that exercise boundaries, refills and tails.
_SIZES = [1, 255, 256, 257, 511, 513, 769, 4096, 30011]
_CASES = ([(n, f) for f in (5, 6) for n in _SIZES] + [(n, f) for f in (0, 1) for n in (1, 256, 257, 513, 4096, 30011)] +
          [(n, f) for f in (2, 3, 4, 12, 15) for n in (257, 30011)])  # 6: pair-fused ring


@pytest.mark.parametrize("n,form", _CASES)
def test_warp_program_of_every_kernel_form_matches_numpy(simt, n, form):
    rng = np.random.default_rng(1000 * form + n)
    n_ref = 300
    for style in (("mixed", "long", "short", "config2") if n < 20000 else ("mixed", "config2")):  # the emulation costs ~0.4 s per launch at 30 k records
        tid = _streams(n, style, rng, n_ref)
        as0 = rng.integers(-50, 301, n).astype(np.int16)
        as0[rng.integers(0, n, max(1, n // 1000))] = 32767
        as0[rng.integers(0, n, max(1, n // 1000))] = -32768
        xm3 = rng.integers(0, 8, n).astype(np.uint8)
        xm3[rng.integers(0, n, max(1, n // 500))] = 255
        qlen_rec = rng.integers(30, 160, n).astype(np.uint16)
        qlen_chunk = np.repeat(rng.integers(30, 160, (n + 255) // 256), 256)[:n].astype(np.uint16)
        allow = np.zeros(n_ref + 8, np.uint8)
        allow[:n_ref] = rng.random(n_ref) < 0.8
        for oidx, qlen in [(o, q) for o in (None, rng.permutation(n).astype(np.uint32)) for q in (qlen_rec, qlen_chunk)]:
            if form >= 2 and oidx is not None:
                continue  # the library sends records with a file-order index to form 0
            if form == 6 and qlen is not qlen_chunk and n != 1:
                continue  # form 6 exists for the per-chunk len(SEQ) stream (others take form 5)
            soa = packing.SoaHost([], np.zeros(0, np.int32), tid, as0, xm3, qlen, oidx, np.zeros(0, packing.PREC_DTYPE), np.zeros(0, np.uint32), 0,
                                  np.zeros(1, np.uint64)).build_runs(max_fraction=1.0)
            assert soa.run_tid is not None and (soa.chunk_qlen is not None) == (qlen is qlen_chunk or n == 1)
            idx = np.arange(n) if oidx is None else oidx
            known = np.minimum(tid, n_ref)  # ids >= n_ref behave as a filtered allele
            for grid in (1, 3):
                for minscore, max_xm, min_len in ((100, 5, 50), (-32768, 255, 0), (40000, 5, 50), (100, -1, 50), (100, 5, 70000)):
                    if (minscore, max_xm, min_len) != (100, 5, 50) and (grid != 1 or style != "mixed"):
                        continue
                    sum_as = np.zeros(n_ref, np.int64); n_hit = np.zeros(n_ref, np.uint32); first = np.full(n_ref, 0xFFFFFFFF, np.uint32)
                    counters = np.zeros(2, np.uint64)
                    rc = simt.simt_score_runs(form, grid, _ptr(soa.run_tid), _ptr(soa.run_start), int(soa.run_tid.shape[0]), _ptr(soa.chunk_run),
                                              _ptr(as0), _ptr(xm3), _ptr(None if soa.chunk_qlen is not None else qlen), _ptr(soa.chunk_qlen), _ptr(oidx), n, 7,
                                              _ptr(allow), n_ref, minscore, max_xm, min_len, _ptr(sum_as), _ptr(n_hit), _ptr(first), _ptr(counters))
                    assert rc == 0
                    allow_x = np.concatenate([allow[:n_ref], [0]])
                    ws, wc, wf, wt, wi = _np_score(known, as0.astype(np.int64), xm3.astype(int), qlen.astype(int), idx + (7 if oidx is None else 0),
                                                   allow_x, n_ref + 1, minscore, max_xm, min_len)
                    where = (n, style, form, grid, oidx is None, qlen is qlen_chunk, minscore, max_xm, min_len)
                    assert np.array_equal(sum_as, ws[:n_ref]), where
                    assert np.array_equal(n_hit, wc[:n_ref]) and np.array_equal(first, wf[:n_ref]), where
                    assert (int(counters[0]), int(counters[1])) == (wt, wi), where
