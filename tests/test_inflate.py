"""The whole-buffer DEFLATE decoder of the BAM unpacker (csrc/inflate_fast.cpp, `mmlst_inflate_raw`) byte for byte against
zlib: every compression level and strategy (stored / fixed-Huffman / dynamic blocks, long and short distances, runs),
BGZF-sized and larger streams, the payloads of the golden BAMs; and robustness: truncated and bit-flipped streams and a
too-small output buffer must end in an error code, never in an access outside the buffers (the output array is guarded)."""
import ctypes as C
import glob
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from metamlst_b200 import native

GUARD = 64


def inflate(data: bytes, cap: int):
    lib = native.lib()
    out = np.full(cap + 2 * GUARD, 0xA5, np.uint8)
    src = np.frombuffer(data, np.uint8) if data else np.zeros(1, np.uint8)
    got = C.c_size_t(0)
    rc = lib.mmlst_inflate_raw(src.ctypes.data_as(C.c_void_p), len(data), C.c_void_p(out.ctypes.data + GUARD), cap, C.byref(got))
    assert (out[:GUARD] == 0xA5).all() and (out[GUARD + cap:] == 0xA5).all(), "write outside the output buffer"
    return rc, bytes(out[GUARD:GUARD + got.value])


def deflate(raw: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, memlevel, strategy)
    return c.compress(raw) + c.flush()


def _payloads():
    rng = np.random.default_rng(7)
    dna = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 60000))
    yield "empty", b""
    yield "one", b"x"
    yield "run", b"a" * 70000
    yield "dna", dna
    yield "dna_repeats", (dna[:700] * 90)[:65280]
    yield "random", bytes(rng.integers(0, 256, 65280, dtype=np.uint8))
    yield "text", (b"the quick brown fox jumps over the lazy dog; " * 2000)[:65000]
    yield "mixed", bytes(rng.integers(0, 256, 3000, dtype=np.uint8)) + dna[:20000] + b"\0" * 5000 + bytes(rng.integers(0, 4, 30000, dtype=np.uint8))
    yield "big", bytes(rng.integers(0, 16, 400000, dtype=np.uint8)) + dna * 3  # beyond one BGZF block: window-sized distances


@pytest.mark.parametrize("name,raw", list(_payloads()), ids=[n for n, _ in _payloads()])
def test_matches_zlib_on_every_level_and_strategy(name, raw):
    for level in (0, 1, 2, 4, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            for memlevel in (8, 1):  # memLevel 1: many small deflate blocks inside one stream
                comp = deflate(raw, level, strategy, memlevel)
                rc, got = inflate(comp, len(raw))
                assert rc == 0 and got == raw, (name, level, strategy, memlevel, rc, len(got))
                rc2, got2 = inflate(comp, len(raw) + 100)  # a roomier buffer gives the same bytes
                assert rc2 == 0 and got2 == raw
                if len(raw):
                    assert inflate(comp, len(raw) - 1)[0] != 0  # does not fit: error, no overrun (guards checked inside)


def test_bgzf_blocks_of_the_golden_bams():
    n_blocks = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, "*", "sample.bam"))):
        data = open(path, "rb").read()
        p = 0
        while p < len(data):
            xlen = struct.unpack_from("<H", data, p + 10)[0]
            bsize = struct.unpack_from("<H", data, p + 16)[0] + 1
            isize = struct.unpack_from("<I", data, p + bsize - 4)[0]
            payload = data[p + 12 + xlen:p + bsize - 8]
            rc, got = inflate(payload, isize)
            assert rc == 0 and got == zlib.decompress(payload, -15) and len(got) == isize
            n_blocks += 1
            p += bsize
    assert n_blocks >= 9


def test_corrupt_streams_fail_cleanly():
    rng = np.random.default_rng(11)
    raw = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), 30000)) + b"Q" * 3000
    for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)):
        comp = deflate(raw, level, strategy)
        for cut in (0, 1, 2, 5, len(comp) // 3, len(comp) - 1):  # truncation: error (a stored block may be cut mid-copy)
            rc, _ = inflate(comp[:cut], len(raw))
            assert rc != 0, (level, cut)
        ok = bad = 0
        for _ in range(300):  # bit flips: error, or a different / equal output -- never a crash or a write outside
            b = bytearray(comp)
            for _k in range(int(rng.integers(1, 4))):
                i = int(rng.integers(0, len(b)))
                b[i] ^= 1 << int(rng.integers(0, 8))
            rc, got = inflate(bytes(b), len(raw))
            if rc == 0:
                ok += 1  # what it must equal when zlib accepts the stream too: test_agrees_with_zlib_on_what_is_valid
            else:
                bad += 1
        assert bad > 0
    assert inflate(b"\x07", 10)[0] != 0  # reserved block type 3
    assert inflate(b"\x01\x05\x00\xfa\xfe" + b"abcde", 5)[0] != 0  # stored block whose NLEN is not ~LEN


def test_agrees_with_zlib_on_what_is_valid():
    """For damaged streams zlib still accepts, the decoder must produce zlib's bytes (it may also accept a few streams zlib
    rejects for an incomplete code set -- those never occur in files written by a compliant deflater)."""
    rng = np.random.default_rng(13)
    raw = bytes(rng.integers(0, 6, 20000, dtype=np.uint8))
    comp = deflate(raw, 6)
    agree = 0
    for _ in range(400):
        b = bytearray(comp)
        i = int(rng.integers(0, len(b)))
        b[i] ^= 1 << int(rng.integers(0, 8))
        try:
            d = zlib.decompressobj(-15)
            want = d.decompress(bytes(b))
            finished = d.eof
        except zlib.error:
            continue
        if not finished:
            continue
        rc, got = inflate(bytes(b), len(want))
        assert rc == 0 and got == want
        agree += 1
    assert agree > 0
