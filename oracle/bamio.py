"""TEST INFRASTRUCTURE ONLY -- pure-Python BGZF/BAM writer and reader (zlib).

Used by tests/, oracle/shims and oracle/make_golden.py to create and parse small BAM fixtures without
htslib/pysam/samtools (none of which exist in this image; SURVEY.md header table).  The product never imports
this module: its own reader is the C++ unpacker in metamlst_b200/csrc/bam_unpack.cpp.

Format follows the public SAM/BAM specification (SAMv1 section 4); the reference touches BAMs only through
samtools (`/root/reference/metamlst.py:96`) and pysam (`/root/reference/cmseq/cmseq.py:54,527`).
"""
from __future__ import annotations

import struct
import zlib
from typing import Iterator, List, NamedTuple, Sequence, Tuple

import numpy as np

_SEQ_CODE = "=ACMGRSVTWYHKDBN"
_ENC = {c: i for i, c in enumerate(_SEQ_CODE)}
_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
CIGAR_CHARS = "MIDNSHP=X"


def _bgzf_block(data: bytes, level: int = 6) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, bsize)
    return hdr + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def bgzf_write(path: str, payload: bytes, level: int = 6, block: int = 0xFF00) -> None:
    with open(path, "wb") as f:
        for i in range(0, len(payload), block):
            f.write(_bgzf_block(payload[i:i + block], level))
        f.write(_EOF)


def bgzf_read(path: str) -> bytes:
    raw = open(path, "rb").read()
    out = []
    p = 0
    while p < len(raw):
        if raw[p:p + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block at %d" % p)
        xlen = struct.unpack_from("<H", raw, p + 10)[0]
        q = p + 12
        bsize = None
        while q < p + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", raw, q)
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", raw, q + 4)[0]
            q += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC field")
        cdata = raw[p + 12 + xlen:p + bsize + 1 - 8]
        out.append(zlib.decompress(cdata, -15))
        p += bsize + 1
    return b"".join(out)


def reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


class BamRecord(NamedTuple):
    qname: str
    flag: int
    tid: int
    pos: int  # 0-based
    mapq: int
    cigar: Tuple[Tuple[int, int], ...]  # (op, len)
    seq: str
    qual: bytes  # phred values, b"" when absent (0xFF filled)
    aux: Tuple[Tuple[str, str, object], ...]  # (tag, type char, value) in stored order
    mtid: int = -1
    mpos: int = -1
    tlen: int = 0

    def ref_len(self) -> int:
        return sum(l for op, l in self.cigar if op in (0, 2, 3, 7, 8))

    def cigar_string(self) -> str:
        return "".join("%d%s" % (l, CIGAR_CHARS[op]) for op, l in self.cigar) or "*"


def _pack_aux(aux: Sequence[Tuple[str, str, object]]) -> bytes:
    out = []
    for tag, typ, val in aux:
        t = tag.encode()
        if typ == "Z":
            out.append(t + b"Z" + str(val).encode() + b"\0")
        elif typ == "A":
            out.append(t + b"A" + str(val).encode()[:1])
        elif typ == "f":
            out.append(t + b"f" + struct.pack("<f", float(val)))
        elif typ in "cCsSiI":
            out.append(t + typ.encode() + struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[typ], int(val)))
        else:
            raise ValueError("aux type %r" % typ)
    return b"".join(out)


def int_aux(tag: str, v: int) -> Tuple[str, str, int]:
    """Smallest integer type, the way htslib/bowtie2 writers choose it."""
    v = int(v)
    if v >= 0:
        typ = "C" if v <= 0xFF else ("S" if v <= 0xFFFF else "I")
    else:
        typ = "c" if v >= -128 else ("s" if v >= -32768 else "i")
    return (tag, typ, v)


def pack_record(r: BamRecord) -> bytes:
    name = r.qname.encode() + b"\0"
    l_seq = len(r.seq)
    end = r.pos + max(r.ref_len(), 1)
    cig = b"".join(struct.pack("<I", (l << 4) | op) for op, l in r.cigar)
    s = r.seq + ("=" if l_seq & 1 else "")
    sq = bytes((_ENC.get(s[i], 15) << 4) | _ENC.get(s[i + 1], 15) for i in range(0, len(s), 2))
    if l_seq & 1:
        sq = sq[:-1] + bytes([sq[-1] & 0xF0])
    ql = bytes(r.qual) if len(r.qual) == l_seq else b"\xff" * l_seq
    body = struct.pack("<iiBBHHHiiii", r.tid, r.pos, len(name), r.mapq, reg2bin(r.pos, end), len(r.cigar), r.flag,
                       l_seq, r.mtid, r.mpos, r.tlen) + name + cig + sq + ql + _pack_aux(r.aux)
    return struct.pack("<i", len(body)) + body


def write_bam(path: str, ref_names: Sequence[str], ref_lens: Sequence[int], records, sort_order: str = "unknown",
              level: int = 1) -> None:
    text = "@HD\tVN:1.0\tSO:%s\n" % sort_order + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in zip(ref_names, ref_lens))
    text += "@PG\tID:bowtie2\tPN:bowtie2\tVN:2.4.4\n"
    tb = text.encode()
    parts = [b"BAM\1", struct.pack("<i", len(tb)), tb, struct.pack("<i", len(ref_names))]
    for n, l in zip(ref_names, ref_lens):
        nb = n.encode() + b"\0"
        parts.append(struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(l)))
    for r in records:
        parts.append(r if isinstance(r, (bytes, bytearray)) else pack_record(r))
    bgzf_write(path, b"".join(parts), level)


def _parse_aux(buf: bytes, p: int, end: int):
    out = []
    while p < end:
        tag = buf[p:p + 2].decode()
        typ = chr(buf[p + 2])
        p += 3
        if typ == "Z" or typ == "H":
            q = buf.index(b"\0", p)
            out.append((tag, typ, buf[p:q].decode()))
            p = q + 1
        elif typ == "A":
            out.append((tag, typ, chr(buf[p])))
            p += 1
        elif typ == "f":
            out.append((tag, typ, struct.unpack_from("<f", buf, p)[0]))
            p += 4
        elif typ in "cCsSiI":
            fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[typ]
            out.append((tag, typ, struct.unpack_from("<" + fmt, buf, p)[0]))
            p += struct.calcsize(fmt)
        elif typ == "B":
            sub = chr(buf[p])
            cnt = struct.unpack_from("<i", buf, p + 1)[0]
            fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
            vals = struct.unpack_from("<%d%s" % (cnt, fmt), buf, p + 5)
            out.append((tag, typ, (sub, vals)))
            p += 5 + cnt * struct.calcsize(fmt)
        else:
            raise ValueError("aux type %r" % typ)
    return tuple(out)


class BamHeader(NamedTuple):
    text: str
    ref_names: List[str]
    ref_lens: List[int]


def read_bam(path: str) -> Tuple[BamHeader, List[BamRecord]]:
    buf = bgzf_read(path)
    if buf[:4] != b"BAM\1":
        raise ValueError("not a BAM file")
    l_text = struct.unpack_from("<i", buf, 4)[0]
    text = buf[8:8 + l_text].split(b"\0")[0].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", buf, p)[0]
    p += 4
    names, lens = [], []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", buf, p)[0]
        names.append(buf[p + 4:p + 4 + ln - 1].decode())
        lens.append(struct.unpack_from("<i", buf, p + 4 + ln)[0])
        p += 8 + ln
    recs = []
    while p < len(buf):
        bs = struct.unpack_from("<i", buf, p)[0]
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", buf, p + 4)
        q = p + 36
        qname = buf[q:q + l_name - 1].decode()
        q += l_name
        cig = tuple((c & 0xF, c >> 4) for c in struct.unpack_from("<%dI" % n_cig, buf, q))
        q += 4 * n_cig
        sb = buf[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = "".join(_SEQ_CODE[b >> 4] + _SEQ_CODE[b & 15] for b in sb)[:l_seq]
        qual = buf[q:q + l_seq]
        if l_seq and qual[0] == 0xFF:
            qual = b""
        q += l_seq
        aux = _parse_aux(buf, q, p + 4 + bs)
        recs.append(BamRecord(qname, flag, tid, pos, mapq, cig, seq, qual, aux, mtid, mpos, tlen))
        p += 4 + bs
    return BamHeader(text, names, lens), recs


def sam_line(h: BamHeader, r: BamRecord) -> str:
    """One `samtools view` text line (SAMv1 section 1.4); integer aux types all print as `:i:`."""
    rname = h.ref_names[r.tid] if r.tid >= 0 else "*"
    rnext = "*" if r.mtid < 0 else ("=" if r.mtid == r.tid else h.ref_names[r.mtid])
    qual = "".join(chr(q + 33) for q in r.qual) if r.qual else "*"
    f = [r.qname, str(r.flag), rname, str(r.pos + 1), str(r.mapq), r.cigar_string(), rnext, str(r.mpos + 1), str(r.tlen),
         r.seq or "*", qual]
    for tag, typ, val in r.aux:
        if typ in "cCsSiI":
            f.append("%s:i:%d" % (tag, val))
        elif typ == "f":
            f.append("%s:f:%g" % (tag, val))
        elif typ == "B":
            f.append("%s:B:%s,%s" % (tag, val[0], ",".join(str(v) for v in val[1])))
        else:
            f.append("%s:%s:%s" % (tag, typ, val))
    return "\t".join(f)


def table_records(tab) -> Iterator[BamRecord]:
    """AlnTable (metamlst_b200.synth) -> BamRecord stream; bowtie2 aux order AS,XS,XN,XM,XO,XG,NM,YT."""
    for i in range(tab.n):
        c0, c1 = int(tab.cig_off[i]), int(tab.cig_off[i + 1])
        cig = tuple((int(c) & 0xF, int(c) >> 4) for c in tab.cig_ops[c0:c1])
        aux = [int_aux("AS", tab.AS[i])]
        if tab.has_xs[i]:
            aux.append(int_aux("XS", tab.XS[i]))
        aux += [int_aux("XN", tab.XN[i]), int_aux("XM", tab.XM[i]), int_aux("XO", tab.XO[i]), int_aux("XG", tab.XG[i]),
                int_aux("NM", tab.NM[i]), ("YT", "Z", "UU")]
        yield BamRecord("r%d" % int(tab.qname_id[i]), int(tab.flag[i]), int(tab.tid[i]), int(tab.pos[i]), 42, cig,
                        tab.seq[i].tobytes().decode(), tab.qual[i].tobytes(), tuple(aux))


def write_table_bam(path: str, tab, sort_order: str = "unknown") -> None:
    write_bam(path, tab.ref_names, [int(x) for x in tab.ref_lens], table_records(tab), sort_order)
