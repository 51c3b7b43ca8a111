"""pysam stand-in (TEST INFRASTRUCTURE) built on oracle.bamio + oracle.mlst_oracle.PileupEngine.
Implements the surface cmseq uses: cmseq/cmseq.py:50-54,76-88,527-545."""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import bamio  # noqa: E402
from oracle.mlst_oracle import PileupEngine, HTS_MAX_DEPTH_DEFAULT, refuse_proper_pairs  # noqa: E402

__version__ = "0.0-shim"


class AlignedSegment:
    def __init__(self, rec):
        self._r = rec
        self.query_name = rec.qname
        self.query_sequence = rec.seq if rec.seq else None
        self.query_qualities = list(rec.qual) if rec.qual else None
        self.query_length = len(rec.seq)
        self.flag = rec.flag
        self.reference_start = rec.pos

    def get_tag(self, tag):
        for t, _typ, v in self._r.aux:
            if t == tag:
                return v
        raise KeyError("tag '%s' not present" % tag)


class PileupRead:
    def __init__(self, aln, qpos, kind):
        self.alignment = aln
        self.is_del = 1 if kind != 0 else 0
        self.is_refskip = 1 if kind == 2 else 0
        self.query_position = qpos if kind == 0 else None
        self.query_position_or_next = qpos


class PileupColumn:
    def __init__(self, tid, pos, plp, min_base_quality, cache):
        self.reference_id = tid
        self.pos = self.reference_pos = pos
        self._plp = plp
        self._minq = min_base_quality
        self._cache = cache
        self.nsegments = self.n = len(plp)

    @property
    def pileups(self):
        out = []
        for r, qpos, kind in self._plp:
            # pysam pileup_base_qual_skip(): qpos >= l_qseq counts as quality 0; missing quals are 0xFF
            if qpos < len(r.seq):
                c = r.qual[qpos] if r.qual else 255
            else:
                c = 0
            if c < self._minq:
                continue
            a = self._cache.get(id(r))
            if a is None:
                a = self._cache[id(r)] = AlignedSegment(r)
            out.append(PileupRead(a, qpos, kind))
        return out


class AlignmentFile:
    def __init__(self, path, mode="rb", **kw):
        self.filename = path
        self.header, self._records = bamio.read_bam(path)
        self.references = tuple(self.header.ref_names)
        self.lengths = tuple(self.header.ref_lens)
        self.nreferences = len(self.references)
        self._closed = False

    def _contig_records(self, contig):
        tid = self.references.index(contig)
        recs = [r for r in self._records if r.tid == tid]
        for a, b in zip(recs, recs[1:]):
            if b.pos < a.pos:
                raise ValueError("fetch called on bamfile without index")  # pileup needs a sorted+indexed BAM
        return tid, recs

    def count(self, contig=None, read_callback="nofilter", **kw):
        return len(self._contig_records(contig)[1])

    def pileup(self, contig=None, stepper="all", min_base_quality=13, max_depth=HTS_MAX_DEPTH_DEFAULT,
               ignore_overlaps=True, **kw):
        if stepper != "nofilter":
            raise NotImplementedError("shim implements stepper='nofilter' only (cmseq/cmseq.py:527)")
        tid, recs = self._contig_records(contig)
        if ignore_overlaps:
            refuse_proper_pairs(recs)  # H2
        eng = PileupEngine(tid, max_depth)
        cache = {}
        for pos, plp in eng.columns(recs):
            yield PileupColumn(tid, pos, plp, min_base_quality, cache)

    def close(self):
        self._closed = True


def index(path, *a, **kw):
    return None


def sort(*a, **kw):
    raise NotImplementedError
