"""Seam S2' -- the slice of cmseq's API that sits on the pileup hot path, over the GPU count tensor.

Mirrors `cmseq/cmseq.py` for users of cmseq other than MetaMLST: `BamFile(bam, filterInputList=...)`
(`cmseq/cmseq.py:35-88`) -> `get_contig_by_label(name)` -> `BamContig.get_base_stats` (`:507-569`),
`reference_free_consensus` (`:226-241`), `majority_rule` / `majority_rule_polymorphicLoci` (`:202-224`),
`polymorphism_rate` (`:430-456`), `breadth_and_depth_of_coverage` / `depth_of_coverage` / `breadth_of_coverage`
(`:459-505`), `get_all_base_values` (`:572-578`).  Same names, arguments, return shapes and key order.

What replaces pysam: the BAM is unpacked once per base-quality threshold by the native unpacker (`bam.unpack_bam`; the
threshold is folded into the bit-planes, H3) and a contig's five counters per column come from the bit-sliced pileup
kernel (`mmlst_pileup_consensus`).  Everything after the counters (ratio, binomial p-value, dict layout, the consensus
rule) is the reference's own per-column arithmetic in Python floats, so values are identical, not merely close.

Files MetaMLST itself would crash on are accepted the way pysam accepts them (`lenient_tags`): records without integer 1st / 4th aux fields
or without AS:i / XM:i tags take part in the pileup; only a call that passes a `BAM_tagFilter` on such a file raises the reference's KeyError.
Proper-pair mates are still refused (`MMLST_E_PAIRED`): pysam's `ignore_overlaps` handling (H2) is not implemented.

Not carried over (each raises, nothing degrades silently): `trimReads` (the bit-planes carry no read coordinate),
`BAM_tagFilter` entries other than `('AS','loc_gte',x)` / `('XM','loc_lte',y)` (the only ones MetaMLST passes,
`metaMLST_functions.py:259`), `stepper='all'`, the GFF/codon functions (`parse_gff`, `baseline_PSR`,
`get_base_stats_for_poly`, `easy_polymorphism_rate`) and the multiprocessing wrapper -- off the hot path.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import api, bam as bam_mod, native


class CMSEQ_DEFAULTS:  # cmseq/cmseq.py:25-32
    minqual = 30
    mincov = 1
    minlen = 0
    poly_error_rate = 0.001
    poly_pvalue_threshold = 0.01
    poly_dominant_frq_thrsh = 0.8
    trimReads = None


PYSAM_DEFAULT_MIN_BASE_QUALITY = 13  # what `pileup()` applies when cmseq does not pass one (cmseq/cmseq.py:469)
_NO_MINSCORE, _NO_MAX_XM = -32768, 255  # thresholds every record passes (as_named is i16, xm_named u8)


def _tag_filter(BAM_tagFilter) -> Tuple[int, int]:
    """(minscore, max_xM) of a cmseq BAM_tagFilter.  cmseq evaluates `all(func(tag value, limit))` per read
    (`cmseq/cmseq.py:545`); the kernels implement AS >= x and XM <= y."""
    minscore, max_xm = _NO_MINSCORE, _NO_MAX_XM
    for ent in (BAM_tagFilter or ()):
        tag, func, limit = ent
        if tag == "AS" and func == "loc_gte":
            minscore = max(minscore, int(limit))
        elif tag == "XM" and func == "loc_lte":
            max_xm = min(max_xm, int(limit))
        else:
            raise NotImplementedError("BAM_tagFilter entry %r: only ('AS','loc_gte',x) and ('XM','loc_lte',y) run on the device" % (ent,))
    return minscore, max(max_xm, -1)


def _contig_counts(ctx: native.Context, soa, tid: int, minscore: int, max_xm: int) -> np.ndarray:
    """counts[len][5] (A,C,G,T,N) of one contig from the pileup kernel.  The seam the CPU tests answer with the oracle."""
    ln = int(soa.ref_lens[tid])
    _seqs, _holes, _snps, counts, _off = api.pileup_consensus(ctx, soa, [tid], ["N" * ln], minscore, max_xm, 1, 0, True)
    return counts


class BamContig:
    coverage = None
    consensus = ''
    name = None
    length = None
    stepper = 'nofilter'
    annotations = None

    def __init__(self, bamHandle: "BamFile", contigName: str, contigLength: int, stepper: str = 'nofilter'):
        self.name = contigName
        self.length = contigLength
        self.bam_handle = bamHandle
        self.stepper = stepper
        self.annotations = []

    def set_stepper(self, ns):
        if ns in ['all', 'nofilter']:
            self.stepper = ns

    # Consensus rules (cmseq/cmseq.py:202-224).  Upstream takes `max(sorted(freq), key=freq.get)`: the keys sort to A, C, G, N, T and
    # max() keeps the first maximum, so ties resolve A > C > G > N > T and N competes with the bases (H8).  Written out here as
    # the scan it amounts to.  They are plain functions in the class body, as upstream, so `consensus_rule=BamContig.majority_rule`
    # and calling `rule(column)` both work.
    _TIE_ORDER = ("A", "C", "G", "N", "T")

    def _first_maximum(freq) -> str:
        best, best_n = None, None
        for base in BamContig._TIE_ORDER:
            if base in freq and (best is None or freq[base] > best_n):
                best, best_n = base, freq[base]
        return best

    def majority_rule(data_array):
        freq = data_array['base_freq']
        return BamContig._first_maximum(freq) if max(freq.values()) > 0 else 'N'

    def majority_rule_polymorphicLoci(data_array):
        if data_array['p'] <= 0.05:  # polymorphic under the binomial test: masked
            return "*"
        freq = data_array['base_freq']
        if max(n for base, n in freq.items() if base != 'N') > 0:
            return BamContig._first_maximum(freq)
        return 'N'

    def _counts(self, min_base_quality: int, BAM_tagFilter=None) -> np.ndarray:
        if self.stepper != 'nofilter':
            raise NotImplementedError("stepper=%r: only 'nofilter' (what cmseq passes by default, cmseq/cmseq.py:40) is implemented" % (self.stepper,))
        minscore, max_xm = _tag_filter(BAM_tagFilter)
        h = self.bam_handle
        soa = h._soa(int(min_base_quality))
        if BAM_tagFilter and getattr(soa, "n_untagged", 0):
            # pysam's get_tag raises for the first read without the tag (cmseq/cmseq.py:545); the unpacker counts such reads per FILE, so this is
            # raised for any contig of a file that holds one -- stricter than upstream, never silently different
            raise KeyError("tag 'AS' / 'XM' not present in %d pileup records of %s: a BAM_tagFilter cannot be evaluated (cmseq/cmseq.py:545)" % (soa.n_untagged, h.bamFile))
        return np.asarray(_contig_counts(h.ctx, soa, h._tid[self.name], minscore, max_xm))

    def get_base_stats(self, min_read_depth=CMSEQ_DEFAULTS.mincov, min_base_quality=CMSEQ_DEFAULTS.minqual, error_rate=CMSEQ_DEFAULTS.poly_error_rate,
                       dominant_frq_thrsh=CMSEQ_DEFAULTS.poly_dominant_frq_thrsh, BAM_tagFilter=None, trimReads=None):
        """cmseq/cmseq.py:507-569: {1-based position: {'p', 'ratio_max2all', 'base_cov', 'base_freq': {'A','T','C','G','N'}}} for
        the columns whose A+C+G+T count reaches min_read_depth, in ascending position order."""
        if trimReads:
            raise NotImplementedError("trimReads is not carried by the packed pileup stream")
        if min_read_depth < 1:
            # upstream divides by base_sum for every column htslib visits, covered by counted bases or not (cmseq/cmseq.py:554-555)
            raise ValueError("min_read_depth must be >= 1")
        counts = self._counts(min_base_quality, BAM_tagFilter)
        base_stats = defaultdict(dict)
        acgt = counts[:, :4].astype(np.int64)
        base_sums = acgt.sum(axis=1)
        binom = None
        for col in np.nonzero(base_sums >= min_read_depth)[0].tolist():
            a, c, g, t, n = (int(x) for x in counts[col])
            base_freq = {'A': a, 'T': t, 'C': c, 'G': g, 'N': n}
            base_sum = a + t + c + g
            base_max = float(max(a, t, c, g))
            r = base_max / base_sum
            if r < dominant_frq_thrsh:
                if binom is None:
                    from scipy import stats  # the reference imports it at module level (cmseq/cmseq.py:9)
                    binom = stats.binom
                p = binom.cdf(base_max, base_sum, 1.0 - error_rate)
            else:
                p = 1.0
            pos = col + 1
            base_stats[pos]['p'] = p
            base_stats[pos]['ratio_max2all'] = r
            base_stats[pos]['base_cov'] = base_sum
            base_stats[pos]['base_freq'] = base_freq
        return base_stats

    def reference_free_consensus(self, consensus_rule=majority_rule, mincov=CMSEQ_DEFAULTS.mincov, minqual=CMSEQ_DEFAULTS.minqual,
                                 dominant_frq_thrsh=CMSEQ_DEFAULTS.poly_dominant_frq_thrsh, noneCharacter='-', BAM_tagFilter=None, trimReads=None):
        """cmseq/cmseq.py:226-241: the rule's call for every column that has statistics, `noneCharacter` elsewhere; the string is
        also kept in `self.consensus`."""
        stats = self.get_base_stats(min_read_depth=mincov, min_base_quality=minqual, dominant_frq_thrsh=dominant_frq_thrsh,
                                    BAM_tagFilter=BAM_tagFilter, trimReads=trimReads, error_rate=CMSEQ_DEFAULTS.poly_error_rate)
        calls = [noneCharacter] * self.length
        for pos1, column in stats.items():
            calls[pos1 - 1] = consensus_rule(column)
        self.consensus = ''.join(calls)
        return self.consensus

    def polymorphism_rate(self, mincov=CMSEQ_DEFAULTS.mincov, minqual=CMSEQ_DEFAULTS.minqual, pvalue=CMSEQ_DEFAULTS.poly_pvalue_threshold,
                          error_rate=CMSEQ_DEFAULTS.poly_error_rate, dominant_frq_thrsh=CMSEQ_DEFAULTS.poly_dominant_frq_thrsh):
        """cmseq/cmseq.py:430-456: a column is polymorphic when its p-value is under `pvalue` AND its dominant base is under
        `dominant_frq_thrsh`; the distribution keys appear only when there is at least one such column (same keys, same order)."""
        stats = self.get_base_stats(min_read_depth=mincov, min_base_quality=minqual, error_rate=error_rate, dominant_frq_thrsh=dominant_frq_thrsh)
        n_cov = len(stats)
        ratios = [col['ratio_max2all'] for col in stats.values() if col['p'] < pvalue and col['ratio_max2all'] < dominant_frq_thrsh]
        rv = {'total_covered_bases': n_cov, 'total_polymorphic_bases': 0}
        if n_cov == 0:
            return rv
        rv['total_polymorphic_bases'] = len(ratios)
        rv['total_polymorphic_rate'] = float(len(ratios)) / float(n_cov)
        if ratios:
            rv['ratios'] = ratios
            rv['dominant_allele_distr_mean'] = np.mean(ratios)
            rv['dominant_allele_distr_sd'] = np.std(ratios)
            for pct in (10, 20, 30, 40, 50, 60, 70, 80, 90, 95, 98, 99):
                rv['dominant_allele_distr_perc_' + str(pct)] = np.percentile(ratios, pct)
        return rv

    def breadth_and_depth_of_coverage(self, mincov=10, minqual=30, trunc=0):
        """cmseq/cmseq.py:459-490.  Upstream calls pileup() WITHOUT min_base_quality, so pysam's default 13 drops reads first
        and the loop then asks for quality >= minqual: the effective threshold is max(13, minqual); N bases and deletions do
        not count; no tag filter."""
        if mincov < 1:
            raise ValueError("mincov must be >= 1")
        # window of the contig that counts: `trunc` columns dropped at both ends when the contig is long enough for that
        lo, hi = (int(trunc), int(self.length - trunc)) if self.length > trunc * 2 else (0, int(self.length))
        counts = self._counts(max(PYSAM_DEFAULT_MIN_BASE_QUALITY, int(minqual)))
        window = counts[lo:hi, :4].astype(np.int64).sum(axis=1)
        hit = np.nonzero(window >= mincov)[0]
        if hit.size == 0:
            return (np.nan, np.nan, np.nan, [np.nan])
        depth_at = dict(zip((hit + lo).tolist(), window[hit].tolist()))   # {0-based column: depth}, ascending like upstream's dict
        vals = list(depth_at.values())
        return (float(hit.size) / (hi - lo), np.mean(vals), np.median(vals), depth_at.values())

    def depth_of_coverage(self, mincov=10, minqual=30):
        return self.breadth_and_depth_of_coverage(mincov, minqual)[1]

    def breadth_of_coverage(self, mincov=10, minqual=30):
        return self.breadth_and_depth_of_coverage(mincov, minqual)[0]

    def get_all_base_values(self, stats_value, *f_args, **f_kwargs):
        """cmseq/cmseq.py:572-578."""
        base_stats = self.get_base_stats(*f_args, **f_kwargs)
        return [base_stats[k].get(stats_value, 'NaN') for k in base_stats]


def _wanted_contigs(spec):
    """cmseq/cmseq.py:58-66: `filterInputList` is a list of contig names, the path of a FASTA file (its record ids: first word of every
    '>' line), or a comma-separated string.  None = no filter."""
    if spec is None:
        return None
    if isinstance(spec, list):
        return set(spec)
    if os.path.isfile(spec):
        with open(spec, "r") as fh:
            heads = (ln[1:].split() for ln in fh if ln.startswith(">"))
            return {h[0] if h else "" for h in heads}
    return set(spec.split(","))


class BamFile:
    """cmseq/cmseq.py:35-88.  `sort` / `index` are accepted and need nothing: the unpacker puts the records in
    `samtools sort` order in memory and the pileup needs no `.bai` (no `.sorted` / `.bai` files are written)."""
    bam_handle = None
    bamFile = None
    contigs = {}

    def __init__(self, bamFile, sort=False, index=False, stepper='nofilter', minlen=CMSEQ_DEFAULTS.minlen, filterInputList=None,
                 minimumReadsAligning=None, ctx: Optional[native.Context] = None, device: int = 0, minqual: int = CMSEQ_DEFAULTS.minqual):
        if not os.path.isfile(bamFile):
            raise Exception(bamFile + ' is not accessible, or is not a file')
        self.bamFile = bamFile
        self.bam_handle = self
        self._own_ctx = ctx is None
        self.ctx = ctx if ctx is not None else native.Context(device)
        self._soas: Dict[int, object] = {}
        first = self._soa(int(minqual))  # header + the streams of the most likely threshold
        self.references = tuple(first.ref_names)
        self.lengths = tuple(int(x) for x in first.ref_lens)
        self._tid = {r: i for i, r in enumerate(self.references)}
        keep = _wanted_contigs(filterInputList)  # upstream scans a list per reference (O(n_ref x n)); a set gives the same answer
        n_on = np.bincount(np.asarray(first.tid, dtype=np.int64), minlength=len(self.references)) if minimumReadsAligning else None
        self.contigs = dict((r, BamContig(self, r, l, stepper)) for i, (r, l) in enumerate(zip(self.references, self.lengths))
                            if l > minlen and (keep is None or r in keep) and (not minimumReadsAligning or int(n_on[i]) >= minimumReadsAligning))

    def _soa(self, minqual: int):
        s = self._soas.get(minqual)
        if s is None:
            s = self._soas[minqual] = bam_mod.unpack_bam(self.bamFile, minqual=minqual, want_qhash=False, lenient_tags=True)
        return s

    def get_contigs(self):
        return iter(self.contigs.keys())

    def get_contigs_obj(self):
        return iter(self.contigs.values())

    def get_contig_by_label(self, contigID):
        return (self.contigs[contigID] if contigID in self.contigs else None)

    def close(self):
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
        self.ctx = None
        self._soas.clear()
