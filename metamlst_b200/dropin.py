"""Drop-in names for the reference's scripts: `metamlst.py` / `metamlst-merge.py` keep their command line, SQL and writers and
differ from upstream only in what four names resolve to (INTEGRATION.md shows the three-line diff).

    buildConsensus(bamFile, chromosomeList, filterScore, max_xM, debugMode)    metaMLST_functions.py:249 -- same signature
    score_bam(bam_path, minscore, max_xM, min_read_len, species_filter)       replaces the inline loop metamlst.py:96-151
    sort_index(bamFile)                                                        metaMLST_functions.py:237 -- nothing to do: the unpacker
                                                                               orders the records in memory and never rewrites the file
    stringDiff(s1, s2), closest_allele(...), defineProfile(conn, geneList)

A BAM is unpacked ONCE per (path, mtime, size) -- stage 1 and stage 2 of the same run share it -- by the native unpacker
(bam.unpack_bam); every per-record / per-base step then runs in libmmlst on the device named by MMLST_DEVICE (default 0).
There is no CPU fallback: without the library or a GPU every call raises.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

from . import api, bam, native

_ctx: Optional[native.Context] = None
_cache: "OrderedDict[tuple, object]" = OrderedDict()
_CACHE_SAMPLES = 2


def context() -> native.Context:
    global _ctx
    if _ctx is None:
        _ctx = native.Context(int(os.environ.get("MMLST_DEVICE", "0")))
    return _ctx


def unpacked(bam_path: str, presorted: bool = False):
    """The sample's packed streams (SoaHost), unpacked on first use and kept for the next seam of the same run."""
    st = os.stat(bam_path)
    key = (os.path.realpath(bam_path), st.st_mtime_ns, st.st_size, bool(presorted))
    soa = _cache.get(key)
    if soa is None:
        soa = bam.unpack_bam(bam_path, presorted=presorted, want_qhash=True)
        _cache[key] = soa
        while len(_cache) > _CACHE_SAMPLES:
            _cache.popitem(last=False)
    else:
        _cache.move_to_end(key)
    return soa


def score_bam(bam_path: str, minscore: int, max_xM: int, min_read_len: int, species_filter: Optional[str] = None, penalty: int = 100,
              presorted: bool = False):
    """Seam S1 (metamlst.py:96-151).  Returns (cel, sequenceBankSums, totalReads, ignoredReads): cel[species][gene][allele] =
    (localScore, n, round(avg, 1)) in first-appearance order (H5); sequenceBankSums['species_gene'] = sum over unique read names
    of len(SEQ) (what `sum(sequenceBank[key].values())` gives at metamlst.py:228, H7)."""
    soa = unpacked(bam_path, presorted)
    index = api.AlleleIndex(soa.ref_names)
    cel, total, ignored, _raw = api.score_soa(context(), soa, index, minscore, max_xM, min_read_len, species_filter, penalty)
    bank = api.coverage_sums(context(), soa, index, minscore, max_xM, min_read_len, species_filter, stream_resident=True)
    return cel, bank, total, ignored


def buildConsensus(bamFile, chromosomeList, filterScore, max_xM, debugMode):  # noqa: N802,N803 -- the reference's names
    """Seam S2 (metaMLST_functions.py:249-281), same signature and return shape: a list of records with .id, .seq (mutable),
    .description == 'CI::<holes>_SP::<snps>' in chromosomeList order."""
    return api.build_consensus(context(), unpacked(bamFile), chromosomeList, filterScore, max_xM, debugMode)


def sort_index(bamFile):  # noqa: N803
    """metaMLST_functions.py:237-247 rewrote the input with `samtools sort` + `samtools index` for pysam's sake.  The unpacker puts
    the records in that order in memory; the input file is left alone."""
    return None


def stringDiff(s1, s2):  # noqa: N802
    """metaMLST_functions.py:230-234 -- one pair through the Hamming kernel (zip truncation, character compare: H9)."""
    return api.stringDiff(s1, s2, context())


_hamming: Dict[Tuple[str, int], api.HammingIndex] = {}


def closest_allele(conn, bacterium: str, gene: str, seq: str) -> Tuple[int, int]:
    """Seam S3: (min stringDiff over sequencesGetAll(conn, bacterium, gene), alleleVariant of the first row reaching it); the
    reference's `any(stringDiff(..) <= z)` (metamlst-merge.py:177-181) is `closest_allele(..)[0] <= z`."""
    key = (bacterium, id(conn))
    idx = _hamming.get(key)
    if idx is None:
        idx = _hamming[key] = api.HammingIndex.from_sqlite(context(), conn, bacterium)
    return idx.closest_allele(bacterium, gene, seq)


_profiles: Dict[int, api.ProfileIndex] = {}


def defineProfile(conn, geneList):  # noqa: N802,N803
    """Seam S4 (metaMLST_functions.py:205-216, H11 included) on the device."""
    idx = _profiles.get(id(conn))
    if idx is None:
        idx = _profiles[id(conn)] = api.ProfileIndex(context(), conn)
    if not geneList:
        return api.define_profile(conn, geneList)  # the reference's own failure mode for an empty list
    return idx.define_profiles([list(geneList)])[0]
