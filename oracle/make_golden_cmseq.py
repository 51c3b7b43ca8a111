#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- golden vectors for the cmseq seam (metamlst_b200/cmseq_api.py).

    python oracle/make_golden_cmseq.py          # needs /root/reference (build container only)

Runs the UNMODIFIED /root/reference/cmseq/cmseq.py (BamFile / BamContig) over the pysam + Bio + samtools shims of
oracle/shims on the committed golden BAMs (tests/golden/<scenario>/sample.bam, coordinate-sorted copies) and stores what
its methods return as tests/golden/cmseq_api.json.gz.  The pileup ENGINE under the shim is the oracle's restatement of
htslib (unpinned residue, see oracle/mlst_oracle.py); every line of cmseq itself is the reference's own.
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.environ.get("MMLST_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
SCENARIOS = ("basic", "strict", "two_org", "no_xs", "lowcov", "deep")
MLST_FILTER = [["AS", "loc_gte", 80], ["XM", "loc_lte", 5]]  # metaMLST_functions.py:259 with the CLI defaults

# (method, kwargs) evaluated on every chosen contig; tuples in tag filters are rebuilt on both sides
CASES = [
    ("get_base_stats", {}),
    ("get_base_stats", {"min_read_depth": 1, "min_base_quality": 20, "dominant_frq_thrsh": 0.4, "BAM_tagFilter": MLST_FILTER}),
    ("get_base_stats", {"min_read_depth": 5, "min_base_quality": 30, "dominant_frq_thrsh": 0.95, "error_rate": 0.01}),
    ("get_base_stats", {"min_read_depth": 2, "min_base_quality": 0, "BAM_tagFilter": [["XM", "loc_lte", 2]]}),
    ("reference_free_consensus", {}),
    ("reference_free_consensus", {"mincov": 1, "minqual": 20, "dominant_frq_thrsh": 0.4, "noneCharacter": "N", "BAM_tagFilter": MLST_FILTER}),
    ("reference_free_consensus", {"consensus_rule": "majority_rule_polymorphicLoci", "mincov": 3, "minqual": 20, "dominant_frq_thrsh": 0.99}),
    ("polymorphism_rate", {}),
    ("polymorphism_rate", {"mincov": 2, "minqual": 20, "dominant_frq_thrsh": 0.995, "pvalue": 0.9}),
    ("breadth_and_depth_of_coverage", {}),
    ("breadth_and_depth_of_coverage", {"mincov": 3, "minqual": 30, "trunc": 10}),
    ("breadth_and_depth_of_coverage", {"mincov": 1, "minqual": 5, "trunc": 400}),
    ("depth_of_coverage", {"mincov": 2, "minqual": 20}),
    ("breadth_of_coverage", {"mincov": 2, "minqual": 20}),
    ("get_all_base_values", {"stats_value": "ratio_max2all", "min_base_quality": 25}),
]


def plain(x):
    """JSON-able copy: numpy scalars -> Python, dict views -> lists, int keys -> str (json does that anyway)."""
    import numpy as np
    if isinstance(x, dict):
        return {str(k): plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)) or type(x).__name__ in ("dict_values", "dict_keys"):
        return [plain(v) for v in x]
    if isinstance(x, np.generic):
        return x.item()
    return x


def inner():
    """Runs with the shims first on sys.path: imports the reference's cmseq and evaluates CASES."""
    import numpy as np
    from cmseq import cmseq  # namespace package from /root/reference
    assert cmseq.__file__.startswith(REF), cmseq.__file__
    import pysam
    assert "shim" in pysam.__version__
    out = {}
    for scen in SCENARIOS:
        src = os.path.join(GOLD, scen, "sample.bam")
        with tempfile.TemporaryDirectory() as td:
            bam = os.path.join(td, "sorted.bam")
            subprocess.check_call(["samtools", "sort", src, "-o", bam])
            af = pysam.AlignmentFile(bam, "rb")
            per = sorted(((af.count(contig=r), r) for r in af.references), key=lambda t: (-t[0], t[1]))
            chosen = [r for _n, r in per[:3]] + [per[len(per) // 2][1], per[-1][1]]
            chosen = list(dict.fromkeys(chosen))
            bf = cmseq.BamFile(bam, filterInputList=list(chosen))
            res = {"contigs": chosen, "n_references": len(af.references), "kept_minreads_50": sorted(
                cmseq.BamFile(bam, minimumReadsAligning=50, minlen=100).contigs.keys()), "cases": []}
            for meth, kw in CASES:
                for c in chosen:
                    k2 = dict(kw)
                    if "BAM_tagFilter" in k2:
                        k2["BAM_tagFilter"] = [tuple(e) for e in k2["BAM_tagFilter"]]
                    if "consensus_rule" in k2:
                        k2["consensus_rule"] = getattr(cmseq.BamContig, k2["consensus_rule"])
                    contig = bf.get_contig_by_label(c)
                    if "stats_value" in k2:
                        sv = k2.pop("stats_value")
                        val = contig.get_all_base_values(sv, **k2)
                    else:
                        val = getattr(contig, meth)(**k2)
                    res["cases"].append({"method": meth, "kwargs": kw, "contig": c, "result": plain(val)})
            out[scen] = res
    import gzip
    with gzip.GzipFile(os.path.join(GOLD, "cmseq_api.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", os.path.join(GOLD, "cmseq_api.json.gz"), {s: len(v["cases"]) for s, v in out.items()})


def main():
    if os.environ.get("MMLST_GOLDEN_INNER") == "1":
        inner()
        return
    e = dict(os.environ)
    e["PYTHONPATH"] = os.path.join(ROOT, "oracle", "shims") + os.pathsep + REF + os.pathsep + ROOT
    e["PATH"] = os.path.join(ROOT, "oracle", "shims", "bin") + os.pathsep + e["PATH"]
    e["MMLST_GOLDEN_INNER"] = "1"
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=e))


if __name__ == "__main__":
    main()
