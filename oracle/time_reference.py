#!/usr/bin/env python3
"""TEST / MEASUREMENT INFRASTRUCTURE -- wall time of the UNMODIFIED reference (metamlst.py + cmseq) over the shims.

    python oracle/time_reference.py [--reads 20000] [--read-len 100] [--alleles 256]     # needs /root/reference

SURVEY.md 8d asks for the reference's own Python timed beside the GPU path on configs[0] (7-locus E. coli scheme, 100 bp reads,
K = 4 records per read), split into (i) time inside the shims (BAM decode + pileup engine: stands in for htslib/samtools C
code and is NOT representative of their speed) and (ii) time inside the reference's own Python lines (representative).
/root/reference exists only in the build container, so this runs here, single process / single thread as the reference is,
and the result is recorded in DESIGN.md; the GPU box times the C port of the oracle instead (`bench.py --impl reference`).
"""
import argparse
import cProfile
import json
import os
import pstats
import runpy
import shutil
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.environ.get("MMLST_REFERENCE", "/root/reference")


def inner(bam, db, out):
    """Runs with the shims first on sys.path: executes metamlst.py in-process under cProfile."""
    sys.argv = ["metamlst.py", bam, "-d", db, "-o", out, "--quiet"]
    if os.environ.get("MMLST_REF_PLAIN") == "1":  # plain wall time, no profiler
        t0 = time.perf_counter()
        try:
            runpy.run_path(os.path.join(REF, "metamlst.py"), run_name="__main__")
        except SystemExit:
            pass
        print(json.dumps({"wall_s": time.perf_counter() - t0}))
        return
    prof = cProfile.Profile()
    t0 = time.perf_counter()
    try:
        prof.runcall(runpy.run_path, os.path.join(REF, "metamlst.py"), run_name="__main__")
    except SystemExit:
        pass
    wall = time.perf_counter() - t0
    st = pstats.Stats(prof)
    ref_t = shim_t = other_t = 0.0
    for (fn, _ln, _name), (_cc, _nc, tt, _ct, _callers) in st.stats.items():  # tt = time inside the function itself
        if fn.startswith(REF):
            ref_t += tt
        elif fn.startswith(os.path.join(ROOT, "oracle")):
            shim_t += tt
        else:
            other_t += tt
    print(json.dumps({"wall_s": wall, "in_reference_lines_s": ref_t, "in_shims_s": shim_t, "elsewhere_s": other_t}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--alleles", type=int, default=256)
    ap.add_argument("--inner", nargs=3)
    a = ap.parse_args()
    if a.inner:
        inner(*a.inner)
        return
    sys.path.insert(0, ROOT)
    import subprocess
    from metamlst_b200 import synth
    from oracle import bamio
    work = tempfile.mkdtemp(prefix="mmlst_ref_")
    try:
        db = synth.make_db(("ecoli",), alleles_per_locus=a.alleles, n_profiles=2048, seed=1001)
        dbp = os.path.join(work, "db.sqlite")
        db.write_sqlite(dbp)
        tab = synth.make_sample(db, a.reads, a.read_len, seed=1001, K=4)
        bam = os.path.join(work, "sample.bam")
        bamio.write_table_bam(bam, tab)
        e = dict(os.environ)
        e["PYTHONPATH"] = os.path.join(ROOT, "oracle", "shims") + os.pathsep + REF + os.pathsep + ROOT
        e["PATH"] = os.path.join(ROOT, "oracle", "shims", "bin") + os.pathsep + e["PATH"]
        def run_inner(plain, tag):
            bam2 = os.path.join(work, tag + ".bam")
            shutil.copy(bam, bam2)  # metamlst.py sorts its input in place
            e2 = dict(e, MMLST_REF_PLAIN="1" if plain else "0")
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--inner", bam2, dbp, os.path.join(work, "out_" + tag)], env=e2, cwd=work,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            line = [l for l in p.stdout.decode().splitlines() if l.startswith("{")]
            assert p.returncode == 0 and line, p.stderr.decode()[-2000:]
            return json.loads(line[-1])
        plain = run_inner(True, "plain")
        r = run_inner(False, "prof")
        r["wall_profiled_s"] = r.pop("wall_s")
        r["wall_s"] = plain["wall_s"]
        nfo = os.path.join(work, "out_plain", "plain.nfo")
        r.update({"records": int(tab.n), "reads": a.reads, "read_len": a.read_len, "alleles_per_locus": a.alleles, "loci": 7,
                  "records_per_s": tab.n / r["wall_s"], "fraction_in_reference_lines": r["in_reference_lines_s"] / max(r["in_reference_lines_s"] + r["in_shims_s"] + r["elsewhere_s"], 1e-9),
                  "nfo_written": os.path.exists(nfo), "cores": 1, "note": "wall_s is the plain run; the split comes from a second run under cProfile (wall_profiled_s)"})
        print(json.dumps(r))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
