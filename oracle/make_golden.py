#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- generate tests/golden/ by running the UNMODIFIED reference scripts over the shims.

    python oracle/make_golden.py            # needs /root/reference (build container only)

For every scenario: build a seeded synthetic DB + BAM (metamlst_b200.synth), build the sqlite DB twice (directly,
and with /root/reference/metamlst-index.py from FASTA + typings -- the two must be identical), run
/root/reference/metamlst.py and /root/reference/metamlst-merge.py with `oracle/shims` on PYTHONPATH and
`oracle/shims/bin` on PATH, and store inputs + outputs.  /root/reference does not exist on the GPU box, so the
outputs are committed; tests/test_oracle_golden.py pins oracle/mlst_oracle.py against them.
"""
import glob
import json
import os
import shutil
import sqlite3
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
REF = os.environ.get("MMLST_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

from metamlst_b200 import synth  # noqa: E402
from oracle import bamio  # noqa: E402

import numpy as np  # noqa: E402

SHORT = {"ecoli": [("adk", 80), ("fumC", 70)]}

SCENARIOS = {
    # name: (db kwargs, sample kwargs, metamlst.py extra args, order)
    "basic": (dict(organisms=("ecoli",), alleles_per_locus=12, n_profiles=40, seed=11),
              dict(n_reads=1200, read_len=100, seed=11), ["--log"], "name"),
    "strict": (dict(organisms=("ecoli",), alleles_per_locus=12, n_profiles=40, seed=12),
               dict(n_reads=1200, read_len=100, seed=12, sub_err=0.02, n_frac=0.03),
               ["--log", "--minscore", "178", "--max_xM", "3", "-a"], "name"),
    "two_org": (dict(organisms=("ecoli", "saureus"), alleles_per_locus=8, n_profiles=30, seed=13),
                dict(n_reads=2000, read_len=100, seed=13, org_props=(0.6, 0.4)), ["--log"], "name"),
    "filter": (dict(organisms=("ecoli", "saureus"), alleles_per_locus=8, n_profiles=30, seed=13),
               dict(n_reads=2000, read_len=100, seed=13, org_props=(0.6, 0.4)), ["--log", "--filter", "saureus"], "name"),
    "presorted": (dict(organisms=("ecoli",), alleles_per_locus=12, n_profiles=40, seed=14),
                  dict(n_reads=1000, read_len=100, seed=14), ["--presorted", "--log"], "coord"),
    "no_xs": (dict(organisms=("ecoli",), alleles_per_locus=12, n_profiles=40, seed=15),
              dict(n_reads=800, read_len=100, seed=15, frac_indel=0.2), ["--log"], "name"),  # H4
    "lowcov": (dict(organisms=("ecoli",), alleles_per_locus=12, n_profiles=40, seed=16),
               dict(n_reads=25, read_len=100, seed=16), ["--log", "--min_accuracy", "0.2"], "name"),  # holes
    "deep": (dict(organisms=("ecoli",), alleles_per_locus=4, n_profiles=6, seed=17, schemes=SHORT),
             dict(n_reads=24000, read_len=50, seed=17, K=1, frac_clip=0.0, frac_indel=0.0, novel_loci=1),
             ["--presorted", "--log"], "coord"),  # H1: depth > 8000
}


def env():
    e = dict(os.environ)
    e["PYTHONPATH"] = os.path.join(ROOT, "oracle", "shims") + os.pathsep + REF + os.pathsep + ROOT
    e["PATH"] = os.path.join(ROOT, "oracle", "shims", "bin") + os.pathsep + e["PATH"]
    return e


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, env=env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return p.returncode, p.stdout.decode(errors="replace")


def dump_db(path):
    c = sqlite3.connect(path)
    out = {}
    for t, cols in (("organisms", "organismkey,label"), ("genes", "geneName,bacterium"),
                    ("alleles", "recID,bacterium,gene,sequence,alignedSequence,alleleVariant"),
                    ("profiles", "recID,profileCode,bacterium,alleleCode")):
        out[t] = sorted(tuple(r) for r in c.execute("SELECT %s FROM %s" % (cols, t)))
    c.close()
    return out


def build_db_with_reference(db, out_db, work):
    """metamlst-index.py over shims: one -s FASTA, one -t typings file per organism."""
    for oi, o in enumerate(db.organisms):
        sub = synth.SynthDB([o], {o: db.loci[o]}, db.row_org, db.row_locus, db.row_variant, db.seq_off, db.seq,
                            db.locus_names, db.locus_row0, {o: db.profiles[o]})
        fa = os.path.join(work, "all.fa")
        ty = os.path.join(work, "typ_%s.txt" % o)
        # FASTA holds every organism's alleles (written once); typings one organism per file
        with open(ty, "w") as f:
            f.write("#%s|Synthetic %s\n" % (o, o))
            f.write("ST\t" + "\t".join(g for g, _ in db.loci[o]) + "\n")
            for st in range(db.profiles[o].shape[0]):
                f.write("%d\t%s\n" % (st + 1, "\t".join(str(int(v)) for v in db.profiles[o][st])))
        if oi == 0:
            with open(fa, "w") as f:
                for r, name in enumerate(db.ref_names()):
                    f.write(">%s\n%s\n" % (name, db.row_seq(r)))
            open(out_db, "w").close()  # metamlst-index.py only creates tables when the file exists (:60)
            rc, log = run([sys.executable, os.path.join(REF, "metamlst-index.py"), "-d", out_db, "-s", fa], work)
            assert rc == 0, log
        rc, log = run([sys.executable, os.path.join(REF, "metamlst-index.py"), "-d", out_db, "-t", ty], work)
        assert rc == 0, log
    del sub


def main():
    assert os.path.isdir(REF), "reference not mounted"
    os.makedirs(GOLD, exist_ok=True)
    manifest = {}
    for name, (dbkw, skw, extra, order) in SCENARIOS.items():
        d = os.path.join(GOLD, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        work = os.path.join(d, "_work")
        os.makedirs(work)
        db = synth.make_db(**dbkw)
        direct = os.path.join(work, "direct.db")
        db.write_sqlite(direct)
        dbpath = os.path.join(d, "db.sqlite")
        build_db_with_reference(db, dbpath, work)
        assert dump_db(direct) == dump_db(dbpath), "direct sqlite build differs from metamlst-index.py build"
        tab = synth.make_sample(db, **skw)
        if name == "no_xs":
            tab.has_xs = (np.arange(tab.n) % 3) != 0
        if order == "coord":
            tab = tab.sorted_by_coord()
        bam = os.path.join(d, "sample.bam")
        bamio.write_table_bam(bam, tab, "coordinate" if order == "coord" else "unknown")
        run_bam = os.path.join(work, "sample.bam")
        shutil.copy(bam, run_bam)  # metamlst.py sorts its input IN PLACE (metaMLST_functions.py:244-245)
        outdir = os.path.join(work, "out")
        rc, log = run([sys.executable, os.path.join(REF, "metamlst.py"), run_bam, "-d", dbpath, "-o", outdir] + extra, work)
        open(os.path.join(d, "metamlst.stdout"), "w").write(log)
        assert rc == 0, log
        nfo = os.path.join(outdir, "sample.nfo")
        if os.path.exists(nfo):
            shutil.copy(nfo, os.path.join(d, "sample.nfo"))
        for f in glob.glob(os.path.join(outdir, "sample_*.out")):
            shutil.copy(f, os.path.join(d, "sample.out"))
        manifest[name] = {"db": dbkw if "schemes" not in dbkw else {**dbkw, "schemes": "SHORT"}, "sample": skw,
                          "args": extra, "order": order, "records": tab.n, "rc": rc,
                          "truth_st": tab.truth["st"]}
        shutil.rmtree(work)
        print(name, "ok", tab.n, "records")

    # merge scenario: a cohort folder of .nfo files from several samples typed by the reference
    d = os.path.join(GOLD, "cohort")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    work = os.path.join(d, "_work")
    os.makedirs(work)
    db = synth.make_db(organisms=("ecoli",), alleles_per_locus=12, n_profiles=40, seed=21)
    dbpath = os.path.join(d, "db.sqlite")
    db.write_sqlite(dbpath)
    outdir = os.path.join(d, "nfo")
    os.makedirs(outdir)
    # samples: 0,1 same strain seed (recurring new alleles), 2 known ST (no novel loci), 3 far novel allele (> z SNPs)
    specs = [("s0", dict(seed=31)), ("s1", dict(seed=31)), ("s2", dict(seed=32, novel_loci=0)), ("s3", dict(seed=33)),
             ("s4", dict(seed=34, novel_loci=1))]
    for sname, kw in specs:
        tab = synth.make_sample(db, n_reads=1500, read_len=100, **kw)
        if sname == "s1":
            tab2 = synth.make_sample(db, n_reads=1500, read_len=100, seed=31)
            tab = tab2  # identical strain; different file name
        if sname == "s3":
            # push one locus far from every known allele: handled below by editing the .nfo sequence
            pass
        bam = os.path.join(work, sname + ".bam")
        bamio.write_table_bam(bam, tab)
        rc, log = run([sys.executable, os.path.join(REF, "metamlst.py"), bam, "-d", dbpath, "-o", outdir, "--quiet"], work)
        assert rc == 0, log
    # make s3's first non-empty sequence differ at 9 sites from everything (rejected at -z 5)
    p3 = os.path.join(outdir, "s3.nfo")
    line = open(p3, newline="").read()
    parts = line.rstrip("\r\n").split("\t")
    for i in range(2, len(parts)):
        f = parts[i].split("::")
        if f[1]:
            s = list(f[1])
            for k in range(9):
                j = 7 + 11 * k
                s[j] = {"A": "C", "C": "G", "G": "T", "T": "A"}.get(s[j].upper(), "A")
            f[1] = "".join(s)
            parts[i] = "::".join(f)
            break
    open(p3, "w", newline="").write("\t".join(parts) + "\r\n")
    rc, log = run([sys.executable, os.path.join(REF, "metamlst-merge.py"), outdir, "-d", dbpath], work)
    open(os.path.join(d, "merge.stdout"), "w").write(log)
    assert rc == 0, log
    shutil.rmtree(work)
    manifest["cohort"] = {"samples": [s for s, _ in specs], "rc": rc}
    print("cohort ok")
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1, sort_keys=True, default=str)


if __name__ == "__main__":
    main()
