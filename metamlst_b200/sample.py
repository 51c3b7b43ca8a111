"""Whole-sample and cohort driver over the four seams: everything `metamlst.py` does between
opening the BAM and closing the DB (metamlst.py:85-299), with the per-record / per-base work on the GPU.

    typer = SampleTyper("db.sqlite", device=0)            # DB connection + context stay open across samples
    res   = typer.type_bam("sample.bam", out_dir)         # appends out_dir/sample.nfo, optional .out log
    type_cohort(bams, "db.sqlite", out_dir, devices=[0, 1, ...])   # the implicit `for bam in cohort: metamlst.py bam`

Output files are byte-identical to the reference's: the `.nfo` line (metamlst.py:284-285, append mode, '\\r\\n'), the `.out`
log (metamlst.py:160-172) and -- for callers that want the screen output too -- the stdout text including the colour
escapes and the coverage column (metamlst.py:176-296), returned as a string and never printed here.  What the reference
computes per record or per base runs in libmmlst (score, coverage dedupe, pileup, consensus); what it computes once per
allele or locus in Python floats and strings (round, str(round(..,4)*100), ljust/rjust) is done the same way here so the
text is identical by construction (H6).

The cohort form overlaps the three phases of consecutive samples: BAM unpack (C++ threads, GIL released) of sample s+1,
GPU work of sample s, file writing of sample s-1; with several devices, samples are dealt round-robin to one worker
thread per device (SURVEY.md 8e "Cohort: samples independent, replicas").
"""
from __future__ import annotations

import os
import queue
import sqlite3
import threading
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import api, bam, native

# terminal colours of the reference (metaMLST_functions.py: class bcolors)
HEADER, OKBLUE, OKGREEN, WARNING, FAIL, ENDC = "\033[95m", "\033[94m", "\033[92m", "\033[93m", "\033[91m", "\033[0m"


def _status_line(mesg: str, label: str, colour: str) -> str:
    """metamlst_print(mesg, label, type) for messages under 65 characters (metaMLST_functions.py:122-128)."""
    assert len(mesg) < 65
    return mesg.ljust(66) + (colour + "[ - " + label.center(5) + " - ]" + ENDC).ljust(14) + "\r\n"


@dataclass
class SampleResult:
    sample: str
    nfo_lines: List[str] = field(default_factory=list)   # one per organism that passed every gate, in `cel` order
    out_log: Optional[str] = None                        # text of the .out file (when log=True)
    stdout: str = ""                                     # what metamlst.py would have printed (quiet=False)
    cel: Optional[dict] = None
    total_reads: int = 0
    ignored_reads: int = 0
    broken_db: bool = False                              # the reference's sys.exit(0) at metamlst.py:190
    seconds: Dict[str, float] = field(default_factory=dict)


class SampleTyper:
    """One SQLite connection + one GPU context, reused for every sample typed through it."""

    def __init__(self, db_path: str, device: int = 0, minscore: int = 80, max_xM: int = 5, min_read_len: int = 50,
                 min_accuracy: float = 0.90, penalty: int = 100, nloci: int = 100, species_filter: Optional[str] = None,
                 write_known: bool = False, log: bool = False, presorted: bool = False, debug: bool = False,
                 unpack_threads: int = 0, ctx: Optional[native.Context] = None):
        if not os.path.isfile(db_path):
            raise IOError("Failed to connect to the database: please check your database file!")  # metamlst.py:72-73
        self.db_path = db_path
        self.conn = sqlite3.connect(db_path, check_same_thread=False)
        self.conn.row_factory = sqlite3.Row
        self.ctx = ctx if ctx is not None else native.Context(device)
        self._own_ctx = ctx is None
        self.minscore, self.max_xM, self.min_read_len = int(minscore), int(max_xM), int(min_read_len)
        self.min_accuracy, self.penalty, self.nloci = float(min_accuracy), int(penalty), int(nloci)
        self.species_filter, self.write_known, self.log, self.presorted, self.debug = species_filter, write_known, log, presorted, debug
        self.unpack_threads = int(unpack_threads)

    def close(self):
        self.conn.close()
        if self._own_ctx:
            self.ctx.close()

    # -- the SQL the reference runs on this path, unchanged -------------------------------------------------------
    def _genes(self, bacterium: str) -> List[str]:  # metamlst.py:184
        return [r["geneName"] for r in self.conn.execute("SELECT geneName FROM genes WHERE bacterium = ?", (bacterium,))]

    def _longest_allele(self, bacterium: str, gene: str) -> int:  # metamlst.py:225-226
        return self.conn.execute("SELECT LENGTH(sequence) as L FROM alleles WHERE bacterium = ? AND gene = ? ORDER BY L DESC LIMIT 1",
                                 (bacterium, gene)).fetchone()["L"]

    def _unal_sequence(self, bacterium: str, gene: str, allele: str):  # metaMLST_functions.py:186-194
        row = self.conn.execute("SELECT sequence FROM alleles WHERE bacterium = ? AND gene = ? AND alleleVariant = ?",
                                (bacterium, gene, allele)).fetchone()
        return row["sequence"] if row is not None else None

    def _sequence_find(self, bacterium: str, sequence: str):  # metaMLST_functions.py:196-203
        row = self.conn.execute("SELECT gene,alleleVariant FROM alleles WHERE sequence = ? AND bacterium = ?", (str(sequence), bacterium)).fetchone()
        return row["gene"] if row else 0

    # ---------------------------------------------------------------------------------------------------------------
    def unpack(self, bam_path: str):
        """Phase 1 (host, C++ threads): BAM -> packed streams.  Safe to call from another thread than type_unpacked."""
        return bam.unpack_bam(bam_path, presorted=self.presorted, threads=self.unpack_threads)

    def type_bam(self, bam_path: str, out_dir: str, want_stdout: bool = False, timestamp: Optional[int] = None) -> SampleResult:
        t0 = time.perf_counter()
        soa = self.unpack(bam_path)
        res = self.type_unpacked(soa, bam_path, want_stdout=want_stdout)
        res.seconds["unpack"] = res.seconds.pop("_t_start") - t0 if "_t_start" in res.seconds else 0.0
        self.write(res, bam_path, out_dir, timestamp)
        return res

    def type_unpacked(self, soa, bam_path: str, want_stdout: bool = False) -> SampleResult:
        """Phase 2 (GPU + per-locus host arithmetic): metamlst.py:101-296 on an unpacked sample."""
        t_start = time.perf_counter()
        fileName = bam_path.split("/")[-1].split(".")[0]  # metamlst.py:89
        res = SampleResult(sample=fileName)
        index = api.AlleleIndex(soa.ref_names)
        cel, res.total_reads, res.ignored_reads, _raw = api.score_soa(self.ctx, soa, index, self.minscore, self.max_xM, self.min_read_len,
                                                                      self.species_filter, self.penalty)
        res.cel = cel
        t_score = time.perf_counter()
        if self.log:  # metamlst.py:160-172
            head = ("SAMPLE:\t\t\t\t\t" + bam_path + "\r\n" + "VERSION:\t\t\t\t\t1.1\r\n" + "PENALTY:\t\t\t\t" + repr(self.penalty) + "\r\n" +
                    "MIN-THRESHOLD SCORE:\t\t\t\t" + repr(self.minscore) + "\r\n" + "TOTAL ALIGNED READS:\t\t\t\t" + repr(res.total_reads) + "\r\n" +
                    " - OF WHICH IGNORED:\t\t\t\t" + repr(res.ignored_reads) + " BAM READS\r\n\r\n------------------------------  RESULTS ------------------------------\r\n")
            rows = []
            for speciesKey, species in cel.items():
                for geneKey, geneInfo in species.items():
                    for alleleKey, (score, geneLen, average) in sorted(geneInfo.items(), key=lambda x: x[1]):
                        rows.append("\t".join(map(str, [speciesKey, geneKey, alleleKey, score, geneLen, average])) + "\r\n")
            res.out_log = head + "".join(rows)
        out: List[str] = []
        if want_stdout:
            out.append(OKBLUE + "Sample file: " + ENDC + os.path.realpath(fileName) + "\n")
            out.append(OKBLUE + "MetaMLST Database file: " + ENDC + os.path.basename(self.db_path) + "\n\n")
        coverage = None
        for speciesKey, species in cel.items():
            tVar = dict((g, 0) for g in self._genes(speciesKey))
            if len(tVar) < len(species.keys()):  # metamlst.py:188-190: message, then sys.exit(0) -- the sample ends here
                if want_stdout:
                    out.append("Database is broken for" + speciesKey + FAIL + "[ - EXITING - ]".rjust(75, " ") + ENDC + "\n")
                res.broken_db = True
                break
            for sk in species.keys():
                tVar[sk] = 1
            vals = sum(tVar.values())
            passed = int((float(vals) / float(len(tVar))) * 100) >= self.nloci
            if want_stdout:
                out.append((OKGREEN if passed else FAIL) + " " + speciesKey.ljust(18, " ") + ENDC + " Detected Loci: " +
                           ", ".join(OKGREEN + sk + ENDC for sk, v in sorted(tVar.items(), key=lambda x: x[0]) if v == 1) + "\n")
                if any(v == 0 for v in tVar.values()):
                    out.append((" " * 20) + "Missing Loci : " + ", ".join(FAIL + sk + ENDC for sk, v in sorted(tVar.items(), key=lambda x: x[0]) if v == 0) + "\n")
                out.append("\n")
            if not passed:
                continue
            if want_stdout:  # closest-allele table with the coverage column (metamlst.py:206-231)
                if coverage is None:
                    coverage = api.coverage_sums(self.ctx, soa, index, self.minscore, self.max_xM, self.min_read_len, self.species_filter,
                                                 stream_resident=True)
                out.append(_status_line("Closest allele identification", "...", HEADER))
                out.append("\r\n  " + "Locus".ljust(7) + "Avg. Coverage".rjust(15) + "Score".rjust(7) + "Hits".rjust(6) + " Reference Allele(s)".ljust(36) + "\n")
                for geneKey, geneInfo in sorted(species.items(), key=lambda x: x[0]):
                    top = max(avg for (_v, _l, avg) in geneInfo.values())
                    best = dict((k, v) for k, v in geneInfo.items() if v[2] == top)
                    close = ",".join(str(a) for a in sorted(best.keys(), key=lambda x: int(x))[:5]) + ("... (" + str(len(best)) + " more)" if len(best) > 5 else "")
                    genL = self._longest_allele(speciesKey, geneKey)
                    cov = coverage[speciesKey + "_" + geneKey]
                    out.append("  " + WARNING + geneKey.ljust(7) + ENDC + str(round(float(cov) / float(genL), 2)).rjust(15) + ENDC + HEADER + str(top).rjust(7) +
                               str(list(best.values())[0][1]).rjust(6) + ENDC + OKBLUE + " " + close.ljust(36) + ENDC + "\n")
                out.append("\n")
                out.append(_status_line("Building Consensus Sequences", "...", HEADER))
            # chosen allele per locus and its DB sequence (metamlst.py:244), consensus on the GPU (seam S2)
            chromosomeList = {}
            for g, a in api.select_alleles(species):
                seq = self._unal_sequence(speciesKey, g, a)
                if seq is None and want_stdout:
                    out.append(_status_line(" > " + speciesKey + "_" + g + "_" + a + " was not found in the database!", "!!!", WARNING))
                chromosomeList[speciesKey + "_" + g + "_" + a] = seq
            consenSeq = api.build_consensus(self.ctx, soa, chromosomeList, self.minscore, self.max_xM, self.debug)
            finWrite = 1
            if want_stdout:
                out.append("\r")  # buildConsensus ends with print('\r', end='') (metaMLST_functions.py:278)
                out.append("\r\n  " + "Locus".ljust(7) + "Ref.".ljust(7) + "Length".rjust(7) + "Ns".rjust(7) + "SNPs".rjust(7) + "Confidence".rjust(15) + "Notes".rjust(10) + "\n")
            for l in sorted(consenSeq, key=lambda x: x.id):  # metamlst.py:253-276
                holes = str(l.description.split("_")[0].split("::")[1])
                snps = int(l.description.split("_")[1].split("::")[1])
                leng = str(len(l.seq))
                leng_ns = str(round(1 - float(holes) / float(leng), 4) * 100) + " %"
                l.seqLen = len(l.seq)
                if (1 - float(holes) / float(leng)) <= self.min_accuracy:
                    finWrite = 0
                if snps > 0:
                    seqFind = self._sequence_find(speciesKey, l.seq)
                    newAllele = seqFind if seqFind else "NEW"
                else:
                    newAllele = "--"
                    if not self.write_known:
                        l.seq = ""
                if want_stdout:
                    out.append("  " + WARNING + (l.id.split("_")[1]).ljust(7) + ENDC + (l.id.split("_")[2]).ljust(7) + leng.rjust(7) + holes.rjust(7) +
                               str(snps).rjust(7) + leng_ns.rjust(15) + newAllele.rjust(10) + "\n")
            if want_stdout:
                out.append("\n")
            if finWrite:  # metamlst.py:281-287
                if want_stdout:
                    out.append(_status_line("Reconstruction Successful", "WRITE", OKGREEN))
                res.nfo_lines.append(speciesKey + "\t" + fileName + "\t" + "\t".join(
                    recd.id + "::" + str(recd.seq) + "::" + str(round(1 - float(recd.description.split("_")[0].split("::")[1]) / float(recd.seqLen), 4) * 100) +
                    "::" + str(round(float(recd.description.split("_")[1].split("::")[1]) / float(recd.seqLen), 4) * 100) for recd in consenSeq) + "\r\n")
            elif want_stdout:
                out.append(_status_line("Accuracy lower than " + str(round(self.min_accuracy * 100, 2)) + "%", "SKIP", FAIL))
        if want_stdout and len(cel) and not res.broken_db:
            out.append("\033[92m" + "[ - Completed - ]".rjust(80, " ") + "\033[0m" + "\n")
        res.stdout = "".join(out)
        t_end = time.perf_counter()
        res.seconds.update({"score": t_score - t_start, "consensus_and_format": t_end - t_score, "_t_start": t_start})
        return res

    def write(self, res: SampleResult, bam_path: str, out_dir: str, timestamp: Optional[int] = None) -> None:
        """Phase 3: the files metamlst.py leaves in its -o folder."""
        if not os.path.isdir(out_dir):
            os.mkdir(out_dir)  # metamlst.py:91
        if res.out_log is not None:
            ts = int(time.time()) if timestamp is None else int(timestamp)
            with open(out_dir + "/" + res.sample + "_" + str(ts) + ".out", "w", newline="") as f:
                f.write(res.out_log)
        for line in res.nfo_lines:
            with open(out_dir + "/" + res.sample + ".nfo", "a", newline="") as f:  # append mode, one line per organism
                f.write(line)
        res.seconds.pop("_t_start", None)


def type_cohort(bam_paths: Sequence[str], db_path: str, out_dir: str, devices: Sequence[int] = (0,), prefetch: int = 2,
                **params) -> List[SampleResult]:
    """Type a cohort back to back (BASELINE.json configs[3]): per device one worker thread owning a SampleTyper; an unpack
    thread per device runs `prefetch` samples ahead; files are written by the worker after the GPU phase of the NEXT sample
    has been queued.  Samples are dealt round-robin to the devices; results come back in input order."""
    results: List[Optional[SampleResult]] = [None] * len(bam_paths)
    errors: List[BaseException] = []

    def worker(slot: int, device: int):
        mine = [i for i in range(len(bam_paths)) if i % len(devices) == slot]
        if not mine:
            return
        typer = SampleTyper(db_path, device=device, **params)
        q: "queue.Queue" = queue.Queue(maxsize=max(1, prefetch))

        def unpacker():
            for i in mine:
                try:
                    t0 = time.perf_counter()
                    soa = typer.unpack(bam_paths[i])
                    q.put((i, soa, time.perf_counter() - t0, None))
                except BaseException as e:  # noqa: BLE001
                    q.put((i, None, 0.0, e))
                    return
            q.put(None)

        th = threading.Thread(target=unpacker, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                i, soa, t_unpack, err = item
                if err is not None:
                    raise err
                res = typer.type_unpacked(soa, bam_paths[i])
                res.seconds["unpack"] = t_unpack
                del soa
                typer.write(res, bam_paths[i], out_dir)
                results[i] = res
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
        finally:
            typer.close()

    threads = [threading.Thread(target=worker, args=(s, d)) for s, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return [r for r in results if r is not None]
