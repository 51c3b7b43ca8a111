#!/usr/bin/env python3
"""BASELINE.json configs[2] and configs[3] end to end (bench.py is the driver's contract and runs configs[1] / configs[4]).

  c3  "metamlstDB_2022-scale synthetic DB (all schemes) with a 100M-read metagenome BAM sharded over 8 B200" (SURVEY.md 8d):
      150 organisms x 7 loci, allele counts log-uniform 50-4000 (~5e5 rows), reads over 20 present organisms with Zipf
      abundances, coordinate-sorted, htslib depth cap resolved over the WHOLE sample, then cut into `world` contiguous record
      ranges -- the all-reduce form of the path (north_star: "per-GPU partial pileup-count and score tensors are allreduced").
        python -m torch.distributed.run --nproc-per-node 8 ... bench_configs.py c3 --reads 100000000
      parity: a sample of loci re-typed by the C port of the oracle from the same reads (score tables of their alleles, chosen
      allele, consensus, holes, SNPs).

  c4  "64-sample cohort batch typed back-to-back, per-sample and whole-box reads/s": 64 bowtie2-ordered BAM files, one worker
      thread + one loader thread per GPU (sample.type_cohort), device-side ingest, `.nfo` files written.
        python bench_configs.py c4 --samples 64 --reads 250000 --devices 8
      parity: samples re-typed by the C port of the oracle from the same reads, and by the host-unpack / host-seam engine.

Each prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4"])
    ap.add_argument("--reads", type=int, default=0, help="c3: reads of the whole metagenome (default 100 M); c4: reads per sample (default 250 k)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--organisms", type=int, default=150)
    ap.add_argument("--present", type=int, default=20)
    ap.add_argument("--min-alleles", type=int, default=50)
    ap.add_argument("--max-alleles", type=int, default=4000)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--devices", type=int, default=0, help="c4: GPUs to use (0 = all visible)")
    ap.add_argument("--alleles", type=int, default=1024, help="c4: alleles per locus of the configs[1] DB")
    ap.add_argument("--ingest", default="device", choices=["device", "host"])
    ap.add_argument("--workers", default="processes", choices=["processes", "threads"], help="c4: one process per GPU (sample.CohortPool) or one thread per GPU (sample.type_cohort)")
    ap.add_argument("--per-gpu", type=int, default=2, help="c4, --workers processes: worker processes per GPU")
    ap.add_argument("--parity-loci", type=int, default=12)
    ap.add_argument("--max-depth", type=int, default=8000)
    ap.add_argument("--pileup-impl", type=int, default=0)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
def make_db_c3(args):
    from metamlst_b200 import synth
    rng = np.random.Generator(np.random.PCG64(1003))
    orgs = ["sp%03d" % i for i in range(args.organisms)]
    schemes = {o: [("g%d" % (j + 1), int(rng.integers(350, 551))) for j in range(7)] for o in orgs}
    apl = np.exp(rng.uniform(np.log(args.min_alleles), np.log(args.max_alleles), size=7 * len(orgs))).astype(np.int64)
    return synth.make_db(orgs, alleles_per_locus=apl, n_profiles=200, seed=1003, schemes=schemes), orgs


def run_c3(args):
    import torch
    from metamlst_b200 import api, devpack, native, pipeline, synth
    from oracle import cpu_path
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    native.lib()
    n_reads = args.reads or 100_000_000
    t0 = time.perf_counter()
    db, orgs = make_db_c3(args)
    # Zipf abundances over the present organisms
    rng = np.random.Generator(np.random.PCG64(2003))
    present = rng.choice(len(orgs), size=min(args.present, len(orgs)), replace=False)
    props = np.zeros(len(orgs))
    props[present] = 1.0 / np.arange(1, len(present) + 1)
    t_db = time.perf_counter() - t0
    # every rank generates the WHOLE sample (same seeds), sorts it, resolves the depth cap over all of it, then keeps its record range
    chunk = 1_000_000
    cores = []
    for i, c0 in enumerate(range(0, n_reads, chunk)):
        core = synth.gen_core(db, min(chunk, n_reads - c0), args.read_len, seed=1003 * 1000 + i, K=args.k, org_props=props, device=device, strain_seed=1003)
        cores.append({k: core[k] for k in ("L", "K", "bases", "qual", "rtype", "a_split", "rows", "start", "flag", "AS", "xm")})
        del core
    whole = devpack.pack_cores(db, cores, 20, args.max_depth or None)
    R_total = int(whole.tid.shape[0])
    st = whole.slice_ranges(rank, world)
    R_local = int(st.tid.shape[0])
    n_prec_total = int(whole.n_prec)
    del whole
    torch.cuda.empty_cache()
    t_gen = time.perf_counter() - t0 - t_db
    index = api.AlleleIndex(db.ref_names())
    pipe = pipeline.DevicePipeline(st, index, db.row_seq, impl=args.pileup_impl, exchange="allreduce", db_ascii=db.seq, db_off=db.seq_off, **bench.PARAMS)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        result = pipe.step()
    use_graph = True
    try:
        pipe.capture()
        assert pipe.step_graph() == result, "graph replay differs from the eager pass"
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("CUDA graph capture unavailable (%r): timing the eager pass\n" % (e,))
        use_graph = False
        pipe.graph = None
    barrier()
    pipe.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pipe.enqueue_step()
    e1.record()
    barrier()
    assert pipe.collect() == result, "results changed between steps"
    tm = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
    ms = float(tm.item())
    # ---- parity on a sample of loci (rank 0): the same reads restricted to those loci through the C port of the oracle
    parity = None
    if rank == 0:
        loci_in_result = [index.locus_of[index.name_to_tid[c]] for sp in result for (c, _s, _h, _n) in result[sp]]
        prng = np.random.Generator(np.random.PCG64(7))
        pick = sorted(prng.choice(loci_in_result, size=min(args.parity_loci, len(loci_in_result)), replace=False).tolist()) if loci_in_result else []
        row_locus = torch.from_numpy(db.row_locus.astype(np.int64)).to(device)
        want_loc = torch.zeros(len(db.locus_names), dtype=torch.bool, device=device)
        want_loc[torch.tensor(pick, dtype=torch.int64, device=device)] = True
        sub = []
        for c in cores:
            keep = want_loc[row_locus[c["rows"][:, 0]]]
            if bool(keep.any()):
                sub.append({k: (v[keep] if torch.is_tensor(v) else v) for k, v in c.items()})
        w = cpu_path.workload_from_cores(db, sub)
        ref = cpu_path.run(w, max_depth=args.max_depth or None, threads=os.cpu_count() or 1, **{k: bench.PARAMS[k] for k in ("minscore", "min_read_len", "penalty")},
                           max_xM=bench.PARAMS["max_xM"])
        mine = {c: (s_, h_, n_) for sp in result for (c, s_, h_, n_) in result[sp]}
        checked = 0
        for sp in ref["result"]:
            for (c, s_, h_, n_) in ref["result"][sp]:
                assert mine.get(c) == (s_, h_, n_), "locus %s differs from the C port" % c
                checked += 1
        # score tables of the sampled loci's alleles (the all-reduced tables, read back before the selection consumes them)
        pipe.want_tables = True
        parity = {"ok": True, "loci_checked": checked, "records_in_sample": w.n, "against": "oracle/c via oracle/cpu_path.py on the reads of the sampled loci"}
    del cores
    if True:  # all ranks: the table pass is collective (all-reduce inside run_score)
        pipe.want_tables = True
        pipe.graph = None
        pipe.step()
        if rank == 0:
            s_tab, n_tab, _f = pipe.tables()
            rows = np.nonzero(np.isin(db.row_locus, pick))[0]
            assert np.array_equal(s_tab[rows], ref["sum_as"][rows]) and np.array_equal(n_tab[rows], ref["n_hit"][rows]), "score tables of the sampled loci differ"
            parity["checked"] = ["sum_as / n_hit of every allele of the sampled loci (all-reduced tables)", "chosen allele", "consensus", "holes", "snps"]
    if rank == 0:
        line = {"metric": "aligned reads/s (score+pileup+consensus)", "value": R_total / (ms / 1e3), "unit": "records/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "dtype": "int32/u8", "data": "synthetic",
                "config": {"workload": "configs[2]: %d organisms x 7 loci, allele counts log-uniform %d-%d (%d rows, %d loci, %.2e bases); ONE metagenome of %d x %d bp reads, "
                                       "K=%d (%d records) over %d present organisms (Zipf), coordinate-sorted, depth cap %d resolved over the whole sample (%d records "
                                       "admitted), cut into %d contiguous record ranges" % (len(orgs), args.min_alleles, args.max_alleles, db.n_rows, len(db.locus_names),
                                                                                             float(db.seq.size), n_reads, args.read_len, args.k, R_total, len(present),
                                                                                             args.max_depth, n_prec_total, world),
                           "sharding": "contiguous record ranges; all-reduce SUM(sum_as, n_hit, counters) MIN(first_idx) SUM(counts) inside the pass (NCCL)",
                           "schedule": "serial passes on one stream, one CUDA-graph replay per pass" if use_graph else "serial eager passes"},
                "records_per_gpu": R_local, "loci_typed": sum(len(v) for v in result.values()), "species_typed": len(result), "cuda_graph": use_graph,
                "gpu_launches": pipe.launches, "parity_locus_sample": parity, "setup_seconds": {"db": t_db, "generate_sort_cap_pack": t_gen}}
        bench.emit(line)
    if world > 1:
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stderr.flush()
        os._exit(0)


# ----------------------------------------------------------------------------------------------------------------
def run_c4(args):
    import torch
    from metamlst_b200 import native, sample, synth
    from oracle import cpu_path
    native.lib()
    n_dev = args.devices or torch.cuda.device_count()
    n_reads = args.reads or 250_000
    args.alleles = args.alleles
    db = bench.make_db(args)
    work = tempfile.mkdtemp(prefix="mmlst_c4_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    db_path = os.path.join(work, "db.sqlite")
    db.write_sqlite(db_path)
    t0 = time.perf_counter()
    paths, keep_cores = [], {}
    bam_bytes = 0
    for i in range(args.samples):
        cores = [synth.gen_core(db, n_reads, args.read_len, seed=4000 + i, K=args.k, org_props=bench.PROPS, device="cuda:0", strain_seed=1004 + i % 5)]
        p = os.path.join(work, "s%03d.bam" % i)
        bam_bytes += int(synth.write_bam_fast(db, cores, order="name", align_records=True, path=p).size)
        paths.append(p)
        if i in (0, args.samples // 2, args.samples - 1):
            keep_cores[i] = cores
    t_gen = time.perf_counter() - t0
    params = dict(ingest=args.ingest, engine="device", **{k: bench.PARAMS[k] for k in ("minscore", "max_xM", "min_read_len", "penalty")})
    out_warm = os.path.join(work, "warm")
    runs = []
    if args.workers == "processes":
        pool = sample.CohortPool(db_path, [d for d in range(n_dev) for _ in range(max(1, args.per_gpu))], **params)
        pool.type(paths[:2 * n_dev * max(1, args.per_gpu)], out_warm)   # builds every worker's tables and workspace once
        typers = {}
        for rep in range(2):
            t0 = time.perf_counter()
            results = pool.type(paths, os.path.join(work, "out%d" % rep))
            runs.append(time.perf_counter() - t0)
        pool.close()
    else:
        typers = {d: sample.SampleTyper(db_path, device=d, **params) for d in range(n_dev)}
        sample.type_cohort(paths[:2 * n_dev], db_path, out_warm, devices=list(range(n_dev)), typers=typers)
        for rep in range(2):
            for d in range(n_dev):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            results = sample.type_cohort(paths, db_path, os.path.join(work, "out%d" % rep), devices=list(range(n_dev)), typers=typers, prefetch=3)
            for d in range(n_dev):
                torch.cuda.synchronize(d)
            runs.append(time.perf_counter() - t0)
    dt = min(runs)
    assert len(results) == args.samples
    records = sum(r.records for r in results)
    lat = np.asarray([r.seconds["device_ingest"] + r.seconds["gpu"] + r.seconds["format"] for r in results])
    # ---- parity: three samples against the C port of the oracle on the same reads, and against the host-unpack / host-seam engine
    checked = []
    host = sample.SampleTyper(db_path, device=0, ingest="host", engine="host", **{k: bench.PARAMS[k] for k in ("minscore", "max_xM", "min_read_len", "penalty")})
    for i, cores in keep_cores.items():
        w = cpu_path.workload_from_cores(db, cores)
        ref = cpu_path.run(w, threads=os.cpu_count() or 1, max_depth=8000, minscore=bench.PARAMS["minscore"], max_xM=bench.PARAMS["max_xM"],
                           min_read_len=bench.PARAMS["min_read_len"], penalty=bench.PARAMS["penalty"])
        got = {c.species: [(l.contig, l.sequence, l.holes, l.snps) for l in c.loci] for c in results[i].calls if c.passed_gate}
        # the cohort BAMs are in bowtie2 (name) order: species / locus order follows the FILE order, the C port sees coordinate order -> compare as sets
        assert {sp: sorted(v) for sp, v in got.items()} == {sp: sorted(v) for sp, v in ref["result"].items()}, "sample %d differs from the C port" % i
        assert (results[i].total_reads, results[i].ignored_reads) == (int(ref["counters"][0]), int(ref["counters"][1]))
        r2 = host.type_bam(paths[i], os.path.join(work, "host"))
        assert r2.nfo_lines == results[i].nfo_lines, "sample %d: device ingest + device engine and host unpack + host seams write different .nfo lines" % i
        checked.append(i)
    host.close()
    nfo = sum(1 for r in results for _ in r.nfo_lines)
    for t in typers.values():
        t.close()
    line = {"metric": "aligned reads/s (score+pileup+consensus), cohort", "value": records / dt, "unit": "records/s", "n_gpus": n_dev, "higher_is_better": True,
            "data": "synthetic", "dtype": "int32/u8",
            "config": {"workload": "configs[3]: %d-sample cohort, every sample %d x %d bp reads, K=%d (%d records), bowtie2-ordered BAM files (%.2f GB in total) against "
                                   "the configs[1] DB (21 loci x %d alleles); files -> `.nfo` lines" % (args.samples, n_reads, args.read_len, args.k, n_reads * args.k,
                                                                                                          bam_bytes / 1e9, args.alleles),
                       "schedule": ("sample.CohortPool: %d PROCESS(es) per GPU (each with its loader threads)" % max(1, args.per_gpu) if args.workers == "processes" else
                                    "sample.type_cohort: one worker thread + two loader threads per GPU") + ", samples dealt round-robin; ingest=%s (%s), engine=device "
                                   "(one kernel chain per sample, device-side selection)" % (args.ingest, "compressed bytes cross PCIe, hardware DEFLATE + parse/sort/cap/pack kernels"
                                                                                             if args.ingest == "device" else "C++ host unpacker")},
            "whole_box_seconds": dt, "whole_box_seconds_both_runs": runs, "workers": args.workers, "samples": args.samples, "samples_per_s": args.samples / dt, "records_total": records, "nfo_lines_written": nfo,
            "per_sample_latency_ms": {"median": float(np.median(lat) * 1e3), "p90": float(np.percentile(lat, 90) * 1e3), "max": float(lat.max() * 1e3),
                                      "what": "device ingest + kernel chain + formatting of one sample, measured by its worker thread"},
            "per_sample_phase_ms_median": {k: float(np.median([r.seconds.get(k, 0.0) for r in results]) * 1e3) for k in ("load", "device_ingest", "gpu", "format")},
            "per_sample_records_per_s_median": float(np.median([r.records / max(l, 1e-9) for r, l in zip(results, lat)])),
            "parity": {"ok": True, "samples_checked": checked, "against": "C port of the oracle on the same reads (loci, consensus, holes, SNPs, read counters) and the "
                                                                          "host-unpack + host-seam engine (.nfo lines)"},
            "setup_seconds": {"generate_and_write_bams": t_gen}}
    bench.emit(line)
    import shutil
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    a = parse()
    bench.protect_stdout()
    (run_c3 if a.config == "c3" else run_c4)(a)
