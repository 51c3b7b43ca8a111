"""streams.DeviceStreams.from_soa: the bridge from an unpacked sample (SoaHost) to the device pipeline, whole or sharded.

GPU: shards of one sample ("ranges": any contiguous record ranges; "loci": contig-aligned) typed one after another on one
device must add up, integer for integer, to the whole sample -- what the N-rank all-reduce relies on -- and the whole
sample must equal the C port of the oracle.  A real 2-process run (torchrun, NCCL) over a golden BAM is at the end: it
needs two GPUs and is skipped on a one-GPU box.
CPU: the shard arithmetic itself (which records / plane words a rank takes) on a CPU tensor device.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from helpers import lut_from_db, small_case
from metamlst_b200 import api, packing, streams
from oracle import corc


def _soa(seed=23, n_reads=5000, max_depth=60):
    db, tab = small_case(seed=seed, n_reads=n_reads, L=100, K=4, orgs=("ecoli", "saureus"), apl=5)
    return db, tab, packing.pack_table(tab, max_depth=max_depth)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", ["ranges", "loci"])
def test_shards_partition_the_sample_on_cpu(mode, world):
    db, tab, soa = _soa()
    index = api.AlleleIndex(soa.ref_names)
    seen = np.zeros(soa.n_rec, np.int32)
    words = 0
    for r in range(world):
        st = streams.DeviceStreams.from_soa(soa, "cpu", rank=r, world=world, mode=mode, locus_of=index.locus_of)
        n = int(st.tid.shape[0])
        if st.orig_idx is not None:
            gi = st.orig_idx.numpy().view(np.uint32).astype(np.int64)
        else:
            gi = st.idx_base + np.arange(n)
            if soa.orig_idx is not None:
                gi = soa.orig_idx[gi]
        back = np.argsort(soa.orig_idx) if soa.orig_idx is not None else None
        pos_in_stream = gi if back is None else np.searchsorted(soa.orig_idx[back], gi)
        where = pos_in_stream if back is None else back[pos_in_stream]
        assert np.array_equal(soa.tid[where].view(np.int32), st.tid.numpy()) and np.array_equal(soa.as0[where], st.as0.numpy())
        seen[where] += 1
        # pileup stream: rebased rows decode to the same words
        P = int(st.n_prec)
        if mode == "loci":
            assert P == soa.n_prec
        recs = st.p_recs.numpy().reshape(-1).view(packing.PREC_DTYPE)
        for i in (0, P // 2, P - 1) if P else ():
            g = i + (soa.n_prec * r) // world if mode == "ranges" else i
            a, b = int(recs["row_off"][i]), int(soa.p_recs["row_off"][g])
            nw = 3 * int(recs["nw"][i])
            assert np.array_equal(st.planes.numpy()[a:a + nw].view(np.uint32), soa.planes[b:b + nw])
        cs = st.contig_start.astype(np.int64)
        assert cs[0] == 0 and cs[-1] == P and np.all(np.diff(cs) >= 0)
        words += int(st.planes.shape[0]) - packing.PLANE_SLACK_WORDS
        if st.run_tid is not None:  # run arrays rebuilt for the shard describe the shard
            rs = st.run_start.numpy().view(np.uint32)
            assert rs[0] == 0 and rs[-1] == n and np.array_equal(np.repeat(st.run_tid.numpy(), np.diff(rs.astype(np.int64))), st.tid.numpy())
    assert np.all(seen == 1)
    if mode == "ranges":
        assert words == int(soa.p_row_off[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("mode", ["ranges", "loci"])
def test_shards_add_up_to_the_whole_sample_and_to_the_oracle(mode, world):
    from metamlst_b200 import pipeline
    db, tab, soa = _soa()
    index = api.AlleleIndex(soa.ref_names)
    st_sorted = tab.sorted_by_coord()
    allow, locus_of, n_loci = lut_from_db(db)
    ws, wc, wf, wcnt = corc.score(tab, allow, locus_of, n_loci, 80, 5, 50)  # file order == the order stage 1 reads
    chosen = api.fast_select(index, ws, wc, wf, 100)
    tids = [t for _sp, ts in chosen for t in ts]
    tot_s, tot_c, tot_f, tot_k, tot_counts = 0, 0, None, 0, 0
    for r in range(world):
        st = streams.DeviceStreams.from_soa(soa, "cuda:0", rank=r, world=world, mode=mode, locus_of=index.locus_of)
        pipe = pipeline.DevicePipeline(st, index, db.row_seq, minscore=80, max_xM=5, min_read_len=50, penalty=100)
        pipe.run_score()
        torch.cuda.synchronize()
        s, c = pipe.sum_as.cpu().numpy(), pipe.n_hit.cpu().numpy().view(np.uint32).astype(np.int64)
        f, k = pipe.first_idx.cpu().numpy().view(np.uint32), pipe.counters.cpu().numpy().view(np.uint64).astype(np.int64)
        if mode == "loci" and world > 1:
            assert np.all((index.locus_of[np.nonzero(c)[0]] % world) == r)
        tot_s, tot_c, tot_k = tot_s + s, tot_c + c, tot_k + k
        tot_f = f if tot_f is None else np.minimum(tot_f, f)
        pipe.run_pileup_consensus(tids)
        ncol = int(sum(int(soa.ref_lens[t]) for t in tids))
        cnt = pipe.counts[: ncol * 5].cpu().numpy().astype(np.int64)
        tot_counts = tot_counts + cnt if mode == "ranges" else np.maximum(tot_counts, cnt)  # loci shards all hold the whole pileup stream
    assert np.array_equal(tot_s, ws) and np.array_equal(tot_c, wc.astype(np.int64)) and np.array_equal(tot_f, wf)
    assert tot_k.tolist() == wcnt.astype(np.int64).tolist()
    want = np.concatenate([corc.contig_counts(st_sorted, t, 20, 80, 5, 60)[0].reshape(-1) for t in tids]).astype(np.int64)
    assert np.array_equal(tot_counts, want)


@pytest.mark.gpu
def test_one_real_sample_sharded_over_two_ranks(tmp_path):
    """torchrun x 2 (NCCL): both ranks unpack the same golden BAM, each types its record range through
    SampleTyper(engine="device"), the tables are all-reduced; rank 0 writes files identical to the reference's."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "out")
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29871", os.path.join(ROOT, "tests", "run_sharded_sample.py"), out], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout.decode()[-3000:]
    assert "SHARDED-OK" in p.stdout.decode()


@pytest.mark.parametrize("world", [2, 5])
def test_slicing_resident_streams_equals_sharding_the_host_sample(world):
    """DeviceStreams.slice_ranges (a device-resident sample cut after packing: bench_configs.py c3) == from_soa(mode="ranges")."""
    db, tab, soa = _soa()
    whole = streams.DeviceStreams.from_soa(soa, "cpu")
    for r in range(world):
        a = whole.slice_ranges(r, world)
        b = streams.DeviceStreams.from_soa(soa, "cpu", rank=r, world=world, mode="ranges")
        for k in ("tid", "as0", "xm3", "qlen", "p_recs", "planes"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
        assert (a.orig_idx is None) == (b.orig_idx is None) and (a.orig_idx is None or torch.equal(a.orig_idx, b.orig_idx))
        assert a.idx_base == b.idx_base and a.n_prec == b.n_prec and np.array_equal(a.contig_start, b.contig_start)
        assert (a.run_tid is None) == (b.run_tid is None) and (a.run_tid is None or torch.equal(a.run_tid, b.run_tid))


@pytest.mark.gpu
def test_c_abi_allreduce_over_two_gpus(tmp_path):
    """mmlst_comm_* / mmlst_allreduce (include/mmlst.h): two plain processes, no torch.distributed."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "run_c_allreduce.py"), str(r), "2", str(tmp_path)], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "C-ALLREDUCE-OK rank %d" % r in o, o[-3000:]
