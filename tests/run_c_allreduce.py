"""Worker of tests/test_streams.py::test_c_abi_allreduce_over_two_gpus: two plain processes (NO torch.distributed), one GPU each, joined by the C-ABI
communicator (mmlst_comm_unique_id / mmlst_comm_create / mmlst_allreduce); each scores and piles up its record range of one sample, the all-reduced
tables must be the oracle's tables of the whole sample."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import lut_from_db, small_case  # noqa: E402
from metamlst_b200 import api, native, packing, pipeline, streams  # noqa: E402
from oracle import corc  # noqa: E402


def main():
    rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    lib = native.lib()
    torch.cuda.set_device(rank)
    idb = np.zeros(128, np.uint8)
    idf = os.path.join(tmp, "nccl_id.bin")
    if rank == 0:
        native.check(lib.mmlst_comm_unique_id(native.ptr(idb)))
        with open(idf + ".tmp", "wb") as fh:
            fh.write(idb.tobytes())
        os.replace(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            assert time.time() - t0 < 120, "no id from rank 0"
            time.sleep(0.05)
        idb = np.frombuffer(open(idf, "rb").read(), np.uint8).copy()
    comm = C.c_void_p()
    native.check(lib.mmlst_comm_create(native.ptr(idb), rank, world, rank, C.byref(comm)))
    db, tab = small_case(seed=23, n_reads=5000, L=100, K=4, orgs=("ecoli", "saureus"), apl=5)
    soa = packing.pack_table(tab, max_depth=60)
    index = api.AlleleIndex(soa.ref_names)
    st = streams.DeviceStreams.from_soa(soa, "cuda:%d" % rank, rank=rank, world=world, mode="ranges")
    pipe = pipeline.DevicePipeline(st, index, db.row_seq, minscore=80, max_xM=5, min_read_len=50, penalty=100)
    assert not pipe.dist
    s = torch.cuda.current_stream().cuda_stream
    pipe.reset_tables()
    pipe._score_call()
    native.check(lib.mmlst_allreduce(comm, native.ptr(pipe.zscore), int(pipe.zscore.shape[0]), native.ptr(pipe.first_idx), pipe.n_ref, None, 0, s))
    torch.cuda.synchronize()
    allow, locus_of, n_loci = lut_from_db(db)
    ws, wc, wf, wcnt = corc.score(tab, allow, locus_of, n_loci, 80, 5, 50)
    assert np.array_equal(pipe.sum_as.cpu().numpy(), ws) and np.array_equal(pipe.n_hit.cpu().numpy().view(np.uint32), wc)
    assert np.array_equal(pipe.first_idx.cpu().numpy().view(np.uint32), wf) and pipe.counters.cpu().numpy().view(np.uint64).tolist() == wcnt.tolist()
    chosen = api.fast_select(index, ws, wc, wf, 100)
    tids = [t for _sp, ts in chosen for t in ts]
    pipe.run_pileup_consensus(tids)
    ncol = int(sum(int(soa.ref_lens[t]) for t in tids))
    native.check(lib.mmlst_allreduce(comm, None, 0, None, 0, native.ptr(pipe.counts), ncol * 5, s))
    torch.cuda.synchronize()
    srt = tab.sorted_by_coord()
    want = np.concatenate([corc.contig_counts(srt, t, 20, 80, 5, 60)[0].reshape(-1) for t in tids])
    assert np.array_equal(pipe.counts[: ncol * 5].cpu().numpy().view(np.uint32), want)
    lib.mmlst_comm_destroy(comm)
    print("C-ALLREDUCE-OK rank %d" % rank)


if __name__ == "__main__":
    main()
