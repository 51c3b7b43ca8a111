"""N>1 host logic on CPU: world_size-2 and -4 `gloo` processes run the same reductions the GPU ranks run over NCCL
(metamlst_b200/dist.py).  Per-rank partial tables come from the C oracle on contig-aligned shards; the reduced tables
must equal the oracle on the whole sample, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from helpers import lut_from_db, small_case
from metamlst_b200 import api, dist
from oracle import corc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db, tab = small_case(seed=61, n_reads=3000, L=100, K=4, orgs=("ecoli", "saureus"), apl=5)
        st = tab.sorted_by_coord()
        allow, locus_of, n_loci = lut_from_db(db)
        per_contig = np.bincount(st.tid, minlength=db.n_rows)
        shards = dist.shard_contigs(per_contig, locus_of, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(db.n_rows))
        mine = st.take(np.nonzero(np.isin(st.tid, shards[rank]))[0])
        # global record index of every local record (H5 first-record order must survive the sharding)
        orig = np.nonzero(np.isin(st.tid, shards[rank]))[0].astype(np.uint32)
        s, c, f, cnt = corc.score(mine, allow, locus_of, n_loci, 80, 5, 50, orig_idx=orig)
        t_s, t_c = torch.from_numpy(s.copy()), torch.from_numpy(c.view(np.int32).copy())
        t_f, t_cnt = torch.from_numpy(f.view(np.int32).copy()), torch.from_numpy(cnt.view(np.int64).copy())
        dist.allreduce_score_tables(t_s, t_c, t_f, t_cnt)
        ws, wc, wf, wcnt = corc.score(st, allow, locus_of, n_loci, 80, 5, 50)
        assert np.array_equal(t_s.numpy(), ws) and np.array_equal(t_c.numpy().view(np.uint32), wc)
        assert np.array_equal(t_f.numpy().view(np.uint32), wf), "first-record indices (H5) differ after MIN reduce"
        assert np.array_equal(t_cnt.numpy().view(np.uint64), wcnt)
        # the two-call form: [sum_as | counters | n_hit pairs] as ONE int64 SUM + first_idx MIN
        n_ref = s.shape[0]
        zb = torch.zeros(n_ref + 2 + (n_ref + 1) // 2, dtype=torch.int64)
        zb[:n_ref] = torch.from_numpy(s.copy()); zb[n_ref:n_ref + 2] = torch.from_numpy(cnt.view(np.int64).copy())
        zb[n_ref + 2:].view(torch.int32)[:n_ref] = torch.from_numpy(c.view(np.int32).copy())
        f2 = torch.from_numpy(f.view(np.int32).copy())
        dist.allreduce_score_block(zb, f2)
        assert np.array_equal(zb[:n_ref].numpy(), ws) and np.array_equal(zb[n_ref:n_ref + 2].numpy().view(np.uint64), wcnt)
        assert np.array_equal(zb[n_ref + 2:].view(torch.int32)[:n_ref].numpy().view(np.uint32), wc) and np.array_equal(f2.numpy().view(np.uint32), wf)
        # every rank now selects the same alleles
        index = api.AlleleIndex(tab.ref_names)
        chosen = api.fast_select(index, t_s.numpy(), t_c.numpy().view(np.uint32), t_f.numpy().view(np.uint32), 100)
        tids = [t for _sp, ts in chosen for t in ts]
        # owner mode: each rank selects among ITS loci from ITS partial tables only; one all-gather of the per-rank results and
        # the merge (gate + H5 order) must reproduce the whole-sample selection
        sp_names = []
        for sp, _g in index.locus_names:
            if sp not in sp_names:
                sp_names.append(sp)
        gdb = [sum(1 for sp2, _g in index.locus_names if sp2 == sp) for sp in sp_names]
        local = api.fast_select(index, s, c, f, 100)
        blk = {"tid": [], "species": [], "first": [], "payload": []}
        for sp, ts in local:
            for t in ts:
                l = int(index.locus_of[t])
                rows_l = np.nonzero((index.locus_of == l) & (c > 0))[0]
                blk["tid"].append(int(t)); blk["species"].append(sp_names.index(sp)); blk["first"].append(int(f[rows_l].min())); blk["payload"].append(int(t) * 7)
        parts = [None] * world
        td.all_gather_object(parts, blk)
        for nloci in (100, 60, 0):
            merged = dist.merge_owner_blocks(parts, sp_names, gdb, nloci)
            want = [(sp, [(t, t * 7) for t in ts]) for sp, ts in chosen if int((float(len(ts)) / float(gdb[sp_names.index(sp)])) * 100) >= nloci]
            assert merged == want, (nloci, merged, want)
        # pileup counts: a rank contributes the contigs it owns (depth cap local to the contig), zeros elsewhere
        lens = [int(st.ref_lens[t]) for t in tids]
        col_off = np.concatenate([[0], np.cumsum(lens)])
        counts = np.zeros((int(col_off[-1]), 5), np.int32)
        whole = np.zeros_like(counts)
        own = set(int(t) for t in shards[rank])
        for i, t in enumerate(tids):
            w, _ = corc.contig_counts(st, t, 20, 80, 5, 60)
            whole[col_off[i]:col_off[i + 1]] = w
            if t in own:
                m, _ = corc.contig_counts(mine, t, 20, 80, 5, 60)
                counts[col_off[i]:col_off[i + 1]] = m
        t_counts = torch.from_numpy(counts)
        dist.allreduce_counts(t_counts)
        assert np.array_equal(t_counts.numpy(), whole)
        # Hamming: rows sharded, best = (dist << 32 | global row) reduced with MIN
        rng = np.random.default_rng(5)
        rows = ["".join("ACGT"[x] for x in rng.integers(0, 4, int(n))) for n in rng.integers(40, 90, 300)]
        rows[17] = rows[250]  # a tie between shards must resolve to the lowest global row
        qs = [rows[250], rows[3][:50], "ACGT" * 12]
        flat = np.frombuffer("".join(rows).encode(), np.uint8)
        off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
        lo, hi = dist.shard_rows(len(rows), world, rank)
        d, a = corc.hamming_min([q.encode() for q in qs], flat, off, [(lo, hi)] * len(qs))
        best = torch.from_numpy(((d.astype(np.uint64) << np.uint64(32)) | a.astype(np.uint64)).view(np.int64).copy())
        if rank == 1:
            best[2] = -1  # "nothing found on this shard" (preset ~0) must lose against any real key
        dist.allreduce_best(best)
        wd, wa = corc.hamming_min([q.encode() for q in qs], flat, off, [(0, len(rows))] * len(qs))
        got = best.numpy().view(np.uint64)
        # query 2: the best over every shard but rank 1's
        keys = []
        for r in range(world):
            if r != 1:
                d0, a0 = corc.hamming_min([qs[2].encode()], flat, off, [dist.shard_rows(len(rows), world, r)])
                keys.append((int(d0[0]), int(a0[0])))
        wd[2], wa[2] = min(keys)
        assert np.array_equal(got >> np.uint64(32), wd.astype(np.uint64)) and np.array_equal(got & np.uint64(0xffffffff), wa.astype(np.uint64))
        assert int(got[0] & np.uint64(0xffffffff)) == 17
        assert dist.max_over_ranks(float(rank + 1), "cpu") == float(world)
        ret[rank] = "ok"
    except BaseException as e:  # noqa: BLE001
        ret[rank] = repr(e)
        raise
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_reductions_over_ranks_equal_the_whole_sample(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_shard_helpers():
    rec = [100, 5, 5, 90, 1, 1, 50, 50]
    loc = [0, 0, 1, 1, 2, 2, 3, 3]
    sh = dist.shard_contigs(rec, loc, 2)
    assert sorted(np.concatenate(sh).tolist()) == list(range(8))
    for s in sh:  # whole loci stay together
        assert all((loc[a] in {loc[b] for b in s}) for a in s)
    loads = [sum(rec[i] for i in s) for s in sh]
    assert abs(loads[0] - loads[1]) <= 100
    assert dist.shard_rows(100, 4, 0) == (0, 32) and dist.shard_rows(100, 4, 3) == (96, 100)
    assert dist.shard_rows(10, 4, 2) == (10, 10)
