"""Device-resident driver of the hot path: score -> best-allele selection -> pileup -> consensus.

One instance per GPU / process.  Inputs are the packed streams already in HBM (torch tensors); the kernels are
launched through the `*_dev` C-ABI on torch's current stream.  With torch.distributed initialised (NCCL, one process
per GPU) the partial integer tables are combined with all-reduce (SUM for scores/hits/counts, MIN for first indices),
exactly the exchange SURVEY.md 8e names; shards must be contig-aligned so that the htslib depth cap stays local.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import api, native

def build_chunks(contig_start: np.ndarray, chosen_tid: Sequence[int], col_off: np.ndarray) -> np.ndarray:
    """mmlst_chunk descriptors (8 x u32 each) for the chosen contigs of a device-resident pileup stream."""
    out = []
    total = sum(int(contig_start[t + 1]) - int(contig_start[t]) for t in chosen_tid)
    CHUNK_RECORDS = int(native.lib().mmlst_chunk_records(total))  # whole tiles, <= 63 tiles, >= 4 chunks per SM
    for l, t in enumerate(chosen_tid):
        r0, r1 = int(contig_start[t]), int(contig_start[t + 1])
        for b in range(r0, r1, CHUNK_RECORDS):
            out.append((b, min(r1, b + CHUNK_RECORDS), int(col_off[l]), int(col_off[l + 1] - col_off[l]), 0, 0, 0, 0))
    return np.asarray(out, dtype=np.uint32).reshape(-1, 8)


class DevicePipeline:
    def __init__(self, streams, index: api.AlleleIndex, dbseq_of: Callable[[int], str], minscore: int = 80, max_xM: int = 5,
                 min_read_len: int = 50, penalty: int = 100, species_filter: Optional[str] = None, mincov: int = 1,
                 impl: int = 0, idx_base: int = 0, group=None):
        self.s = streams
        self.index = index
        self.dbseq_of = dbseq_of
        self.minscore, self.max_xM, self.min_read_len, self.penalty = int(minscore), int(max_xM), int(min_read_len), int(penalty)
        self.mincov, self.impl, self.idx_base = int(mincov), int(impl), int(idx_base)
        self.group = group
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        dev = streams.tid.device
        self.dev = dev
        n_ref = len(index.ref_names)
        self.n_ref = n_ref
        self.allow = torch.from_numpy(index.allow_mask(species_filter)).to(dev)
        self.locus_of = torch.from_numpy(index.locus_of.astype(np.int32)).to(dev)
        # one int64 block so a single D2H / all-reduce moves the three score tables
        self.sum_as = torch.zeros(n_ref, dtype=torch.int64, device=dev)
        self.n_hit = torch.zeros(n_ref, dtype=torch.int32, device=dev)
        self.first_idx = torch.zeros(n_ref, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(2, dtype=torch.int64, device=dev)
        self.max_cols = int(np.sort(np.asarray(streams.ref_lens))[::-1][: index.n_loci].sum())
        self.counts = torch.zeros(self.max_cols * 5 + 8, dtype=torch.int32, device=dev)
        self.cons = torch.zeros(self.max_cols + 8, dtype=torch.uint8, device=dev)
        self.holes = torch.zeros(index.n_loci + 1, dtype=torch.int32, device=dev)
        self.snps = torch.zeros(index.n_loci + 1, dtype=torch.int32, device=dev)
        self._db_cache: Dict[Tuple[int, ...], tuple] = {}
        self.lib = native.lib()
        self.timers: Optional[Dict[str, list]] = None  # name -> [(start_event, end_event)]
        self.launches = 0

    # ------------------------------------------------------------------------------------------------------------
    def _timed(self, name: str, fn):
        if self.timers is None:
            fn()
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        self.timers.setdefault(name, []).append((a, b))

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def run_score(self):
        s = self.s
        self.sum_as.zero_(); self.n_hit.zero_(); self.first_idx.fill_(-1); self.counters.zero_()
        def k():
            native.check(self.lib.mmlst_score_dev(native.ptr(s.tid), native.ptr(s.as0), native.ptr(s.xm3), native.ptr(s.qlen), 0,
                                                  int(s.tid.shape[0]), self.idx_base, native.ptr(self.allow), native.ptr(self.locus_of),
                                                  self.n_ref, self.minscore, self.max_xM, self.min_read_len, native.ptr(self.sum_as),
                                                  native.ptr(self.n_hit), native.ptr(self.first_idx), native.ptr(self.counters),
                                                  self._stream()))
        self._timed("score", k)
        self.launches += 1
        if self.dist:
            d = torch.distributed
            d.all_reduce(self.sum_as, op=d.ReduceOp.SUM, group=self.group)
            d.all_reduce(self.n_hit, op=d.ReduceOp.SUM, group=self.group)
            self.first_idx.bitwise_xor_(-2147483648)  # u32 order -> i32 order
            d.all_reduce(self.first_idx, op=d.ReduceOp.MIN, group=self.group)
            self.first_idx.bitwise_xor_(-2147483648)
            d.all_reduce(self.counters, op=d.ReduceOp.SUM, group=self.group)

    def select(self):
        sum_as = self.sum_as.cpu().numpy()
        n_hit = self.n_hit.cpu().numpy().view(np.uint32)
        first = self.first_idx.cpu().numpy().view(np.uint32)
        return api.fast_select(self.index, sum_as, n_hit, first, self.penalty), (sum_as, n_hit, first)

    def _db_for(self, tids: Tuple[int, ...]):
        hit = self._db_cache.get(tids)
        if hit is None:
            lens = [int(self.s.ref_lens[t]) for t in tids]
            col_off = np.zeros(len(tids) + 1, dtype=np.uint32)
            col_off[1:] = np.cumsum(lens)
            db = np.zeros(int(col_off[-1]) + 8, dtype=np.uint8)
            for i, t in enumerate(tids):
                sq = self.dbseq_of(t)
                if len(sq) < lens[i]:
                    raise IndexError("string index out of range: BAM LN > DB sequence length for %s" % self.index.ref_names[t])
                db[col_off[i]:col_off[i + 1]] = np.frombuffer(sq[:lens[i]].encode("latin-1"), dtype=np.uint8)
            hit = (col_off, torch.from_numpy(db).to(self.dev), torch.from_numpy(col_off.view(np.int32)).to(self.dev))
            if len(self._db_cache) > 64:
                self._db_cache.clear()
            self._db_cache[tids] = hit
        return hit

    def run_pileup_consensus(self, tids: Sequence[int]):
        tids = tuple(int(t) for t in tids)
        s = self.s
        col_off, db_d, col_off_d = self._db_for(tids)
        total = int(col_off[-1])
        chunks = build_chunks(s.contig_start, tids, col_off)
        self.counts[: total * 5].zero_()
        if chunks.shape[0]:
            chunks_d = torch.from_numpy(chunks.view(np.int32)).to(self.dev, non_blocking=True)
            def k():
                native.check(self.lib.mmlst_pileup_dev(native.ptr(s.p_pos), native.ptr(s.p_row_off), native.ptr(s.p_reflen), native.ptr(s.p_as),
                                                       native.ptr(s.p_xm), native.ptr(s.planes), native.ptr(chunks_d), int(chunks.shape[0]),
                                                       int(s.max_row_words), self.minscore, self.max_xM, native.ptr(self.counts), total,
                                                       self.impl, self._stream()))
            self._timed("pileup", k)
            self.launches += 1
        if self.dist:
            torch.distributed.all_reduce(self.counts[: total * 5], op=torch.distributed.ReduceOp.SUM, group=self.group)
        def k2():
            native.check(self.lib.mmlst_consensus_dev(native.ptr(self.counts), native.ptr(db_d), native.ptr(col_off_d), len(tids), self.mincov,
                                                      native.ptr(self.cons), native.ptr(self.holes), native.ptr(self.snps), self._stream()))
        self._timed("consensus", k2)
        self.launches += 1
        cons = self.cons[:total].cpu().numpy()
        holes = self.holes[: len(tids)].cpu().numpy()
        snps = self.snps[: len(tids)].cpu().numpy()
        return [cons[col_off[i]:col_off[i + 1]].tobytes().decode("latin-1") for i in range(len(tids))], holes, snps, col_off

    def step(self):
        """One pass of the hot path over the resident sample.  Returns {species: [(contig, consensus, holes, snps)]}."""
        self.run_score()
        chosen, _raw = self.select()
        tids = [t for _sp, ts in chosen for t in ts]
        out: Dict[str, list] = {}
        if tids:
            seqs, holes, snps, _ = self.run_pileup_consensus(tids)
            i = 0
            for sp, ts in chosen:
                for t in ts:
                    out.setdefault(sp, []).append((self.index.ref_names[t], seqs[i], int(holes[i]), int(snps[i])))
                    i += 1
        return out
