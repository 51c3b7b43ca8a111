"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` (NCCL on the GPUs,
gloo in the CPU tests).  Everything exchanged is an integer table, so every reduction is bit-exact and independent of
the order ranks arrive in.

    score      records sharded by contig (allele row) => per-rank partial (sum_as, n_hit, first_idx, counters)
               all-reduce SUM / SUM / MIN / SUM                                   (metamlst.py:118-130 aggregated)
    depth cap  sequential per contig (H1) => contig-aligned shards keep it rank-local, no exchange
    pileup     per-rank partial count tensor, all-reduce SUM                        (cmseq/cmseq.py:541-548 aggregated)
    owner mode contig-aligned shards hold WHOLE loci, so a rank can select, pile up and call its own loci with no partial
               tensor at all; the only exchange is ONE all-gather of the per-rank result blocks (chosen contig, first
               record, consensus, holes, SNPs), after which the --nloci gate and the H5 order are applied on the merged
               list (merge_owner_blocks).  The all-reduce form stays for shards that cut through loci.
    Hamming    DB rows sharded, queries replicated, best[q] = (distance << 32 | global row) all-reduce MIN
               (ties -> lowest row, metamlst-merge.py:177-181 order)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as td

SIGN32 = -2147483648
SIGN64 = -9223372036854775808


def active(group=None) -> bool:
    return td.is_available() and td.is_initialized() and td.get_world_size(group) > 1


def shard_contigs(records_per_contig: Sequence[int], locus_of: Sequence[int], world: int) -> List[np.ndarray]:
    """Contig-aligned shards: whole LOCI are dealt to ranks (largest first, always to the lightest rank), so that
    every contig's records -- and with them the sequential htslib depth-cap admission -- stay on one rank, and the
    per-locus scoring tables of a rank are complete for its loci.  Returns the contig (tid) list of every rank."""
    rec = np.asarray(records_per_contig, dtype=np.int64)
    loc = np.asarray(locus_of, dtype=np.int64)
    n_loci = int(loc.max()) + 1 if loc.size else 0
    per_locus = np.bincount(loc, weights=rec, minlength=n_loci)
    load = np.zeros(world, dtype=np.float64)
    owner = np.zeros(n_loci, dtype=np.int64)
    for l in sorted(range(n_loci), key=lambda l: (-per_locus[l], l)):
        r = int(np.argmin(load))
        owner[l] = r
        load[r] += per_locus[l]
    return [np.nonzero(owner[loc] == r)[0] for r in range(world)]


def shard_rows(n_rows: int, world: int, rank: int, align: int = 32) -> Tuple[int, int]:
    """Row range [lo, hi) of a rank for the Hamming search, aligned to the 32-row tiles of the DB layout."""
    per = -(-n_rows // world)
    per = -(-per // align) * align
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def allreduce_score_tables(sum_as: torch.Tensor, n_hit: torch.Tensor, first_idx: torch.Tensor, counters: torch.Tensor,
                           group=None) -> None:
    """In place.  sum_as int64 SUM, n_hit int32 SUM, counters int64 SUM, first_idx (u32 carried in int32) MIN in
    UNSIGNED order: x ^ 0x80000000 maps unsigned order onto the signed order the backends reduce in."""
    if not active(group):
        return
    td.all_reduce(sum_as, op=td.ReduceOp.SUM, group=group)
    td.all_reduce(n_hit, op=td.ReduceOp.SUM, group=group)
    first_idx.bitwise_xor_(SIGN32)
    td.all_reduce(first_idx, op=td.ReduceOp.MIN, group=group)
    first_idx.bitwise_xor_(SIGN32)
    td.all_reduce(counters, op=td.ReduceOp.SUM, group=group)


def allreduce_score_block(zscore: torch.Tensor, first_idx: torch.Tensor, group=None) -> None:
    """Same exchange in two calls: `zscore` is the contiguous int64 block [sum_as | counters | n_hit pairs] (two u32 hit
    counts per int64 word add without carrying into each other as long as every total fits u32, which n_hit is by
    type), first_idx MIN as above."""
    if not active(group):
        return
    td.all_reduce(zscore, op=td.ReduceOp.SUM, group=group)
    first_idx.bitwise_xor_(SIGN32)
    td.all_reduce(first_idx, op=td.ReduceOp.MIN, group=group)
    first_idx.bitwise_xor_(SIGN32)


def merge_owner_blocks(blocks: Sequence[dict], species_names: Sequence[str], genes_in_db: Sequence[int], nloci_pct: int):
    """Owner mode: per-rank results -> the sample's result in the reference's order.

    blocks[r] = {"tid": [...], "species": [...species id...], "first": [...first passing record of the locus...],
                 "payload": [...anything per locus...]} for the loci rank r owns and detected.
    metamlst.py:184-206: a species is processed iff int(detected / loci_in_db * 100) >= nloci ("Database is broken" when
    more loci are detected than the DB lists); dict order (H5) = species by their first passing record, loci inside a
    species by theirs.  Returns [(species name, [(tid, payload)])]."""
    per_species = {}
    for b in blocks:
        for t, sp, f, pay in zip(b["tid"], b["species"], b["first"], b["payload"]):
            per_species.setdefault(int(sp), []).append((int(f), int(t), pay))
    out = []
    for sp, lst in per_species.items():
        det, tot = len(lst), int(genes_in_db[sp])
        if tot < det:
            raise RuntimeError("Database is broken: a species has more detected loci than the genes table lists (metamlst.py:188)")
        if int((float(det) / float(tot)) * 100) < nloci_pct:
            continue
        lst.sort()
        out.append((lst[0][0], species_names[sp], [(t, pay) for _f, t, pay in lst]))
    out.sort(key=lambda x: x[0])
    return [(name, lst) for _f, name, lst in out]


def allreduce_counts(counts: torch.Tensor, group=None) -> None:
    if active(group):
        td.all_reduce(counts, op=td.ReduceOp.SUM, group=group)


def allreduce_best(best: torch.Tensor, group=None) -> None:
    """best[q] = (distance << 32 | global row) as u64 carried in int64 (preset ~0 = -1): MIN in unsigned order."""
    if not active(group):
        return
    best.bitwise_xor_(SIGN64)
    td.all_reduce(best, op=td.ReduceOp.MIN, group=group)
    best.bitwise_xor_(SIGN64)


def max_over_ranks(ms: float, device, group=None) -> float:
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if active(group):
        td.all_reduce(t, op=td.ReduceOp.MAX, group=group)
    return float(t.item())
