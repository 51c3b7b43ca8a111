"""Seeded synthetic metamlstDB + pre-aligned read generator.

Produces (a) a synthetic allele/profile database with the reference's 4-table
schema (`/root/reference/metamlst-index.py:62-65`) and (b) a table of bowtie2
style alignment records against that database (`AlnTable`).  The generator is
written with torch ops so the same code makes small fixtures on the CPU (tests,
oracle) and the 10 M-read bench workload on the GPU in under a second.  It is
input plumbing: nothing here is on the measured path.

BAM reference names are `organism_gene_allele` (`metamlst.py:107`,
`metaMLST_functions.py:157`), reference order = allele row order.
"""
from __future__ import annotations

import sqlite3
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# Achtman E. coli / S. aureus / K. pneumoniae scheme lengths (generator parameters)
SCHEMES: Dict[str, List[Tuple[str, int]]] = {
    "ecoli": [("adk", 536), ("fumC", 469), ("gyrB", 460), ("icd", 518), ("mdh", 452), ("purA", 478), ("recA", 510)],
    "saureus": [("arcC", 456), ("aroE", 456), ("glpF", 465), ("gmk", 417), ("pta", 474), ("tpi", 402), ("yqiL", 516)],
    "kpneumoniae": [("gapA", 450), ("infB", 318), ("mdh", 477), ("pgi", 432), ("phoE", 420), ("rpoB", 501), ("tonB", 414)],
}

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class SynthDB:
    """Allele rows in BAM-reference order (organism-major, locus-major, variant ascending)."""

    organisms: List[str]
    loci: Dict[str, List[Tuple[str, int]]]  # organism -> [(gene, length)]
    row_org: np.ndarray  # int32 [A] organism index
    row_locus: np.ndarray  # int32 [A] global locus index (organism-major)
    row_variant: np.ndarray  # int32 [A] alleleVariant (1-based)
    seq_off: np.ndarray  # int64 [A+1] offsets into seq
    seq: np.ndarray  # uint8 ASCII, concatenated allele sequences
    locus_names: List[Tuple[str, str]]  # global locus index -> (organism, gene)
    locus_row0: np.ndarray  # int64 [n_loci+1] first row of each locus
    profiles: Dict[str, np.ndarray]  # organism -> int32 [n_st, n_loci_of_org] alleleVariant per locus; ST code = index+1

    @property
    def n_rows(self) -> int:
        return int(self.row_org.shape[0])

    def row_len(self) -> np.ndarray:
        return (self.seq_off[1:] - self.seq_off[:-1]).astype(np.int32)

    def row_seq(self, r: int) -> str:
        return self.seq[self.seq_off[r]:self.seq_off[r + 1]].tobytes().decode()

    def ref_names(self) -> List[str]:
        out = []
        for r in range(self.n_rows):
            o, g = self.locus_names[int(self.row_locus[r])]
            out.append("%s_%s_%d" % (o, g, int(self.row_variant[r])))
        return out

    def write_sqlite(self, path: str) -> None:
        """Direct sqlite build with the reference schema (`metamlst-index.py:62-65`)."""
        conn = sqlite3.connect(path)
        c = conn.cursor()
        c.execute("CREATE TABLE IF NOT EXISTS organisms (organismkey varchar(255), label VARCHAR(255), PRIMARY KEY(organismkey))")
        c.execute("CREATE TABLE IF NOT EXISTS genes (geneName varchar(255), bacterium VARCHAR(255), PRIMARY KEY(geneName,bacterium))")
        c.execute("CREATE TABLE IF NOT EXISTS alleles (recID INTEGER PRIMARY KEY AUTOINCREMENT,bacterium varchar(255), gene VARCHAR(255), sequence TEXT, alignedSequence TEXT, alleleVariant INT)")
        c.execute("CREATE TABLE IF NOT EXISTS profiles (recID INTEGER PRIMARY KEY AUTOINCREMENT, profileCode INTEGER, bacterium VARCHAR(255), alleleCode INTEGER)")
        for o in self.organisms:
            c.execute("INSERT OR IGNORE INTO organisms (organismkey,label) VALUES (?,?)", (o, "Synthetic " + o))
            for g, _ in self.loci[o]:
                c.execute("INSERT OR IGNORE INTO genes (geneName, bacterium) VALUES (?,?)", (g, o))
        rows = []
        for r in range(self.n_rows):
            o, g = self.locus_names[int(self.row_locus[r])]
            rows.append((g, o, int(self.row_variant[r]), self.row_seq(r)))
        c.executemany("INSERT INTO alleles (gene, bacterium,alleleVariant,sequence) VALUES (?,?,?,?)", rows)
        # recID of row r is r+1 (AUTOINCREMENT from an empty table)
        prof = []
        li0 = 0
        for o in self.organisms:
            nl = len(self.loci[o])
            p = self.profiles[o]
            for st in range(p.shape[0]):
                for j in range(nl):
                    row = int(self.locus_row0[li0 + j]) + int(p[st, j]) - 1
                    prof.append((o, st + 1, row + 1))
            li0 += nl
        c.executemany("INSERT INTO profiles (bacterium, profileCode, alleleCode) VALUES (?,?,?)", prof)
        conn.commit()
        conn.close()

    def write_fasta_typings(self, fasta_path: str, typings_path: str) -> None:
        """FASTA + typings TSV accepted by `metamlst-index.py:92-217`."""
        with open(fasta_path, "w") as f:
            for r, name in enumerate(self.ref_names()):
                f.write(">%s\n%s\n" % (name, self.row_seq(r)))
        with open(typings_path, "w") as f:
            # NB the reference resets `intest` only once per FILE (`metamlst-index.py:147,171-175`), so one
            # typings file carries one organism; callers write one file per organism when several exist.
            assert len(self.organisms) == 1, "one typings file per organism"
            o = self.organisms[0]
            f.write("#%s|Synthetic %s\n" % (o, o))
            f.write("ST\t" + "\t".join(g for g, _ in self.loci[o]) + "\n")
            p = self.profiles[o]
            for st in range(p.shape[0]):
                f.write("%d\t%s\n" % (st + 1, "\t".join(str(int(v)) for v in p[st])))


def make_db(organisms: Sequence[str] = ("ecoli",), alleles_per_locus=256, n_profiles: int = 2048,
            seed: int = 1001, subs_lambda: float = 5.0,
            schemes: Optional[Dict[str, List[Tuple[str, int]]]] = None) -> SynthDB:
    """Each allele = the locus's random base sequence + Poisson(subs_lambda) substitutions.  alleles_per_locus: one count for
    every locus, or a sequence with one count per locus (organism-major order)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    schemes = schemes or SCHEMES
    loci = {o: list(schemes[o]) for o in organisms}
    n_loci_total = sum(len(v) for v in loci.values())
    apl = [int(alleles_per_locus)] * n_loci_total if np.isscalar(alleles_per_locus) else [int(x) for x in alleles_per_locus]
    assert len(apl) == n_loci_total
    row_org, row_locus, row_variant, seqs, locus_names, locus_row0 = [], [], [], [], [], [0]
    li = 0
    for oi, o in enumerate(organisms):
        for g, ln in loci[o]:
            base = rng.integers(0, 4, size=ln, dtype=np.int64)
            seen = set()
            v = 0
            while v < apl[li]:
                s = base.copy()
                k = int(rng.poisson(subs_lambda)) if v > 0 else 0
                if k:
                    p = rng.choice(ln, size=min(k, ln), replace=False)
                    s[p] = (s[p] + rng.integers(1, 4, size=p.shape[0])) % 4
                key = s.tobytes()
                if key in seen:
                    continue  # alleles must be distinct sequences
                seen.add(key)
                v += 1
                row_org.append(oi)
                row_locus.append(li)
                row_variant.append(v)
                seqs.append(_ACGT[s])
            locus_names.append((o, g))
            locus_row0.append(len(row_org))
            li += 1
    seq_off = np.zeros(len(seqs) + 1, dtype=np.int64)
    seq_off[1:] = np.cumsum([len(s) for s in seqs])
    profiles = {}
    li0 = 0
    for o in organisms:
        nl = len(loci[o])
        p = rng.integers(1, np.asarray(apl[li0:li0 + nl]) + 1, size=(n_profiles, nl)).astype(np.int32)
        li0 += nl
        # distinct tuples (duplicates would make defineProfile ambiguous)
        _, idx = np.unique(p, axis=0, return_index=True)
        profiles[o] = p[np.sort(idx)]
    return SynthDB(list(organisms), loci, np.asarray(row_org, np.int32), np.asarray(row_locus, np.int32),
                   np.asarray(row_variant, np.int32), seq_off, np.concatenate(seqs), locus_names,
                   np.asarray(locus_row0, np.int64), profiles)


# ----------------------------------------------------------------------------------------------
# Alignment records
# ----------------------------------------------------------------------------------------------

@dataclass
class AlnTable:
    """Flat, BAM-equivalent alignment records (one row per BAM record), numpy on the host.

    Reads are fixed length L inside one table (synthetic); ragged CIGARs are (cig_off, cig_ops) with the BAM
    encoding `len<<4 | op` (MIDNSHP=X -> 0..8).  `seq` holds ASCII bases as stored in the BAM (reference
    orientation), `qual` phred values.  Aux fields follow bowtie2's order AS,XS,XN,XM,XO,XG,NM,YT; `has_xs`
    False drops XS:i from a record (H4 fixture: the 4th aux field is then XO, `metamlst.py:109-110`).
    """

    ref_names: List[str]
    ref_lens: np.ndarray  # int32 [n_ref]
    tid: np.ndarray  # int32 [n]
    pos: np.ndarray  # int32 [n] 0-based
    flag: np.ndarray  # uint16 [n]
    qname_id: np.ndarray  # int64 [n]  QNAME = "r%d"
    cig_off: np.ndarray  # int64 [n+1]
    cig_ops: np.ndarray  # uint32
    read_len: int
    seq: np.ndarray  # uint8 [n, L] ASCII
    qual: np.ndarray  # uint8 [n, L]
    AS: np.ndarray  # int32
    XS: np.ndarray  # int32
    has_xs: np.ndarray  # bool
    XN: np.ndarray
    XM: np.ndarray
    XO: np.ndarray
    XG: np.ndarray
    NM: np.ndarray
    truth: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.tid.shape[0])

    def take(self, idx: np.ndarray) -> "AlnTable":
        idx = np.asarray(idx, dtype=np.int64)
        ncig = (self.cig_off[1:] - self.cig_off[:-1])[idx]
        new_off = np.zeros(idx.shape[0] + 1, dtype=np.int64)
        new_off[1:] = np.cumsum(ncig)
        src = np.repeat(self.cig_off[:-1][idx] - new_off[:-1], ncig) + np.arange(int(new_off[-1]), dtype=np.int64)
        return AlnTable(self.ref_names, self.ref_lens, self.tid[idx], self.pos[idx], self.flag[idx], self.qname_id[idx],
                        new_off, self.cig_ops[src], self.read_len, self.seq[idx], self.qual[idx], self.AS[idx],
                        self.XS[idx], self.has_xs[idx], self.XN[idx], self.XM[idx], self.XO[idx], self.XG[idx],
                        self.NM[idx], self.truth)

    def coord_order(self) -> np.ndarray:
        """`samtools sort` order: stable by (tid, pos, reverse-strand) -- restated from samtools bam_sort.c
        (not in /root/reference; `metaMLST_functions.py:244` is the call site)."""
        rev = (self.flag >> 4) & 1
        key = (self.tid.astype(np.int64) << 33) | ((self.pos.astype(np.int64) + 1) << 1) | rev
        return np.argsort(key, kind="stable")

    def sorted_by_coord(self) -> "AlnTable":
        return self.take(self.coord_order())


def _nearest_alleles(db: SynthDB, k: int, device: str = "cpu") -> np.ndarray:
    """For every row the k nearest other rows of its locus by Hamming distance (ties -> lower row)."""
    out = np.zeros((db.n_rows, k), dtype=np.int32)
    dev = torch.device(device)
    for li in range(len(db.locus_names)):
        r0, r1 = int(db.locus_row0[li]), int(db.locus_row0[li + 1])
        ln = int(db.seq_off[r0 + 1] - db.seq_off[r0])
        m = torch.from_numpy(db.seq[db.seq_off[r0]:db.seq_off[r1]].reshape(r1 - r0, ln)).to(dev)
        a = r1 - r0
        d = torch.zeros((a, a), dtype=torch.int32, device=dev)
        for c in range(0, a, 64):
            d[c:c + 64] = (m[c:c + 64, None, :] != m[None, :, :]).sum(-1).to(torch.int32)
        d.fill_diagonal_(1 << 30)
        kk = min(k, a - 1)
        nn = torch.sort(d, dim=1, stable=True).indices[:, :kk].cpu().numpy()
        out[r0:r1, :kk] = nn + r0
        if kk < k:
            out[r0:r1, kk:] = out[r0:r1, :1] if kk else np.arange(r0, r1)[:, None]
    return out


def _cached_nearest(db: SynthDB, k: int, device: str) -> np.ndarray:
    cache = db.__dict__.setdefault("_nearest_cache", {})
    if k not in cache:
        cache[k] = _nearest_alleles(db, k, device)
    return cache[k]


def gen_core(db: SynthDB, n_reads: int, read_len: int = 100, seed: int = 1001, K: int = 4,
             org_props: Optional[Sequence[float]] = None, sub_err: float = 0.005, n_frac: float = 0.01,
             novel_loci: int = 2, frac_clip: float = 0.10, frac_indel: float = 0.01,
             device: str = "cpu", strain_seed: Optional[int] = None,
             locus_subset: Optional[Sequence[int]] = None) -> dict:
    """Reads drawn uniformly from one sample strain per organism (one ST, `novel_loci` loci carrying 1-3 novel
    SNPs); K records per read (true allele flag 0/16, K-1 nearest alleles flag 256/272); AS = 2*#M - 6*XM -
    sum(5+3*gaplen).  Returns torch tensors on `device`: per read (bases, qual, start, rtype, a_split) and per
    record [n_reads, K] (rows, AS, XS, xm, flag).  `make_sample` turns them into an AlnTable (host),
    `metamlst_b200.devpack.pack_core` into the packed streams directly on the GPU (bench)."""
    # the sample strain depends on strain_seed only, so a big sample can be generated in chunks of reads
    rng = np.random.Generator(np.random.PCG64(seed if strain_seed is None else strain_seed))
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = read_len
    n_org = len(db.organisms)
    props = np.asarray(org_props if org_props is not None else [1.0 / n_org] * n_org, dtype=np.float64)
    props = props / props.sum()
    lens_all = db.row_len()

    # --- sample strain: per global locus the base allele row + novel substitutions -> strain genome
    strain_row = np.zeros(len(db.locus_names), dtype=np.int64)
    strain_seqs = []
    truth = {"st": {}, "novel": {}}
    li0 = 0
    for o in db.organisms:
        nl = len(db.loci[o])
        st = int(rng.integers(0, db.profiles[o].shape[0]))
        truth["st"][o] = st + 1
        novel = set(rng.choice(nl, size=min(novel_loci, nl), replace=False).tolist())
        for j in range(nl):
            row = int(db.locus_row0[li0 + j]) + int(db.profiles[o][st, j]) - 1
            strain_row[li0 + j] = row
            s = db.seq[db.seq_off[row]:db.seq_off[row + 1]].copy()
            if j in novel:
                k = int(rng.integers(1, 4))
                p = rng.choice(len(s), size=k, replace=False)
                code = np.searchsorted(_ACGT, s[p])
                s[p] = _ACGT[(code + rng.integers(1, 4, size=k)) % 4]
                truth["novel"][db.locus_names[li0 + j]] = sorted(p.tolist())
            strain_seqs.append(s)
        li0 += nl
    n_loci = len(db.locus_names)
    locus_len = np.array([len(s) for s in strain_seqs], dtype=np.int64)
    locus_goff = np.zeros(n_loci + 1, dtype=np.int64)
    locus_goff[1:] = np.cumsum(locus_len)
    genome = torch.from_numpy(np.concatenate(strain_seqs)).to(dev)
    allseq = torch.from_numpy(db.seq).to(dev)
    row_off_t = torch.from_numpy(db.seq_off).to(dev)
    nearest = torch.from_numpy(_cached_nearest(db, max(K - 1, 1), device)).to(dev)

    # --- per read: locus (organism by props, locus by length), type, start
    w = np.zeros(n_loci, dtype=np.float64)
    li0 = 0
    for oi, o in enumerate(db.organisms):
        nl = len(db.loci[o])
        ll = locus_len[li0:li0 + nl].astype(np.float64)
        w[li0:li0 + nl] = props[oi] * ll / ll.sum()
        li0 += nl
    if locus_subset is not None:  # multi-GPU shards own disjoint locus sets (contig-aligned shards)
        keep = np.zeros(n_loci, dtype=bool)
        keep[np.asarray(list(locus_subset), dtype=np.int64)] = True
        w = np.where(keep, w, 0.0)
    wt = torch.from_numpy(w / w.sum()).to(dev)
    locus = torch.multinomial(wt, n_reads, replacement=True, generator=g)
    u = torch.rand(n_reads, generator=g, device=dev)
    # type 0 simple LM, 1 clipped 5S(L-10)M5S, 2 insertion aM1IbM, 3 deletion aM1DbM
    rtype = torch.zeros(n_reads, dtype=torch.int64, device=dev)
    rtype[u < frac_clip + frac_indel] = 1
    rtype[u < frac_indel] = 2
    rtype[u < frac_indel / 2] = 3
    span = torch.full((n_reads,), L, dtype=torch.int64, device=dev)
    span[rtype == 1] = L - 10
    span[rtype == 2] = L - 1
    span[rtype == 3] = L + 1
    llen = torch.from_numpy(locus_len).to(dev)[locus]
    room = (llen - span + 1).clamp(min=1)
    start = (torch.rand(n_reads, generator=g, device=dev) * room).long().clamp(max=(room - 1))
    a_split = 10 + (torch.rand(n_reads, generator=g, device=dev) * (L - 21)).long()  # indel offset in the read

    # --- reference index of every read base (-1: inserted / soft-clipped base)
    j = torch.arange(L, device=dev)[None, :]
    st_ = start[:, None]
    a_ = a_split[:, None]
    t_ = rtype[:, None]
    ref0 = st_ + j
    refidx = ref0.clone()
    refidx = torch.where(t_ == 1, torch.where((j >= 5) & (j < L - 5), st_ + j - 5, torch.full_like(ref0, -1)), refidx)
    refidx = torch.where(t_ == 2, torch.where(j < a_, ref0, torch.where(j == a_, torch.full_like(ref0, -1), ref0 - 1)), refidx)
    refidx = torch.where(t_ == 3, torch.where(j < a_, ref0, ref0 + 1), refidx)
    aligned = refidx >= 0
    goff = torch.from_numpy(locus_goff).to(dev)[locus][:, None]
    acgt_t = torch.from_numpy(_ACGT.copy()).to(dev)
    bases = genome[(goff + refidx.clamp(min=0))]
    rnd = acgt_t[torch.randint(0, 4, (n_reads, L), generator=g, device=dev)]
    bases = torch.where(aligned, bases, rnd)
    # substitution errors: replace by a *different* base
    err = torch.rand((n_reads, L), generator=g, device=dev) < sub_err
    code = torch.zeros_like(bases, dtype=torch.int64)
    code[bases == 67] = 1
    code[bases == 71] = 2
    code[bases == 84] = 3
    sub = acgt_t[(code + torch.randint(1, 4, (n_reads, L), generator=g, device=dev)) % 4]
    bases = torch.where(err, sub, bases)
    isn = torch.rand((n_reads, L), generator=g, device=dev) < n_frac
    bases = torch.where(isn, torch.full_like(bases, 78), bases)
    # qualities: 85 % in 30-40, 10 % in 20-29, 5 % < 20
    uq = torch.rand((n_reads, L), generator=g, device=dev)
    q_hi = torch.randint(30, 41, (n_reads, L), generator=g, device=dev)
    q_mid = torch.randint(20, 30, (n_reads, L), generator=g, device=dev)
    q_lo = torch.randint(2, 20, (n_reads, L), generator=g, device=dev)
    qual = torch.where(uq < 0.85, q_hi, torch.where(uq < 0.95, q_mid, q_lo)).to(torch.uint8)

    # --- K records per read
    true_row = torch.from_numpy(strain_row).to(dev)[locus]
    rows = torch.cat([true_row[:, None], nearest[true_row][:, :K - 1].long()], dim=1) if K > 1 else true_row[:, None]
    rows = rows[:, :K]
    n_m = aligned.sum(1)
    gap = ((rtype == 2) | (rtype == 3)).long()
    xm = torch.zeros((n_reads, K), dtype=torch.int64, device=dev)
    ridx = refidx.clamp(min=0)
    for k in range(K):
        ab = allseq[row_off_t[rows[:, k]][:, None] + ridx]
        xm[:, k] = ((ab != bases) & aligned).sum(1)
    AS = 2 * n_m[:, None] - 6 * xm - gap[:, None] * 8
    strand = (torch.rand(n_reads, generator=g, device=dev) < 0.5).long() * 16
    flag = strand[:, None] + torch.tensor([0] + [256] * (K - 1), device=dev)[None, :]
    # XS = best other alignment score of the read
    if K > 1:
        top2 = torch.topk(AS, 2, dim=1).values
        XS = torch.where(AS == top2[:, :1], top2[:, 1:2].expand(-1, K), top2[:, :1].expand(-1, K))
    else:
        XS = AS.clone()

    return dict(L=L, K=K, n_reads=n_reads, bases=bases, qual=qual, start=start, rtype=rtype, a_split=a_split,
                rows=rows, AS=AS, XS=XS, xm=xm, flag=flag, gap=gap, truth=truth, strain_row=strain_row,
                strain_seqs=strain_seqs, lens_all=lens_all)


def make_sample(db: SynthDB, n_reads: int, read_len: int = 100, seed: int = 1001, K: int = 4, device: str = "cpu",
                **kw) -> AlnTable:
    """AlnTable (host, BAM-equivalent, name-grouped = bowtie2 order) of gen_core's reads; aux order
    AS,XS,XN,XM,XO,XG,NM,YT."""
    core = gen_core(db, n_reads, read_len, seed, K, device=device, **kw)
    L, K = core["L"], core["K"]
    dev = core["bases"].device
    bases, qual, start, rtype, a_split = core["bases"], core["qual"], core["start"], core["rtype"], core["a_split"]
    rows, AS, XS, xm, flag, gap, truth = core["rows"], core["AS"], core["XS"], core["xm"], core["flag"], core["gap"], core["truth"]
    strain_row, strain_seqs, lens_all = core["strain_row"], core["strain_seqs"], core["lens_all"]

    def rep(x):  # [n_reads] -> [n_reads*K]
        return x[:, None].expand(-1, K).reshape(-1)

    n = n_reads * K
    tid = rows.reshape(-1)
    pos = rep(start)
    # CIGARs (same for the K records of a read)
    M, I, D, S = 0, 1, 2, 4
    ncig = torch.ones(n_reads, dtype=torch.int64, device=dev)
    ncig[rtype != 0] = 3
    cig = torch.zeros((n_reads, 3), dtype=torch.int64, device=dev)
    cig[:, 0] = (L << 4) | M
    m1 = rtype == 1
    cig[m1] = torch.tensor([(5 << 4) | S, ((L - 10) << 4) | M, (5 << 4) | S], device=dev)
    m2 = rtype == 2
    cig[m2, 0] = (a_split[m2] << 4) | M
    cig[m2, 1] = (1 << 4) | I
    cig[m2, 2] = ((L - 1 - a_split[m2]) << 4) | M
    m3 = rtype == 3
    cig[m3, 0] = (a_split[m3] << 4) | M
    cig[m3, 1] = (1 << 4) | D
    cig[m3, 2] = ((L - a_split[m3]) << 4) | M
    cigK = cig[:, None, :].expand(-1, K, -1).reshape(n, 3)
    ncigK = rep(ncig)
    keep = torch.arange(3, device=dev)[None, :] < ncigK[:, None]
    cig_ops = cigK[keep]
    cig_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    cig_off[1:] = torch.cumsum(ncigK, 0)

    def npy(x, dt):
        return x.detach().cpu().numpy().astype(dt, copy=False)

    seqK = bases[:, None, :].expand(-1, K, -1).reshape(n, L)
    qualK = qual[:, None, :].expand(-1, K, -1).reshape(n, L)
    gapK = rep(gap)
    tab = AlnTable(
        ref_names=db.ref_names(), ref_lens=lens_all.astype(np.int32),
        tid=npy(tid, np.int32), pos=npy(pos, np.int32), flag=npy(flag.reshape(-1), np.uint16),
        qname_id=npy(rep(torch.arange(n_reads, device=dev)), np.int64),
        cig_off=npy(cig_off, np.int64), cig_ops=npy(cig_ops, np.uint32), read_len=L,
        seq=npy(seqK, np.uint8), qual=npy(qualK, np.uint8),
        AS=npy(AS.reshape(-1), np.int32), XS=npy(XS.reshape(-1), np.int32), has_xs=np.ones(n, dtype=bool),
        XN=np.zeros(n, np.int32), XM=npy(xm.reshape(-1), np.int32), XO=npy(gapK, np.int32), XG=npy(gapK, np.int32),
        NM=npy(xm.reshape(-1) + gapK, np.int32), truth=truth)
    tab.truth["strain_row"] = strain_row
    tab.truth["strain_seqs"] = [s.tobytes().decode() for s in strain_seqs]
    return tab


# ----------------------------------------------------------------------------------------------
# BAM bytes of a synthetic sample, fast (bench / test plumbing: nothing here is on a measured path)
# ----------------------------------------------------------------------------------------------
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bam_record_bytes(db: SynthDB, cores: List[dict], order: str = "name"):
    """The BAM records of gen_core chunks as one uint8 torch tensor + the record end offsets (int64), built with tensor ops on the
    cores' device.  order="name": bowtie2's output order (the K alignments of a read are adjacent); "coord": `samtools sort`
    order.  Fixed-width fields: QNAME "r%09d", aux AS:s XS:s XN:C XM:C XO:C XG:C NM:C YT:Z:UU (bowtie2's order, H4)."""
    dev = cores[0]["bases"].device
    K, L = cores[0]["K"], cores[0]["L"]
    bases = torch.cat([c["bases"] for c in cores]); qual = torch.cat([c["qual"] for c in cores])
    rtype = torch.cat([c["rtype"] for c in cores]); a = torch.cat([c["a_split"] for c in cores]); start = torch.cat([c["start"] for c in cores])
    rows = torch.cat([c["rows"] for c in cores]); AS = torch.cat([c["AS"] for c in cores]); xm = torch.cat([c["xm"] for c in cores])
    flag = torch.cat([c["flag"] for c in cores])
    XS = torch.cat([c["XS"] for c in cores]) if "XS" in cores[0] else AS
    n_reads = bases.shape[0]
    n = n_reads * K
    read_of = torch.arange(n, device=dev) // K
    tid = rows.reshape(-1); pos = start[read_of]; fl = flag.reshape(-1); as_ = AS.reshape(-1); xs_ = XS.reshape(-1); xm_ = xm.reshape(-1)
    if order == "coord":
        key = (tid << 33) | ((pos + 1) << 1) | ((fl >> 4) & 1)
        perm = torch.sort(key, stable=True).indices
        read_of, tid, pos, fl, as_, xs_, xm_ = read_of[perm], tid[perm], pos[perm], fl[perm], as_[perm], xs_[perm], xm_[perm]
    rt = rtype[read_of]
    ncig = torch.where(rt == 0, 1, 3)
    span = torch.full((n,), L, dtype=torch.int64, device=dev)
    span[rt == 1] = L - 10; span[rt == 2] = L - 1; span[rt == 3] = L + 1
    gap = ((rt == 2) | (rt == 3)).long()
    # per-read pieces shared by the K records: 4-bit packed SEQ, QUAL, CIGAR words
    code = torch.full_like(bases, 15)
    for ch, v in ((65, 1), (67, 2), (71, 4), (84, 8)):
        code[bases == ch] = v
    if L % 2:
        code = torch.cat([code, torch.zeros((n_reads, 1), dtype=code.dtype, device=dev)], dim=1)
    seq4 = ((code[:, 0::2] << 4) | code[:, 1::2]).to(torch.uint8)
    M, I, D, S = 0, 1, 2, 4
    cig = torch.zeros((n_reads, 3), dtype=torch.int64, device=dev)
    cig[:, 0] = (L << 4) | M
    m1, m2, m3 = rtype == 1, rtype == 2, rtype == 3
    cig[m1] = torch.tensor([(5 << 4) | S, ((L - 10) << 4) | M, (5 << 4) | S], device=dev)
    cig[m2, 0] = (a[m2] << 4) | M; cig[m2, 1] = (1 << 4) | I; cig[m2, 2] = ((L - 1 - a[m2]) << 4) | M
    cig[m3, 0] = (a[m3] << 4) | M; cig[m3, 1] = (1 << 4) | D; cig[m3, 2] = ((L - a[m3]) << 4) | M

    def le(x, nbytes):  # little-endian bytes of an integer tensor -> [n, nbytes] uint8
        x = x.to(torch.int64)
        return torch.stack([((x >> (8 * i)) & 0xFF) for i in range(nbytes)], dim=1).to(torch.uint8)

    # UCSC binning of [pos, pos + span)
    beg, end = pos, pos + span - 1
    binv = torch.zeros(n, dtype=torch.int64, device=dev)
    done = torch.zeros(n, dtype=torch.bool, device=dev)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = (~done) & ((beg >> shift) == (end >> shift))
        binv[hit] = base + (beg[hit] >> shift)
        done |= hit
    name = torch.zeros((n, 11), dtype=torch.uint8, device=dev)
    name[:, 0] = ord("r")
    rid = read_of.clone()
    for d in range(9, 0, -1):
        name[:, d] = (48 + rid % 10).to(torch.uint8)
        rid = rid // 10
    aux = torch.cat([torch.tensor(list(b"ASs"), dtype=torch.uint8, device=dev).expand(n, 3), le(as_ & 0xFFFF, 2),
                     torch.tensor(list(b"XSs"), dtype=torch.uint8, device=dev).expand(n, 3), le(xs_ & 0xFFFF, 2),
                     torch.tensor(list(b"XNC"), dtype=torch.uint8, device=dev).expand(n, 3), le(torch.zeros_like(xm_), 1),
                     torch.tensor(list(b"XMC"), dtype=torch.uint8, device=dev).expand(n, 3), le(xm_.clamp(0, 255), 1),
                     torch.tensor(list(b"XOC"), dtype=torch.uint8, device=dev).expand(n, 3), le(gap, 1),
                     torch.tensor(list(b"XGC"), dtype=torch.uint8, device=dev).expand(n, 3), le(gap, 1),
                     torch.tensor(list(b"NMC"), dtype=torch.uint8, device=dev).expand(n, 3), le((xm_ + gap).clamp(0, 255), 1),
                     torch.tensor(list(b"YTZUU\0"), dtype=torch.uint8, device=dev).expand(n, 6)], dim=1)
    nseq = (L + 1) // 2
    fixed_tail = nseq + L + aux.shape[1]
    size = 4 + 32 + 11 + 4 * ncig + fixed_tail  # bytes of the record including its block_size field
    head = torch.cat([le(size - 4, 4), le(tid, 4), le(pos, 4), le(torch.full_like(tid, 11), 1), le(torch.full_like(tid, 42), 1), le(binv, 2),
                      le(ncig, 2), le(fl, 2), le(torch.full_like(tid, L), 4), le(torch.full_like(tid, 0xFFFFFFFF), 4), le(torch.full_like(tid, 0xFFFFFFFF), 4),
                      le(torch.zeros_like(tid), 4), name], dim=1)  # 36 + 11 bytes
    ends = torch.cumsum(size, 0)
    off = ends - size
    out = torch.empty(int(ends[-1]) if n else 0, dtype=torch.uint8, device=dev)
    chunk = 1 << 20
    for c0 in range(0, n, chunk):
        sl = slice(c0, min(n, c0 + chunk))
        ro = read_of[sl]
        for g in (1, 3):
            sel = torch.nonzero(ncig[sl] == g)[:, 0]
            if sel.numel() == 0:
                continue
            r = ro[sel]
            rec = torch.cat([head[sl][sel], le(cig[r][:, :g].reshape(-1), 4).reshape(sel.numel(), 4 * g), seq4[r], qual[r], aux[sl][sel]], dim=1)
            dest = off[sl][sel][:, None] + torch.arange(rec.shape[1], device=dev)[None, :]
            out[dest.reshape(-1)] = rec.reshape(-1)
    return out, ends


def write_bam_fast(db: SynthDB, cores: List[dict], order: str = "name", align_records: bool = True, level: int = 1, threads: int = 0,
                   block: int = 0xFF00, path: Optional[str] = None) -> np.ndarray:
    """BAM file bytes (numpy uint8) of a synthetic sample: records built with tensor ops, BGZF blocks deflated by a thread pool.
    align_records: BGZF blocks end at record boundaries (what htslib writes); False: cut every `block` bytes wherever that falls
    (what oracle.bamio writes)."""
    import os
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    recs, ends = bam_record_bytes(db, cores, order)
    names, lens = db.ref_names(), db.row_len()
    text = ("@HD\tVN:1.0\tSO:%s\n" % ("coordinate" if order == "coord" else "unsorted") + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in zip(names, lens)) +
            "@PG\tID:bowtie2\tPN:bowtie2\tVN:2.4.4\n").encode()
    head = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", len(names))]
    for n, l in zip(names, lens):
        nb = n.encode() + b"\0"
        head.append(struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(l)))
    head = b"".join(head)
    body = recs.cpu().numpy()
    ends_h = ends.cpu().numpy()
    cuts = [0]
    if align_records and ends_h.size:
        p = 0
        while p < body.size:
            j = int(np.searchsorted(ends_h, p + block, side="right")) - 1
            q = int(ends_h[j]) if j >= 0 and int(ends_h[j]) > p else min(body.size, p + block)
            cuts.append(q)
            p = q
    else:
        cuts = list(range(0, body.size, block)) + [body.size]
    pieces = [head[i:i + block] for i in range(0, len(head), block)]  # the header in its own blocks, as htslib flushes it
    mv = memoryview(body)

    def deflate(data) -> bytes:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(data) + co.flush()
        return (struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, len(comp) + 25) + comp +
                struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))

    spans = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]
    with ThreadPoolExecutor(threads or (os.cpu_count() or 1)) as pool:
        blocks = list(pool.map(lambda ab: deflate(mv[ab[0]:ab[1]]), spans, chunksize=64))
    raw = b"".join([deflate(p) for p in pieces] + blocks + [_BGZF_EOF])
    if path:
        with open(path, "wb") as fh:
            fh.write(raw)
    return np.frombuffer(raw, dtype=np.uint8)
