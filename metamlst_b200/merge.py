"""Cohort merge: allele classification + ST assignment of a folder of `.nfo` files, i.e. what
`metamlst-merge.py FOLDER -d DB [-z ED] [--filter ..] [--meta ..]` does between reading the folder and writing
`merged/<organism>_ST.txt` / `merged/<organism>_report.txt` (metamlst-merge.py:90-341), with the per-character work on
the GPU.

The reference walks the cohort line by line and, for every reconstructed locus, runs (a) an un-indexed full-table SQL
scan for an exact match (`sequenceExists` / `sequenceLocate`, metaMLST_functions.py:168-172,218-222) and (b) for a
sequence not in the DB, a Python per-character loop against every allele of the locus (`sequencesGetAll` + `stringDiff`,
metamlst-merge.py:174-181).  Here the cohort is handled in three phases:

  1. device : parse the folder (host); every distinct reconstructed sequence of an organism goes through ONE exact-match call
              against the organism's resident 2-bit rows (`HammingIndex.exact_first`, csrc/st_match.cu: same length + same
              characters, case-sensitive like SQLite's `=`, H10; lowest rowid wins like `fetchone()`).  Without a GPU context
              (CPU tests that inject the distance function) the same answers come from one pass over `alleles`;
  2. device : every DISTINCT sequence that is not in the DB goes, once, through ONE batched closest-allele search
              (`api.HammingIndex.search`: 2-bit XOR + popcount, non-ACGT letters on the exact path, H9) against the rows
              of its own locus -> min zip-Hamming distance; the reference's early-exit `any(d <= z)` is `min <= z`;
  3. host   : the sequential bookkeeping of the reference (new allele / profile numbering depends on encounter order),
              with every lookup answered from the two tables above; ST assignment (defineProfile, H11) for all lines made of DB
              alleles in ONE batched device call (`api.ProfileIndex`), memoised per label tuple.

Files and screen text are byte-identical to the reference's (tests/test_merge_driver.py against the golden cohort that
the unmodified script produced).  The sequence writers behind `--outseqformat` (metamlst-merge.py:345-494: FASTA/CSV
dumps, no arithmetic) are not part of the hot path and are left to the reference script.
"""
from __future__ import annotations

import os
import sqlite3
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

from . import api

OKBLUE, OKGREEN, WARNING, FAIL, ENDC = "\033[94m", "\033[92m", "\033[93m", "\033[91m", "\033[0m"

Closest = Callable[[str, Sequence[Tuple[str, str]]], List[int]]  # (bacterium, [(gene, sequence)]) -> [min distance]


def read_nfo_folder(folder: str, species_filter: Optional[str] = None) -> Dict[str, list]:
    """metamlst-merge.py:90-107: {organism: [({label: (SEQ upper-cased, confidence, snp pct)}, sample)]} in
    os.listdir / file-line order.  `--filter` is a substring test on the raw argument, as upstream (`:102`)."""
    cel: Dict[str, list] = {}
    for name in os.listdir(folder):
        if name.split(".")[-1] != "nfo":
            continue
        with open(folder + "/" + name, "r") as fh:
            for line in fh:
                cols = line.split()
                organism, sample_name = cols[0], cols[1]
                if species_filter and organism not in species_filter:
                    continue
                loci = {}
                for item in cols[2:]:
                    parts = item.split("::")
                    loci[parts[0]] = (parts[1].upper(), parts[2], parts[3])
                cel.setdefault(organism, []).append((loci, sample_name))
    return cel


@dataclass
class OrganismMerge:
    """State of one organism after classification: what the writers of metamlst-merge.py:253-341 consume."""
    bacterium: str
    label: str
    loci: List[str]                                   # sorted(lastGenes)
    old_profiles: Dict[int, list] = field(default_factory=dict)          # profileCode -> [hits, {gene: alleleVariant}]
    new_profiles: Dict[int, list] = field(default_factory=dict)          # id -> [{gene: (allele, category)}, hits, category]
    isolates: List[Tuple[int, float, str]] = field(default_factory=list)
    new_sequences: Dict[str, List[Tuple[str, str]]] = field(default_factory=dict)  # gene -> [(new label, sequence)]
    n_searched: int = 0                               # distinct novel sequences sent to the GPU search


class CohortMerger:
    """DB connection + resident allele index, reused for every organism of the cohort."""

    def __init__(self, db_path: str, ctx=None, z: Optional[int] = 5, closest: Optional[Closest] = None):
        if not os.path.isfile(db_path):
            raise IOError("Failed to connect to the database: please check your database file!")  # metamlst-merge.py:66-68
        self.db_path = db_path
        self.conn = sqlite3.connect(db_path)
        self.conn.row_factory = sqlite3.Row
        self.ctx = ctx
        self.z = z
        self._closest = closest
        self._index: Dict[str, api.HammingIndex] = {}
        self._profile_memo: Dict[Tuple[str, ...], list] = {}
        self._profiles: Optional[api.ProfileIndex] = None
        self.device_lookups = ctx is not None and closest is None   # rows a10 / a11 on the GPU too (exact match, ST assignment)

    def close(self):
        self.conn.close()

    # -- look-up tables -----------------------------------------------------------------------------------------------
    def organism_label(self, bacterium: str) -> str:
        """db_getOrganisms(conn, bacterium) (metaMLST_functions.py:422-426); KeyError for an organism without profiles."""
        t = {}
        for r in self.conn.execute("SELECT label,organismkey,COUNT(DISTINCT profileCode) AS totalProfiles FROM organisms,profiles "
                                   "WHERE organismkey = bacterium GROUP BY label,organismkey"):
            t[r["organismkey"]] = r["label"] if r["label"] is not None else "(" + r["organismkey"] + ")"
        return t[bacterium]

    def exact_table(self, bacterium: str) -> Dict[str, str]:
        """{sequence: str(alleleVariant) of the first row holding it}: sequenceExists (metaMLST_functions.py:168-172) is
        `seq in table`, sequenceLocate (:218-222) is `table[seq]` -- organism-wide, any gene, case-sensitive."""
        table: Dict[str, str] = {}
        for r in self.conn.execute("SELECT sequence,alleleVariant FROM alleles WHERE bacterium = ? ORDER BY rowid", (bacterium,)):
            if r["sequence"] is not None:  # SQL `=` never matches NULL
                table.setdefault(str(r["sequence"]), str(r["alleleVariant"]))
        return table

    def _organism_index(self, bacterium: str) -> api.HammingIndex:
        idx = self._index.get(bacterium)
        if idx is None:
            idx = self._index[bacterium] = api.HammingIndex.from_sqlite(self.ctx, self.conn, bacterium)
        return idx

    def exact_lookup(self, bacterium: str, sequences: Sequence[str]) -> Dict[str, str]:
        """{sequence: str(alleleVariant)} for those of `sequences` the organism's alleles hold -- the answers `exact_table` gives,
        from ONE device call over the resident 2-bit DB (`api.HammingIndex.exact_first`, csrc/st_match.cu)."""
        idx = self._organism_index(bacterium)
        rng = idx.organism_range(bacterium)
        seqs = list(dict.fromkeys(s for s in sequences if s != ""))
        if rng is None or not seqs:
            return {}
        rows = idx.exact_first(seqs, [rng] * len(seqs))
        return {s: str(idx.rows[int(r)][2]) for s, r in zip(seqs, rows) if int(r) != api.NO_IDX}

    def closest_distances(self, bacterium: str, items: Sequence[Tuple[str, str]]) -> List[int]:
        """min over the alleles of (bacterium, gene) of stringDiff(seq, allele) for every (gene, seq), one device call."""
        if not items:
            return []
        if self._closest is not None:
            return list(self._closest(bacterium, items))
        if self.ctx is None:
            raise RuntimeError("CohortMerger needs a native.Context (GPU) for the closest-allele search; there is no CPU fallback")
        idx = self._organism_index(bacterium)
        # a locus without rows in the DB: the reference's loop body never runs and the allele stays rejected (:173-181)
        have = [i for i, (g, _s) in enumerate(items) if (bacterium, g) in idx.block]
        out = [1 << 30] * len(items)
        if have:
            dist, _rows = idx.search([items[i][1] for i in have], [idx.block[(bacterium, items[i][0])] for i in have])
            for i, d in zip(have, dist):
                out[i] = int(d)
        return out

    def define_profile(self, labels: Sequence[str]):
        key = tuple(labels)
        hit = self._profile_memo.get(key)
        if hit is None:
            self.define_profiles([key])
            hit = self._profile_memo[key]
        return hit

    def define_profiles(self, label_lists: Sequence[Sequence[str]]) -> None:
        """Memoise defineProfile for every list: ONE batched device call (`api.ProfileIndex`, csrc/st_match.cu) with a GPU context,
        else the reference's SQL statement per list."""
        todo = [tuple(l) for l in dict.fromkeys(tuple(l) for l in label_lists) if tuple(l) not in self._profile_memo]
        if not todo:
            return
        if self.device_lookups:
            if self._profiles is None:
                self._profiles = api.ProfileIndex(self.ctx, self.conn)
            for key, res in zip(todo, self._profiles.define_profiles(todo)):
                self._profile_memo[key] = res
        else:
            for key in todo:
                self._profile_memo[key] = api.define_profile(self.conn, list(key))

    # -- classification -----------------------------------------------------------------------------------------------
    def merge_organism(self, bacterium: str, records: list) -> OrganismMerge:
        """metamlst-merge.py:121-239 for one organism."""
        conn, z = self.conn, self.z
        last_gene = dict((r["gene"], 100000) for r in conn.execute(
            "SELECT gene, MAX(alleleVariant) as maxGene FROM alleles WHERE bacterium = ? GROUP BY gene", (bacterium,)))  # :134
        st = OrganismMerge(bacterium, self.organism_label(bacterium), sorted(last_gene.keys()))
        for r in conn.execute("SELECT profileCode,gene,alleleVariant FROM profiles,alleles WHERE alleleCode = alleles.recID AND alleles.bacterium = ?",
                              (bacterium,)):  # :138-140
            st.old_profiles.setdefault(r["profileCode"], [0, {}])[1][r["gene"]] = r["alleleVariant"]
        if self.device_lookups:
            known = self.exact_lookup(bacterium, [seq for loci, _s in records for (seq, _a, _p) in loci.values()])
            # ST assignment of every line made of DB alleles only, batched (the lines that can reach defineProfile, :199-204)
            self.define_profiles([[bacterium + "_" + label.split("_")[1] + "_" + (known[seq] if seq != "" else label.split("_")[2])
                                   for label, (seq, _a, _p) in loci.items()]
                                  for loci, _s in records if all(seq == "" or seq in known for (seq, _a, _p) in loci.values())])
        else:
            known = self.exact_table(bacterium)
        # phase 2: every distinct sequence that is not a DB sequence, searched once (first gene it appears with: a later
        # appearance hits `genesBase` upstream and never reaches the distance test)
        accepted: Dict[str, bool] = {}
        if z is not None:
            todo: Dict[str, str] = {}
            for loci, _sample in records:
                for label, (seq, _acc, _snp) in loci.items():
                    if seq != "" and seq not in known and seq not in todo:
                        todo[seq] = label.split("_")[1]
            items = [(g, s) for s, g in todo.items()]
            for (g, s), d in zip(items, self.closest_distances(bacterium, items)):
                accepted[s] = d <= z
            st.n_searched = len(items)
        # phase 3: the reference's sequential bookkeeping
        new_label_of: Dict[str, str] = {}  # genesBase
        last_profile = 100000
        for loci, sample_name in records:
            line: Dict[str, tuple] = {}
            new_alleles: List[str] = []
            recurrent = False
            acc_sum = 0.0
            for label, (seq, acc, _snp) in loci.items():
                organism, gene, allele = label.split("_")
                acc_sum += float(acc)
                if seq == "" or seq in known:  # :156-161
                    line[gene] = (known[seq] if seq != "" else allele, 0)
                elif seq in new_label_of:      # :163-165
                    line[gene] = (new_label_of[seq].split("_")[2], 2)
                    recurrent = True
                else:                          # :167-195
                    category = 1 if (z is None or accepted[seq]) else 3
                    number = str(last_gene[gene] + 1)
                    last_gene[gene] += 1
                    new_label_of[seq] = organism + "_" + gene + "_" + number
                    line[gene] = (number, category)
                    new_alleles.append(gene)
                    st.new_sequences.setdefault(gene, []).append((new_label_of[seq], seq))
            mean_acc = acc_sum / float(len(loci))
            if not new_alleles:  # :199-224
                if not recurrent:
                    tried = self.define_profile([bacterium + "_" + k + "_" + v[0] for k, v in line.items()])
                    if tried and tried[0][1] == 100:
                        st.old_profiles[tried[0][0]][0] += 1
                        st.isolates.append((tried[0][0], mean_acc, sample_name))
                        continue
                signature = [k + str(v[0]) for k, v in sorted(line.items())]
                found = 0
                for key, (element, _hits, _cat) in st.new_profiles.items():
                    if signature == [k + str(v[0]) for k, v in sorted(element.items())]:
                        found = key  # no break upstream: the LAST equal profile wins
                if found:
                    st.new_profiles[found][1] += 1
                    st.isolates.append((found, mean_acc, sample_name))
                else:
                    last_profile += 1
                    st.new_profiles[last_profile] = [line, 1, 2]
                    st.isolates.append((last_profile, mean_acc, sample_name))
            else:  # :225-238
                last_profile += 1
                cat = 3 if (z is not None and any(c == 3 for (_v, c) in line.values())) else 1
                st.new_profiles[last_profile] = [line, 1, cat]
                if cat != 3:
                    st.isolates.append((last_profile, mean_acc, sample_name))
        return st

    # -- writers (metamlst-merge.py:253-341) --------------------------------------------------------------------------
    @staticmethod
    def st_table(st: OrganismMerge) -> str:
        """merged/<organism>_ST.txt (note upstream's mixed line ends: '\\r\\n' for header and known STs, '\\n' for new)."""
        out = ["ST\t" + "\t".join(st.loci) + "\r\n"]
        for code, (_hits, profile) in st.old_profiles.items():
            out.append(str(code) + "\t" + "\t".join(str(v) for _k, v in sorted(profile.items())) + "\r\n")
        for pid, (profile, _hits, cat) in st.new_profiles.items():
            if cat in (1, 2):
                out.append(str(pid) + "\t" + "\t".join(str(v[0]) for _k, v in sorted(profile.items())) + "\n")
        return "".join(out)

    @staticmethod
    def _coloured(profile: dict) -> str:
        def paint(v):
            return (OKBLUE if v[1] == 1 else OKGREEN if v[1] == 2 else FAIL) + str(v[0]) + ENDC if v[1] in (1, 2, 3) else str(v[0])
        return "\t".join(paint(v) for _k, v in sorted(profile.items()))

    def screen_text(self, st: OrganismMerge) -> str:
        """What the script prints for one organism (metamlst-merge.py:114-116, 254-292)."""
        hdr = "ST\t" + "\t".join(st.loci) + "\tHits"
        out = [OKBLUE + "+" + ("-" * 78) + "+" + ENDC, OKBLUE + "|" + ENDC + str(st.label).center(78) + OKBLUE + "|" + ENDC,
               OKBLUE + "+" + ("-" * 78) + "+" + ENDC, "KNOWN MLST profiles found:\n" + hdr]
        for code, (hits, profile) in st.old_profiles.items():
            if hits > 0:
                out.append(FAIL + str(code) + ENDC + "\t" + "\t".join(str(v) for _k, v in sorted(profile.items())) + "\t" + str(hits))
        out.append("\n\nNEW MLST profiles found:\n" + hdr)
        for pid, (profile, hits, cat) in st.new_profiles.items():
            if cat in (1, 2):
                out.append((WARNING if cat == 1 else OKGREEN) + str(pid) + ENDC + "\t" + self._coloured(profile) + "\t" + str(hits))
        out.append("\n\nREJECTED NEW MLST profiles, as they have > SNPs than max-threshold (-z " + str(self.z) + ")\n" + hdr)
        for pid, (profile, hits, cat) in st.new_profiles.items():
            if cat == 3:
                out.append(str(pid) + "\t" + self._coloured(profile) + "\t" + str(hits))
        out.append("")
        text = "\n".join(out) + "\n"
        return text + "Outputing results".ljust(66) + (ENDC + "[ - " + "...".center(5) + " - ]" + ENDC).ljust(14) + "\r\n"

    @staticmethod
    def report(st: OrganismMerge, meta_path: Optional[str] = None, id_field: int = 0) -> str:
        """merged/<organism>_report.txt (metamlst-merge.py:296-339)."""
        keys: List[str] = []
        ident: Dict[str, Dict[str, str]] = {}
        if meta_path:
            first = True
            with open(meta_path) as fh:
                for line in fh:
                    if line == "":
                        continue
                    if first:
                        first = False
                        keys = [str(x).strip() for x in line.split("\t")]
                    else:
                        cols = line.strip().split("\t")
                        if len(cols) == len(keys):
                            ident[cols[id_field]] = dict((keys[i], cols[i]) for i in range(len(keys)))
        out = ["ST\tConfidence\t" + "\t".join(keys) + "\n"]
        for code, mean_acc, sample_name in st.isolates:
            if sample_name.endswith(".fna"):
                sample_name = sample_name.split(".")[0]
            if sample_name in ident:
                out.append(str(code) + "\t" + str(round(mean_acc, 2)) + "\t" + "\t".join(ident[sample_name][k] for k in keys) + "\n")
            else:
                out.append(str(code) + "\t" + str(round(mean_acc, 2)) + "\t" + str(sample_name) + "\n")
        return "".join(out)


def merge_folder(folder: str, db_path: str, ctx=None, z: Optional[int] = 5, species_filter: Optional[str] = None,
                 meta_path: Optional[str] = None, id_field: int = 0, closest: Optional[Closest] = None,
                 write: bool = True) -> Tuple[Dict[str, OrganismMerge], str]:
    """`metamlst-merge.py folder -d db_path [-z z] [--filter f] [--meta m --idField i]` up to the report files.
    Returns ({organism: OrganismMerge}, screen text)."""
    merger = CohortMerger(db_path, ctx=ctx, z=z, closest=closest)
    try:
        if write and not os.path.isdir(folder + "/merged"):
            os.makedirs(folder + "/merged")  # metamlst-merge.py:81
        cel = read_nfo_folder(folder, species_filter)
        screen = [OKBLUE + "MetaMLST Database file: " + ENDC + os.path.basename(db_path) + "\n\n"]
        states: Dict[str, OrganismMerge] = {}
        for bacterium, records in cel.items():
            st = states[bacterium] = merger.merge_organism(bacterium, records)
            screen.append(merger.screen_text(st))
            if write:
                with open(folder + "/merged/" + bacterium + "_ST.txt", "w", newline="") as fh:
                    fh.write(merger.st_table(st))
                with open(folder + "/merged/" + bacterium + "_report.txt", "w", newline="") as fh:
                    fh.write(merger.report(st, meta_path, id_field))
        screen.append("Colour Legend:\n" + "-" * 80 + "\n")
        screen.append("Alleles:" + "\t" + "[Known]" + "\t" + OKBLUE + "[NEW]" + ENDC + "\t" + OKGREEN + "[NEW-RECURRING]" + ENDC + "\n")
        screen.append("Profiles:" + "\t" + FAIL + "[Known]" + ENDC + "\t" + WARNING + "[NEW]" + ENDC + "\t" + OKGREEN + "[NEW*]" + ENDC + "\n")
        screen.append("New* profiles are composed by Known and Recurring alleles only\n" + "-" * 80 + "\nCompleted! Have a nice day.\n")
        return states, "".join(screen)
    finally:
        merger.close()
