/*
 * TEST INFRASTRUCTURE ONLY -- plain scalar C restatement of the hot path's integer arithmetic, for parity tests at
 * sizes the pure-Python oracle (oracle/mlst_oracle.py, pinned against the reference over shims) cannot finish, and
 * for bench.py's cpu_baseline / --impl reference legs.  Never linked into the product.  It consumes UNPACKED
 * records (ASCII bases, phred bytes, BAM CIGAR words) so that nothing of the product's bit-plane packing is shared.
 * tests/test_oracle_c.py checks every function against the Python oracle on small inputs.
 *
 * Reference lines (paths under /root/reference): see each function.  The pileup engine follows htslib
 * bam_plp_push/bam_plp_next (third-party, un-pinned: "parity unpinned" for H1-H3, as in the Python oracle).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* metamlst.py:101-130 + integer half of :133-151 */
void orc_score(uint64_t n, const int32_t* tid, const int32_t* aux0, const int32_t* aux3, const int32_t* qlen,
               const uint32_t* orig_idx, const uint8_t* allow, const uint32_t* locus_of, int minscore, int max_xm,
               int min_read_len, int64_t* sum_as, uint32_t* n_hit, uint32_t* first_idx, uint64_t* counters, uint64_t idx_base) {
    /* idx_base: file-order index of record 0 of this range when orig_idx is NULL (range-split scoring) */
    for (uint64_t i = 0; i < n; ++i) {
        const int32_t t = tid[i];
        if (!allow[t]) continue;                                   /* :114 */
        if (aux0[i] >= minscore && qlen[i] >= min_read_len && aux3[i] <= max_xm) { /* :115 */
            sum_as[t] += aux0[i];
            n_hit[t] += 1;
            const uint32_t oi = orig_idx ? orig_idx[i] : (uint32_t)(idx_base + i);
            if (oi < first_idx[t]) first_idx[t] = oi;
        } else {
            counters[1] += 1;                                      /* :129 */
        }
        counters[0] += 1;                                          /* :130 */
    }
}

/* htslib bam_plp_push / bam_plp_next mempool accounting (H1) for records of ONE contig fetch, coordinate-sorted.
 * A faithful event simulation (not the closed form the product uses). */
typedef struct { int32_t beg, end; } node_t;
int orc_depth_cap_sim(uint64_t n, int32_t tid, const int32_t* pos, const int32_t* reflen, uint32_t maxcnt,
                      uint32_t sentinel_nodes, uint8_t* admitted) {
    node_t* buf = (node_t*)malloc(sizeof(node_t) * (n + 1));
    uint64_t nb = 0;
    uint64_t cnt = sentinel_nodes;
    int64_t iter_tid = 0, iter_pos = 0, max_tid = -1, max_pos = -1;
    for (uint64_t i = 0; i <= n; ++i) {
        /* bam_plp_auto: drain bam_plp_next until it needs more input */
        const int eof = (i == n);
        while (eof || max_tid > iter_tid || (max_tid == iter_tid && max_pos > iter_pos)) {
            if (eof && nb == 0) break;
            uint64_t k = 0;
            for (uint64_t j = 0; j < nb; ++j) {
                if (tid < iter_tid || (tid == iter_tid && buf[j].end <= iter_pos)) { --cnt; continue; } /* mp_free */
                buf[k++] = buf[j];
            }
            nb = k;
            if (nb) {
                if (iter_tid < tid) { iter_tid = tid; iter_pos = buf[0].beg; }
                else if (iter_pos < buf[0].beg) iter_pos = buf[0].beg;
                else ++iter_pos;
            } else ++iter_pos;
        }
        if (eof) break;
        /* bam_plp_push */
        if (iter_tid == tid && iter_pos == pos[i] && cnt > maxcnt) { admitted[i] = 0; continue; }
        admitted[i] = 1;
        if (pos[i] < max_pos && max_tid == tid) { free(buf); return -1; } /* "The input is not sorted" */
        max_tid = tid; max_pos = pos[i];
        const int32_t end = pos[i] + reflen[i];
        if (end > iter_pos || tid > iter_tid) { buf[nb].beg = pos[i]; buf[nb].end = end; ++nb; ++cnt; }
    }
    free(buf);
    return 0;
}

/* cmseq/cmseq.py:527-548 for ONE admitted record: walk its CIGAR, bin every aligned base into counts[len][5] (A,C,G,T,N). */
static void pile_one(int64_t r, const uint32_t* cig, int64_t n_cig, const uint8_t* seq, const uint8_t* qual, int L, int pass,
                     int minqual, int32_t contig_len, uint32_t* counts) {
    int q = 0;
    for (int64_t c = 0; c < n_cig; ++c) {
        const uint32_t op = cig[c] & 0xF, len = cig[c] >> 4;
        if (op == 0 || op == 7 || op == 8) {
            for (uint32_t k = 0; k < len; ++k, ++r, ++q) {
                if (r < 0 || r >= contig_len) continue;
                const int qv = (q < L) ? qual[q] : 0;  /* pysam pileup_base_qual_skip (H3) */
                if (qv < minqual) continue;
                uint8_t b = (q < L) ? seq[q] : 'N';
                if (b >= 'a' && b <= 'z') b -= 32;                       /* .upper(), :537 */
                int bin = 4;
                if (pass) { if (b == 'A') bin = 0; else if (b == 'C') bin = 1; else if (b == 'G') bin = 2; else if (b == 'T') bin = 3; }
                counts[r * 5 + bin] += 1;
            }
        } else if (op == 1 || op == 4) q += len;      /* I, S: query only */
        else if (op == 2 || op == 3) r += len;        /* D, N: is_del / is_refskip are skipped, :535 */
    }
}

/* the admitted records of one contig.  seq/qual: n x L row-major (ASCII / phred); cigar: BAM words at cig_off[i]..cig_off[i+1]. */
void orc_pileup(uint64_t n, const int32_t* pos, const int64_t* cig_off, const uint32_t* cig, const uint8_t* seq,
                const uint8_t* qual, int L, const int32_t* as_named, const int32_t* xm_named, const uint8_t* admitted,
                int minqual, int minscore, int max_xm, int32_t contig_len, uint32_t* counts) {
    for (uint64_t i = 0; i < n; ++i) {
        if (!admitted[i]) continue;
        const int pass = (as_named[i] >= minscore) && (xm_named[i] <= max_xm);  /* metaMLST_functions.py:259 */
        pile_one(pos[i], cig + cig_off[i], cig_off[i + 1] - cig_off[i], seq + i * (uint64_t)L, qual + i * (uint64_t)L, L, pass,
                 minqual, contig_len, counts);
    }
}

/* same, for a table that stores SEQ / QUAL / CIGAR once per READ and K alignment records per read (bowtie2 -k): record i
 * belongs to read read_of[i]; cig3 holds up to 3 BAM CIGAR words per read, n_cig[read] of them used. */
void orc_pileup_reads(uint64_t n, const int32_t* pos, const int64_t* read_of, const uint32_t* cig3, const uint8_t* n_cig,
                      const uint8_t* seq, const uint8_t* qual, int L, const int32_t* as_named, const int32_t* xm_named,
                      const uint8_t* admitted, int minqual, int minscore, int max_xm, int32_t contig_len, uint32_t* counts) {
    for (uint64_t i = 0; i < n; ++i) {
        if (!admitted[i]) continue;
        const int pass = (as_named[i] >= minscore) && (xm_named[i] <= max_xm);
        const uint64_t rd = (uint64_t)read_of[i];
        pile_one(pos[i], cig3 + 3 * rd, n_cig[rd], seq + rd * (uint64_t)L, qual + rd * (uint64_t)L, L, pass, minqual, contig_len, counts);
    }
}

/* cmseq/cmseq.py:551-554, 202-209, 234-237 + metaMLST_functions.py:260-276 */
void orc_consensus(const uint32_t* counts, const uint8_t* dbseq, int32_t len, uint32_t mincov, uint8_t* cons,
                   uint32_t* holes, uint32_t* snps) {
    uint32_t h = 0, s = 0;
    for (int32_t i = 0; i < len; ++i) {
        const uint32_t A = counts[i * 5], C = counts[i * 5 + 1], G = counts[i * 5 + 2], T = counts[i * 5 + 3], N = counts[i * 5 + 4];
        uint8_t call = 'N';
        if (A + C + G + T >= mincov && (A || C || G || T || N)) {
            /* max(sorted(freq), key=freq.get): keys sorted A,C,G,N,T; first maximum wins (H8) */
            const uint32_t v[5] = {A, C, G, N, T};
            const char k[5] = {'A', 'C', 'G', 'N', 'T'};
            int best = 0;
            for (int j = 1; j < 5; ++j) if (v[j] > v[best]) best = j;
            call = (uint8_t)k[best];
        }
        if (call == 'N') { uint8_t d = dbseq[i]; cons[i] = (d >= 'A' && d <= 'Z') ? (uint8_t)(d + 32) : d; ++h; }
        else { cons[i] = call; if (call != dbseq[i]) ++s; }
    }
    *holes = h; *snps = s;
}

/* metaMLST_functions.py:230-234 over rows [r0, r1): min distance, lowest row on ties */
void orc_hamming_min(const uint8_t* q, int32_t qlen, const uint8_t* rows, const int64_t* row_off, uint32_t r0, uint32_t r1,
                     uint32_t* min_dist, uint32_t* argmin) {
    uint32_t best = 0xffffffffu, arg = 0xffffffffu;
    for (uint32_t r = r0; r < r1; ++r) {
        const uint8_t* s = rows + row_off[r];
        const int64_t rl = row_off[r + 1] - row_off[r];
        const int64_t m = qlen < rl ? qlen : rl;   /* zip() truncates to the shorter string (H9) */
        uint32_t d = 0;
        for (int64_t i = 0; i < m; ++i) d += (q[i] != s[i]);
        if (d < best) { best = d; arg = r; }
    }
    *min_dist = best; *argmin = arg;
}
