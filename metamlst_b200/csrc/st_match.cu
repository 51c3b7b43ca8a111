// Rows a10 / a11 of the hot path: the two table look-ups the reference runs as un-indexed SQL scans.
//
// (1) exact-sequence lookup -- sequenceExists / sequenceLocate / sequenceFind (metaMLST_functions.py:168-172, 196-203, 218-222):
//     `SELECT .. FROM alleles WHERE sequence = ? AND bacterium = ?` + fetchone() == the FIRST row (table order) of the organism
//     whose sequence equals the query character for character (SQLite `=` on TEXT is case-sensitive, H10).  That is NOT
//     "zip-Hamming distance 0": the truncating distance is also 0 against any row that is a prefix (or an extension) of the
//     query, so the length must match too.  On the resident 2-bit DB: row_len == q_len and every plane word equal; a sequence
//     holding a character outside upper-case ACGT is flagged (bit 15 of its length, H9) and can only equal another flagged
//     sequence -- those pairs compare their stored bytes.
//       exact_clean_kernel   : thread = DB row (coalesced tile loads, planes of one row in registers word by word), the block's
//                              queries staged in shared memory; almost every pair is rejected by the 16-bit length or the first word
//       exact_flagged_kernel : CTA = flagged query, threads = flagged rows
//     first_row[q] = min(key of row) by atomicMin, key = row_key[row] (e.g. the position in table order when the resident rows are
//     grouped differently) or the row index itself; caller presets 0xFFFFFFFF.
//
// (2) ST assignment -- defineProfile (metaMLST_functions.py:205-216): among the rows of `profiles` whose alleleCode is one of the
//     query's allele rows, count per profileCode; return the profiles whose count equals the maximum, with that count.  The
//     `profiles` table is held grouped by profile (prof_start[n_st+1] into prof_allele[]), profiles ascending by code (the
//     order SQLite's GROUP BY emits them in), so a (query, profile) pair is one thread, no atomics except the per-query max:
//       st_count_kernel : grid (profiles / 256, queries): count[q][p] = #{rows of p whose allele is in the query set};
//                         atomicMax(best[q], count)
//       st_emit_kernel  : CTA = query: ordered compaction (block scan) of the profiles with count == best[q] > 0
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kStageQ = 64;   // queries staged per pass in exact_clean_kernel

struct ExactArgs {
    const uint32_t* db_hi; const uint32_t* db_lo; const uint16_t* row_len; uint32_t n_rows; uint32_t W;
    const uint32_t* q_hi; const uint32_t* q_lo; const uint16_t* q_len; uint32_t n_q;
    const uint32_t* blocks; uint32_t row_index_base; const uint32_t* row_key;
    const uint32_t* xr_ids; const uint8_t* xr_bytes; uint32_t n_xr;
    const uint32_t* xq_ids; const uint8_t* xq_bytes; uint32_t n_xq;
    uint32_t* first_row;
};

__global__ void __launch_bounds__(kThreads) exact_clean_kernel(const ExactArgs a) {
    const uint32_t* blk = a.blocks + 4 * blockIdx.y;
    const uint32_t q_begin = blk[0], q_end = blk[1], r_begin = blk[2], r_end = blk[3];
    const uint32_t row = (r_begin & ~31u) + blockIdx.x * kThreads + threadIdx.x;   // tile-aligned so a warp reads one tile
    if ((r_begin & ~31u) + blockIdx.x * kThreads >= r_end) return;
    const bool live = row >= r_begin && row < r_end && row < a.n_rows;
    const uint32_t W = a.W;
    const uint32_t rlen = live ? a.row_len[row] : 0xffffu;                        // raw u16: a flagged row never equals a clean length
    const size_t tbase = static_cast<size_t>(row >> 5) * W * 32 + (row & 31u);
    const uint32_t h0 = live ? a.db_hi[tbase] : 0u, l0 = live ? a.db_lo[tbase] : 0u;
    __shared__ uint32_t s_h0[kStageQ], s_l0[kStageQ], s_len[kStageQ];
    for (uint32_t q0 = q_begin; q0 < q_end; q0 += kStageQ) {
        const uint32_t nq = min(static_cast<uint32_t>(kStageQ), q_end - q0);
        __syncthreads();
        if (threadIdx.x < nq) {
            const uint32_t q = q0 + threadIdx.x;
            s_len[threadIdx.x] = a.q_len[q];
            s_h0[threadIdx.x] = a.q_hi[static_cast<size_t>(q) * W];
            s_l0[threadIdx.x] = a.q_lo[static_cast<size_t>(q) * W];
        }
        __syncthreads();
        if (!live || (rlen & 0x8000u)) continue;
        for (uint32_t j = 0; j < nq; ++j) {
            if (s_len[j] != rlen || s_h0[j] != h0 || s_l0[j] != l0) continue;
            const uint32_t q = q0 + j;
            bool same = true;
            const uint32_t nw = (rlen + 31u) >> 5;
            for (uint32_t w = 1; w < nw && same; ++w)
                same = a.db_hi[tbase + static_cast<size_t>(w) * 32] == a.q_hi[static_cast<size_t>(q) * W + w] &&
                       a.db_lo[tbase + static_cast<size_t>(w) * 32] == a.q_lo[static_cast<size_t>(q) * W + w];
            if (same) atomicMin(&a.first_row[q], a.row_key ? a.row_key[row] : row + a.row_index_base);
        }
    }
}

// flagged query e (blockIdx.x) of block blockIdx.y against the flagged rows inside the block's row range: byte compare
__global__ void __launch_bounds__(kThreads) exact_flagged_kernel(const ExactArgs a) {
    const uint32_t* blk = a.blocks + 4 * blockIdx.y;
    const uint32_t q_begin = blk[0], q_end = blk[1], r_begin = blk[2], r_end = blk[3];
    const uint32_t q = a.xq_ids[blockIdx.x];
    if (q < q_begin || q >= q_end) return;
    const uint32_t qlen = a.q_len[q] & 0x7fffu;
    const uint8_t* qb = a.xq_bytes + static_cast<size_t>(blockIdx.x) * a.W * 32;
    for (uint32_t e = threadIdx.x; e < a.n_xr; e += kThreads) {
        const uint32_t row = a.xr_ids[e];
        if (row < r_begin || row >= r_end) continue;
        if ((a.row_len[row] & 0x7fffu) != qlen) continue;
        const uint8_t* rb = a.xr_bytes + static_cast<size_t>(e) * a.W * 32;
        bool same = true;
        for (uint32_t i = 0; i < qlen && same; ++i) same = rb[i] == qb[i];
        if (same) atomicMin(&a.first_row[q], a.row_key ? a.row_key[row] : row + a.row_index_base);
    }
}

struct StArgs {
    const uint32_t* prof_start; const uint32_t* prof_allele; uint32_t n_st;
    const uint32_t* q_alleles; const uint32_t* q_n; uint32_t l_max; uint32_t n_q;
    uint32_t* count; uint32_t* best; uint32_t* n_best; uint32_t* out_idx; uint32_t max_out;
};

__global__ void __launch_bounds__(256) st_count_kernel(const StArgs a) {
    const uint32_t q = blockIdx.y;
    extern __shared__ uint32_t s_set[];
    const uint32_t n = min(a.q_n[q], a.l_max);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_set[i] = a.q_alleles[static_cast<size_t>(q) * a.l_max + i];
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (p < a.n_st) {
        for (uint32_t r = a.prof_start[p]; r < a.prof_start[p + 1]; ++r) {
            const uint32_t al = a.prof_allele[r];
            bool in = false;                     // `alleleCode IN (..)` is a SET test: a code listed twice still counts the row once
            for (uint32_t i = 0; i < n; ++i) in |= (s_set[i] == al);
            c += in ? 1u : 0u;
        }
        a.count[static_cast<size_t>(q) * a.n_st + p] = c;
    }
    c = __reduce_max_sync(0xffffffffu, c);
    if ((threadIdx.x & 31u) == 0 && c) atomicMax(&a.best[q], c);
}

__global__ void __launch_bounds__(256) st_emit_kernel(const StArgs a) {
    const uint32_t q = blockIdx.x;
    const uint32_t best = a.best[q];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    if (best == 0) { if (threadIdx.x == 0) a.n_best[q] = 0; return; }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t p0 = 0; p0 < a.n_st; p0 += blockDim.x) {
        const uint32_t p = p0 + threadIdx.x;
        const bool hit = p < a.n_st && a.count[static_cast<size_t>(q) * a.n_st + p] == best;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = s_base;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        const uint32_t slot = before + __popc(bal & ((1u << lane) - 1u));
        if (hit && slot < a.max_out) a.out_idx[static_cast<size_t>(q) * a.max_out + slot] = p;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (uint32_t w = 0; w < blockDim.x / 32; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) a.n_best[q] = s_base;
}

}  // namespace

extern "C" int mmlst_exact_match_dev(const uint32_t* db_hi, const uint32_t* db_lo, const uint16_t* row_len, uint32_t n_rows, uint32_t W,
                                     const uint32_t* q_hi, const uint32_t* q_lo, const uint16_t* q_len, uint32_t n_q,
                                     const uint32_t* blocks, uint32_t n_blocks, uint32_t max_block_rows, uint32_t row_index_base,
                                     const uint32_t* row_key, const uint32_t* xr_ids, const uint8_t* xr_bytes, uint32_t n_xr,
                                     const uint32_t* xq_ids, const uint8_t* xq_bytes, uint32_t n_xq,
                                     uint32_t* first_row, void* stream) {
    if (n_q == 0 || n_blocks == 0 || n_rows == 0) return MMLST_OK;
    if (!db_hi || !db_lo || !row_len || !q_hi || !q_lo || !q_len || !blocks || !first_row) { mmlst_set_error("mmlst_exact_match_dev: null pointer"); return MMLST_E_ARG; }
    if (n_blocks > 65535u) { mmlst_set_error("mmlst_exact_match_dev: more than 65535 blocks per call"); return MMLST_E_ARG; }
    if ((n_xr && (!xr_ids || !xr_bytes)) || (n_xq && (!xq_ids || !xq_bytes))) { mmlst_set_error("mmlst_exact_match_dev: flagged lists without data"); return MMLST_E_ARG; }
    const ExactArgs a{db_hi, db_lo, row_len, n_rows, W, q_hi, q_lo, q_len, n_q, blocks, row_index_base, row_key, xr_ids, xr_bytes, n_xr, xq_ids, xq_bytes, n_xq, first_row};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (max_block_rows == 0 || max_block_rows > n_rows) max_block_rows = n_rows;
    const uint32_t gx = (max_block_rows + 31u + kThreads - 1) / kThreads + 1;   // + the tile-alignment slack of the first row
    exact_clean_kernel<<<dim3(gx, n_blocks), kThreads, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    if (n_xq && n_xr) {
        exact_flagged_kernel<<<dim3(n_xq, n_blocks), kThreads, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
    }
    return MMLST_OK;
}

extern "C" int mmlst_st_match_dev(const uint32_t* prof_start, const uint32_t* prof_allele, uint32_t n_st,
                                  const uint32_t* q_alleles, const uint32_t* q_n, uint32_t l_max, uint32_t n_q,
                                  uint32_t* count, uint32_t* best, uint32_t* n_best, uint32_t* out_idx, uint32_t max_out, void* stream) {
    if (n_q == 0) return MMLST_OK;
    if (!prof_start || !q_alleles || !q_n || !count || !best || !n_best || !out_idx || (n_st && !prof_allele)) { mmlst_set_error("mmlst_st_match_dev: null pointer"); return MMLST_E_ARG; }
    if (n_q > 65535u) { mmlst_set_error("mmlst_st_match_dev: more than 65535 queries per call"); return MMLST_E_ARG; }
    if (l_max == 0 || l_max > 4096u) { mmlst_set_error("mmlst_st_match_dev: l_max must be 1..4096"); return MMLST_E_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const StArgs a{prof_start, prof_allele, n_st, q_alleles, q_n, l_max, n_q, count, best, n_best, out_idx, max_out};
    CUDA_TRY(cudaMemsetAsync(best, 0, sizeof(uint32_t) * n_q, st));
    if (n_st) {
        st_count_kernel<<<dim3((n_st + 255u) / 256u, n_q), 256, l_max * sizeof(uint32_t), st>>>(a);
        CUDA_TRY(cudaGetLastError());
    }
    st_emit_kernel<<<n_q, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}
