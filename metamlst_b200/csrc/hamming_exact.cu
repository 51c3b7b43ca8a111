// Stage 3, exception path (hazard H9): stringDiff compares CHARACTERS (metaMLST_functions.py:230-234), so a DB row or a
// query holding anything but upper-case A/C/G/T (IUPAC codes, 'N', lower case) cannot live in the 2-bit planes alone.
// Such sequences are flagged (bit 15 of their length), skipped by hamming_min_kernel, and compared here, exactly:
//   an exceptional character never equals a clean one, so for a pair (row r, query q) over m = min(len) columns
//     distance = popc(mismatch(2-bit planes) & ~(X_r | X_q))        both clean: the usual XOR/OR word
//              + popc(X_r ^ X_q)                                     exactly one side exceptional: always a mismatch
//              + #{ i in X_r & X_q : byte_r[i] != byte_q[i] }        both exceptional: compare the stored bytes
//   X = per-sequence bit-plane of exceptional columns, kept only for the flagged sequences together with a dense ASCII
//   copy (W*32 bytes each).  Clean sequences have X = 0 and no bytes.
// Two launches cover every pair with at least one flagged side exactly once:
//   exact_rows_kernel    : CTA = (flagged row, block); threads = the block's queries (clean or flagged)
//   exact_queries_kernel : CTA = (flagged query, block); threads = the block's CLEAN rows (coalesced tile loads)
// Results merge into best[q] = min(distance << 32 | row) like the fast kernel: ties -> lowest row.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;

struct ExArgs {
    const uint32_t* db_hi; const uint32_t* db_lo; const uint16_t* row_len; uint32_t n_rows; uint32_t W;
    const uint32_t* q_hi; const uint32_t* q_lo; const uint16_t* q_len; uint32_t n_q;
    const uint32_t* blocks; uint32_t n_blocks; uint32_t row_index_base;
    const uint32_t* xr_ids; const uint32_t* xr_x; const uint8_t* xr_bytes; uint32_t n_xr;
    const uint32_t* xq_ids; const uint32_t* xq_x; const uint8_t* xq_bytes; uint32_t n_xq;
    unsigned long long* best;
};

__device__ __forceinline__ uint32_t prefix_mask(int bits) { return bits >= 32 ? 0xffffffffu : (bits <= 0 ? 0u : ((1u << bits) - 1u)); }

// index of `id` in the sorted list, or -1
__device__ __forceinline__ int find_id(const uint32_t* ids, uint32_t n, uint32_t id) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ids[mid] < id) lo = mid + 1; else hi = mid;
    }
    return (lo < n && ids[lo] == id) ? static_cast<int>(lo) : -1;
}

__global__ void __launch_bounds__(kThreads) exact_rows_kernel(const ExArgs a) {
    const uint32_t* blk = a.blocks + 4 * blockIdx.y;
    const uint32_t q_begin = blk[0], q_end = blk[1], r_begin = blk[2], r_end = blk[3];
    const uint32_t e = blockIdx.x;
    const uint32_t row = a.xr_ids[e];
    if (row < r_begin || row >= r_end || row >= a.n_rows) return;
    const uint32_t W = a.W;
    const int rlen = a.row_len[row] & 0x7fff;
    const uint32_t tile = row >> 5, rr = row & 31u;
    const uint32_t* rx = a.xr_x + static_cast<size_t>(e) * W;
    const uint8_t* rb = a.xr_bytes + static_cast<size_t>(e) * W * 32;
    for (uint32_t q = q_begin + threadIdx.x; q < q_end; q += kThreads) {
        const uint32_t qraw = a.q_len[q];
        const int m = min(rlen, static_cast<int>(qraw & 0x7fff));
        const int xe = (qraw & 0x8000u) ? find_id(a.xq_ids, a.n_xq, q) : -1;
        uint32_t d = 0;
        for (uint32_t w = 0; 32 * w < static_cast<uint32_t>(m); ++w) {
            const uint32_t msk = prefix_mask(m - 32 * static_cast<int>(w));
            const uint32_t rh = a.db_hi[(static_cast<size_t>(tile) * W + w) * 32 + rr], rl = a.db_lo[(static_cast<size_t>(tile) * W + w) * 32 + rr];
            const uint32_t qh = a.q_hi[static_cast<size_t>(q) * W + w], ql = a.q_lo[static_cast<size_t>(q) * W + w];
            const uint32_t xr = rx[w] & msk;
            const uint32_t xq = (xe >= 0 ? a.xq_x[static_cast<size_t>(xe) * W + w] : 0u) & msk;
            d += __popc(((rh ^ qh) | (rl ^ ql)) & msk & ~(xr | xq)) + __popc(xr ^ xq);
            uint32_t both = xr & xq;
            while (both) {
                const uint32_t i = 32 * w + (__ffs(both) - 1);
                both &= both - 1;
                d += rb[i] != a.xq_bytes[static_cast<size_t>(xe) * W * 32 + i];
            }
        }
        atomicMin(a.best + q, (static_cast<unsigned long long>(d) << 32) | static_cast<unsigned long long>(a.row_index_base + row));
    }
}

__global__ void __launch_bounds__(kThreads) exact_queries_kernel(const ExArgs a) {
    const uint32_t* blk = a.blocks + 4 * blockIdx.y;
    const uint32_t q_begin = blk[0], q_end = blk[1], r_begin = blk[2], r_end = min(blk[3], a.n_rows);
    const uint32_t e = blockIdx.x;
    const uint32_t q = a.xq_ids[e];
    if (q < q_begin || q >= q_end) return;
    const uint32_t W = a.W;
    const int qlen = a.q_len[q] & 0x7fff;
    const uint32_t* qx = a.xq_x + static_cast<size_t>(e) * W;
    unsigned long long mine = ~0ull;
    for (uint32_t row = r_begin + threadIdx.x; row < r_end; row += kThreads) {
        const uint32_t rraw = a.row_len[row];
        if (rraw & 0x8000u) continue;  // flagged rows x every query: exact_rows_kernel
        const int m = min(qlen, static_cast<int>(rraw & 0x7fff));
        const uint32_t tile = row >> 5, rr = row & 31u;
        uint32_t d = 0;
        for (uint32_t w = 0; 32 * w < static_cast<uint32_t>(m); ++w) {
            const uint32_t msk = prefix_mask(m - 32 * static_cast<int>(w));
            const uint32_t rh = a.db_hi[(static_cast<size_t>(tile) * W + w) * 32 + rr], rl = a.db_lo[(static_cast<size_t>(tile) * W + w) * 32 + rr];
            const uint32_t qh = a.q_hi[static_cast<size_t>(q) * W + w], ql = a.q_lo[static_cast<size_t>(q) * W + w];
            const uint32_t xq = qx[w] & msk;
            d += __popc(((rh ^ qh) | (rl ^ ql)) & msk & ~xq) + __popc(xq);  // clean row: every exceptional query column mismatches
        }
        const unsigned long long key = (static_cast<unsigned long long>(d) << 32) | static_cast<unsigned long long>(a.row_index_base + row);
        mine = key < mine ? key : mine;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = other < mine ? other : mine;
    }
    if ((threadIdx.x & 31u) == 0 && mine != ~0ull) atomicMin(a.best + q, mine);
}

}  // namespace

extern "C" int mmlst_hamming_exact_dev(const uint32_t* db_hi, const uint32_t* db_lo, const uint16_t* row_len, uint32_t n_rows, uint32_t W,
                                       const uint32_t* q_hi, const uint32_t* q_lo, const uint16_t* q_len, uint32_t n_q,
                                       const uint32_t* blocks, uint32_t n_blocks, uint32_t row_index_base,
                                       const uint32_t* xr_ids, const uint32_t* xr_x, const uint8_t* xr_bytes, uint32_t n_xr,
                                       const uint32_t* xq_ids, const uint32_t* xq_x, const uint8_t* xq_bytes, uint32_t n_xq,
                                       unsigned long long* best, void* stream) {
    if (n_q == 0 || n_rows == 0 || n_blocks == 0 || (n_xr == 0 && n_xq == 0)) return MMLST_OK;
    if (!db_hi || !db_lo || !row_len || !q_hi || !q_lo || !q_len || !blocks || !best || (n_xr && (!xr_ids || !xr_x || !xr_bytes)) ||
        (n_xq && (!xq_ids || !xq_x || !xq_bytes))) {
        mmlst_set_error("mmlst_hamming_exact_dev: null pointer");
        return MMLST_E_ARG;
    }
    if (n_blocks > 65535) { mmlst_set_error("mmlst_hamming_exact_dev: more than 65535 blocks per launch"); return MMLST_E_ARG; }
    ExArgs a{db_hi, db_lo, row_len, n_rows, W, q_hi, q_lo, q_len, n_q, blocks, n_blocks, row_index_base,
             xr_ids, xr_x, xr_bytes, n_xr, xq_ids, xq_x, xq_bytes, n_xq, best};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_xr) {
        exact_rows_kernel<<<dim3(n_xr, n_blocks), kThreads, 0, s>>>(a);
        if (int rc = mmlst_cuda_fail(cudaGetLastError(), "exact_rows_kernel")) return rc;
    }
    if (n_xq) {
        exact_queries_kernel<<<dim3(n_xq, n_blocks), kThreads, 0, s>>>(a);
        if (int rc = mmlst_cuda_fail(cudaGetLastError(), "exact_queries_kernel")) return rc;
    }
    return MMLST_OK;
}
