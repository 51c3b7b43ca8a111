// Stage 3: closest allele by zip-truncated Hamming distance (replaces metaMLST_functions.py:230-234 driven by
// metamlst-merge.py:174-181).  2-bit codes stored as two bit-planes => one 32-base mismatch mask costs
// XOR + LOP3 ((qh^rh)|(ql^rl)) and one POPC.  thread = DB row (planes held in registers, loaded coalesced from
// 32-row word-major tiles), CTA = 256 rows, queries staged in shared memory and broadcast with 128-bit LDS;
// per query a warp-level REDUX.MIN of (distance<<8 | row-in-CTA), then one 64-bit atomicMin per (CTA, query) of
// (distance<<32 | row) which also resolves ties to the lowest row.
#include "common.cuh"

namespace {

constexpr int kRowsPerCta = 256;
constexpr int kQChunk = 64;
constexpr int kQSplit = 2048;  // queries handled by one CTA (gridDim.z splits longer query ranges)

struct HamArgs {
    const uint32_t* db_hi; const uint32_t* db_lo; const uint16_t* row_len; uint32_t n_rows;
    const uint32_t* q_hi; const uint32_t* q_lo; const uint16_t* q_len; uint32_t n_q;
    const uint32_t* blocks; uint32_t n_blocks; uint32_t row_index_base;
    unsigned long long* best;
};

template <int W>
__global__ void __launch_bounds__(kRowsPerCta) hamming_min_kernel(const HamArgs a) {
    static_assert(W % 4 == 0, "W must be a multiple of 4 (128-bit query loads)");
    const uint32_t* blk = a.blocks + 4 * blockIdx.y;
    const uint32_t q_begin = blk[0], q_end = blk[1], r_begin = blk[2], r_end = blk[3];
    const uint32_t row0 = (r_begin & ~31u) + blockIdx.x * kRowsPerCta;
    if (row0 >= r_end) return;
    const uint32_t qs = q_begin + blockIdx.z * kQSplit;
    if (qs >= q_end) return;
    const uint32_t qe = min(q_end, qs + kQSplit);

    __shared__ uint4 sq_hi[kQChunk][W / 4];
    __shared__ uint4 sq_lo[kQChunk][W / 4];
    __shared__ uint32_t sq_len[kQChunk];
    __shared__ uint32_t part[kRowsPerCta / 32][kQChunk];
    __shared__ uint32_t s_minw;

    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t row = row0 + threadIdx.x;
    const bool valid = (row >= r_begin) && (row < r_end) && (row < a.n_rows);
    uint32_t rh[W], rl[W];
    uint32_t rlen = 0;
    {
        const uint32_t tile = row >> 5, r = row & 31u;
        const bool inb = row < ((a.n_rows + 31u) & ~31u);
#pragma unroll
        for (int w = 0; w < W; ++w) {
            rh[w] = inb ? a.db_hi[(static_cast<size_t>(tile) * W + w) * 32 + r] : 0u;
            rl[w] = inb ? a.db_lo[(static_cast<size_t>(tile) * W + w) * 32 + r] : 0u;
        }
        if (valid) rlen = a.row_len[row];
    }
    if (threadIdx.x == 0) s_minw = 0xffffffffu;
    __syncthreads();
    {
        uint32_t mw = valid ? (rlen >> 5) : 0xffffffffu;
        mw = __reduce_min_sync(0xffffffffu, mw);
        if (lane == 0) atomicMin(&s_minw, mw);
    }
    __syncthreads();
    const uint32_t cta_minw = s_minw;

    for (uint32_t qc = qs; qc < qe; qc += kQChunk) {
        const uint32_t nq = min(uint32_t(kQChunk), qe - qc);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nq * (W / 4); i += kRowsPerCta) {
            const uint32_t qi = i / (W / 4), w4 = i % (W / 4);
            sq_hi[qi][w4] = reinterpret_cast<const uint4*>(a.q_hi + static_cast<size_t>(qc + qi) * W)[w4];
            sq_lo[qi][w4] = reinterpret_cast<const uint4*>(a.q_lo + static_cast<size_t>(qc + qi) * W)[w4];
        }
        if (threadIdx.x < nq) sq_len[threadIdx.x] = a.q_len[qc + threadIdx.x];
        __syncthreads();
        for (uint32_t qi = 0; qi < nq; ++qi) {
            const uint32_t qlen = sq_len[qi];
            const uint32_t fw = min(cta_minw, qlen >> 5);     // words that are full for every pair of this CTA
            const uint32_t minlen = min(qlen, rlen);
            uint32_t acc = 0;
#pragma unroll
            for (int w4 = 0; w4 < W / 4; ++w4) {
                const uint4 qh = sq_hi[qi][w4];
                const uint4 ql = sq_lo[qi][w4];
                const uint32_t qhv[4] = {qh.x, qh.y, qh.z, qh.w};
                const uint32_t qlv[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int w = 4 * w4 + k;
                    uint32_t m = (rh[w] ^ qhv[k]) | (rl[w] ^ qlv[k]);
                    if (uint32_t(w) >= fw) {  // CTA-uniform branch: tail words are masked to min(len_q, len_r) (H9)
                        const int vb = int(minlen) - 32 * w;
                        const uint32_t msk = vb >= 32 ? 0xffffffffu : (vb <= 0 ? 0u : ((1u << vb) - 1u));
                        m &= msk;
                    }
                    acc += __popc(m);
                }
            }
            uint32_t key = valid ? ((acc << 8) | threadIdx.x) : 0xffffffffu;
            key = __reduce_min_sync(0xffffffffu, key);
            if (lane == 0) part[wib][qi] = key;
        }
        __syncthreads();
        if (threadIdx.x < nq) {
            uint32_t key = 0xffffffffu;
#pragma unroll
            for (int w = 0; w < kRowsPerCta / 32; ++w) key = min(key, part[w][threadIdx.x]);
            if (key != 0xffffffffu) {
                const unsigned long long v = (static_cast<unsigned long long>(key >> 8) << 32) |
                                             static_cast<unsigned long long>(a.row_index_base + row0 + (key & 255u));
                atomicMin(a.best + qc + threadIdx.x, v);
            }
        }
    }
}

template <int W>
int launch(const HamArgs& a, uint32_t max_rows, uint32_t max_q, cudaStream_t s) {
    dim3 grid((max_rows + 31 + kRowsPerCta - 1) / kRowsPerCta + 1, a.n_blocks, (max_q + kQSplit - 1) / kQSplit);
    hamming_min_kernel<W><<<grid, kRowsPerCta, 0, s>>>(a);
    return mmlst_cuda_fail(cudaGetLastError(), "hamming_min_kernel");
}

}  // namespace

// max_block_rows / max_block_queries size the grid (blocks with fewer rows exit early)
extern "C" int mmlst_hamming_min_dev2(const uint32_t* db_hi, const uint32_t* db_lo, const uint16_t* row_len, uint32_t n_rows,
                                      uint32_t W, const uint32_t* q_hi, const uint32_t* q_lo, const uint16_t* q_len,
                                      uint32_t n_q, const uint32_t* blocks, uint32_t n_blocks, uint32_t max_block_rows,
                                      uint32_t max_block_queries, uint32_t row_index_base, unsigned long long* best,
                                      void* stream) {
    if (n_q == 0 || n_rows == 0 || n_blocks == 0) return MMLST_OK;
    if (!db_hi || !db_lo || !row_len || !q_hi || !q_lo || !q_len || !blocks || !best) {
        mmlst_set_error("mmlst_hamming_min_dev: null pointer");
        return MMLST_E_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(q_hi) & 15) || (reinterpret_cast<uintptr_t>(q_lo) & 15)) {
        mmlst_set_error("mmlst_hamming_min_dev: query planes must be 16-byte aligned");
        return MMLST_E_ARG;
    }
    if (n_blocks > 65535) { mmlst_set_error("mmlst_hamming_min_dev: more than 65535 blocks per launch"); return MMLST_E_ARG; }
    HamArgs a{db_hi, db_lo, row_len, n_rows, q_hi, q_lo, q_len, n_q, blocks, n_blocks, row_index_base, best};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (W) {
        case 8: return launch<8>(a, max_block_rows, max_block_queries, s);
        case 16: return launch<16>(a, max_block_rows, max_block_queries, s);
        case 24: return launch<24>(a, max_block_rows, max_block_queries, s);
        case 32: return launch<32>(a, max_block_rows, max_block_queries, s);
        default:
            mmlst_set_error("mmlst_hamming_min_dev: W=%u unsupported (8, 16, 24 or 32 words per plane; rows up to 1024 bases)", W);
            return MMLST_E_ARG;
    }
}

extern "C" int mmlst_hamming_min_dev(const uint32_t* db_hi, const uint32_t* db_lo, const uint16_t* row_len, uint32_t n_rows,
                                     uint32_t W, const uint32_t* q_hi, const uint32_t* q_lo, const uint16_t* q_len,
                                     uint32_t n_q, const uint32_t* blocks, uint32_t n_blocks, uint32_t row_index_base,
                                     unsigned long long* best, void* stream) {
    // conservative grid: every block may span all rows / all queries
    return mmlst_hamming_min_dev2(db_hi, db_lo, row_len, n_rows, W, q_hi, q_lo, q_len, n_q, blocks, n_blocks, n_rows, n_q,
                                  row_index_base, best, stream);
}
