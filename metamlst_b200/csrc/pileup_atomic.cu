// Stage 2, implementation 1: per-base atomics (baseline; kept as the cross-check for the bit-sliced kernel).
// One warp per record, lane = reference offset inside a 32-column word, one RED.ADD per counted base.
// Bound by atomic throughput (~1 increment/lane/1.3 clk per SM), far below the HBM roof -- see DESIGN.md.
#include "common.cuh"
#include "pileup.cuh"

namespace {

__global__ void __launch_bounds__(256) pileup_atomic_kernel(const PileupArgs a) {
    const uint32_t n_chunks = pileup_n_chunks(a);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wib = threadIdx.x >> 5;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    const mmlst_chunk ck = a.chunks[ci];
    for (uint32_t rec = ck.rec_begin + wib; rec < ck.rec_end; rec += wpb) {
        const mmlst_prec pr = a.recs[rec];
        const int p = pr.pos;
        const uint32_t off = pr.row_off + ck.plane_delta;
        const uint32_t rl = pr.reflen;
        const bool pass = (int(pr.as_named) >= a.minscore) && (int(pr.xm_named) <= a.max_xm);
        const uint32_t nw = pr.nw;  // contig words touched; rows are aligned to the contig's 32-column words
        const long long col0 = static_cast<long long>(p >> 5) * 32;
        (void)rl;
        for (uint32_t j = 0; j < nw; ++j) {
            const uint32_t v = a.planes[off + 3 * j + 0];
            const uint32_t b1 = a.planes[off + 3 * j + 1];
            const uint32_t b0 = a.planes[off + 3 * j + 2];
            const uint32_t vb = (v >> lane) & 1u, h = (b1 >> lane) & 1u, l = (b0 >> lane) & 1u;
            if (!(vb | l)) continue;  // not in the column
            const long long col = col0 + 32 * j + lane;
            if (col < 0 || col >= static_cast<long long>(ck.contig_len)) continue;
            const uint32_t bin = (vb && pass) ? (h * 2u + l) : 4u;
            atomicAdd(a.counts + (static_cast<size_t>(ck.col_base) + col) * 5 + bin, 1u);
        }
    }
    }
}

}  // namespace

int launch_pileup_atomic(const PileupArgs& a, cudaStream_t stream) {
    const uint32_t grid = a.n_chunks_dev ? uint32_t(mmlst_num_sms() * 8) : min(a.n_chunks, uint32_t(mmlst_num_sms() * 8));
    pileup_atomic_kernel<<<grid, 256, 0, stream>>>(a);
    return mmlst_cuda_fail(cudaGetLastError(), "pileup_atomic_kernel");
}
