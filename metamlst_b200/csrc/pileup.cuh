// Shared declarations of the two pileup implementations.
#pragma once
#include "common.cuh"

struct PileupArgs {
    const mmlst_prec* recs;
    const uint32_t* planes; const mmlst_chunk* chunks; uint32_t n_chunks; uint32_t max_row_words;
    int minscore, max_xm; uint32_t* counts; uint32_t total_cols;
    const uint32_t* n_chunks_dev;  // when non-null the chunk count is read on the device (written by mmlst_select_dev)
};

int launch_pileup_atomic(const PileupArgs& a, cudaStream_t stream);
int launch_pileup_bitsliced(const PileupArgs& a, cudaStream_t stream);
__device__ __forceinline__ uint32_t pileup_n_chunks(const PileupArgs& a) { return a.n_chunks_dev ? *a.n_chunks_dev : a.n_chunks; }

// words of a plane row for a record touching nw contig words: 3 planes, padded to an odd count (0 stays 0)
__host__ __device__ __forceinline__ uint32_t mmlst_row_words(uint32_t nw) {
    const uint32_t rw = 3u * nw;
    return rw + ((rw != 0u && (rw & 1u) == 0u) ? 1u : 0u);
}
__host__ __device__ __forceinline__ uint32_t mmlst_touched_words(uint32_t pos, uint32_t reflen) {
    return reflen ? (((pos & 31u) + reflen + 31u) >> 5) : 0u;
}
