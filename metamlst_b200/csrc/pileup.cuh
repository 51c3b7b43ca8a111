// Shared declarations of the two pileup implementations.
#pragma once
#include "common.cuh"

struct PileupArgs {
    const int32_t* pos; const uint32_t* row_off; const uint16_t* reflen; const int16_t* as_named; const uint8_t* xm_named;
    const uint32_t* planes; const mmlst_chunk* chunks; uint32_t n_chunks; uint32_t max_row_words;
    int minscore, max_xm; uint32_t* counts; uint32_t total_cols;
    const uint32_t* n_chunks_dev;  // when non-null the chunk count is read on the device (written by mmlst_select_dev)
};

int launch_pileup_atomic(const PileupArgs& a, cudaStream_t stream);
int launch_pileup_bitsliced(const PileupArgs& a, cudaStream_t stream);
__device__ __forceinline__ uint32_t pileup_n_chunks(const PileupArgs& a) { return a.n_chunks_dev ? *a.n_chunks_dev : a.n_chunks; }
