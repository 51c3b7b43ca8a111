// Blackwell's hardware decompression engine (cuMemBatchDecompressAsync, CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE) behind two internal calls shared by
// the BAM ingest (csrc/ingest.cu: BGZF blocks) and the host-buffer score path (csrc/api.cu: DEFLATE-compressed score stream).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

// MMLST_OK when `device` has hardware DEFLATE and the driver exposes the batch call; MMLST_E_CUDA with mmlst_last_error() set otherwise.
int mmlst_de_available(int device);

// Copy h_comp[0, n_bytes) to d_comp in slices on a per-device copy stream and queue the decompression of every slice's blocks on `st` as the
// slice lands (copy engine and decompression engine overlap).  prm[b] (src / dst device pointers already set) must be ordered by source
// offset; src_off[b] = byte offset of block b's payload inside the buffer.  Returns after the LAST copy has finished (the host buffer is
// free again); the decompression itself is stream-ordered on `st`.
int mmlst_h2d_inflate(int device, cudaStream_t st, uint8_t* d_comp, const uint8_t* h_comp, size_t n_bytes,
                      std::vector<CUmemDecompressParams>& prm, const std::vector<uint64_t>& src_off);
