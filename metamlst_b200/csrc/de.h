// Blackwell's hardware decompression engine (cuMemBatchDecompressAsync, CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE) behind two internal calls shared by
// the BAM ingest (csrc/ingest.cu: BGZF blocks) and the host-buffer score path (csrc/api.cu: DEFLATE-compressed score stream).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

// MMLST_OK when `device` has hardware DEFLATE and the driver exposes the batch call; MMLST_E_CUDA with mmlst_last_error() set otherwise.
int mmlst_de_available(int device);

struct MmlstPlainCopy { void* dst; const void* src; size_t bytes; };   // host -> device, queued on the copy lane after the compressed slices

// Copy h_comp[0, n_bytes) to d_comp in slices on a per-device copy stream and queue the decompression of every slice's blocks on `st` as the
// slice lands (copy engine and decompression engine overlap).  prm[b] (src / dst device pointers already set) must be ordered by source
// offset; src_off[b] = byte offset of block b's payload inside the buffer.  Returns after the LAST copy has finished (the host buffer is
// free again); the decompression itself is stream-ordered on `st`.  `plain`: further host -> device copies that follow the compressed slices on the
// copy lane (the part of a stream shipped uncompressed: PCIe carries it while the engine is still inflating); `st` waits for them too.
// `slices`: every slice costs the engine a ramp, every slice fewer delays its start: 8 for a BAM file (hundreds of MB), 3 for the score stream of
// one sample (tens of MB; profiles/r2z_e2e_sweep.json).
int mmlst_h2d_inflate(int device, cudaStream_t st, uint8_t* d_comp, const uint8_t* h_comp, size_t n_bytes,
                      std::vector<CUmemDecompressParams>& prm, const std::vector<uint64_t>& src_off,
                      const std::vector<MmlstPlainCopy>* plain = nullptr, int slices = 8);

// The same for compressed bytes that are NOT one contiguous host range (the chosen contigs of a compressed pileup stream): segment i is
// h_src[i] .. + bytes[i] -> d_dst[i], and owns the blocks prm[first_block[i] .. first_block[i+1]) (src / dst device pointers already set).  The
// segments are dealt over at most `groups` copy/decompress rounds; returns after the last copy has finished.
struct MmlstSegment { uint8_t* d_dst; const uint8_t* h_src; size_t bytes; };
int mmlst_h2d_inflate_segments(int device, cudaStream_t st, const std::vector<MmlstSegment>& segs, std::vector<CUmemDecompressParams>& prm,
                               const std::vector<uint32_t>& first_block, int groups = 6);
