// Stage 3 on the 5th-generation tensor cores: the ALL-PAIRS closest-allele sweep (BASELINE.json configs[4]: 10 k reconstructed loci
// against 1 M DB alleles) as a tcgen05 GEMM.  stringDiff (metaMLST_functions.py:230-234) over 2-bit bases is an inner product:
// with every base c = (b1, b0) written as three signs  f = (s1, s0, s1*s0),  s = 1 - 2b,  and nothing (0, 0, 0) beyond the end
// of a sequence, a pair of aligned bases contributes  s1 s1' + s0 s0' + s1 s0 s1' s0' = 3 if equal, -1 if not  (each kind of
// mismatch flips exactly two of the three products), so over the m = min(len_q, len_r) columns zip() really compares
//     <f(q), f(r)> = 4 * matches - m        =>        distance = m - matches = (3 m - <f(q), f(r)>) / 4 ,
// the zero padding giving the zip truncation (H9) for free.  The three planes are 8-bit floats e4m3 (+1 = 0x38, -1 = 0xB8, 0),
// products and sums of at most 3 * 1024 terms are exact in the FP32 accumulator: the result is the integer the POPC kernel finds.
//
//   operands   queries (A, M) and DB rows (B, N) expanded ONCE into e4m3 tile images in HBM, laid out exactly as the shared-memory
//              image the MMA reads: per (tile, K-block of 128 features) TM rows x 128 bytes, K-major, 128-byte swizzle (16-byte chunk
//              index XOR row mod 8), so a stage is filled by two 1-D TMA bulk copies (cp.async.bulk, no tensor map).  K-block order
//              = position block major, plane minor, which lets a tile stop at the last position block any of its pairs can reach.
//   kernel     persistent, one CTA per SM, 192 threads: warp 0 = TMA producer (4-stage ring of A 16 KB + B 32 KB, mbarrier
//              full/empty), warp 1 = MMA issuer (one lane: tcgen05.mma cta_group::1 kind::f8f6f4, M=128 N=256 K=32, 4 per K-block,
//              accumulators in TMEM, two 256-column buffers; tcgen05.commit releases the smem stage / publishes the accumulator),
//              warps 2-5 = epilogue (tcgen05.ld 32x32b.x32 of the warp's 32 TMEM lanes = 32 queries x 32 rows at a time,
//              distance = (3 min(len) - dot) >> 2, running (distance << 32 | row) minimum per query, ONE 64-bit atomicMin per
//              (query, tile) -- ties resolve to the lowest row like the POPC kernel).
//   schedule   tile t -> (N tile, M tile) with the M tile fastest: CTAs running together share one DB tile (L2 hits), the expanded
//              queries (23 MB at 10 k x 768) stay L2-resident.
// Flagged sequences (non-ACGT letters, bit 15 of the length) are skipped here and handled by hamming_exact.cu, as with the POPC kernel.
#include "common.cuh"

namespace {

constexpr int TC_M = 128, TC_N = 256, TC_KB = 128;       // tile shape, K-block = 128 features = 128 bytes per row
constexpr int TC_STAGES = 4;
constexpr uint32_t A_STAGE = TC_M * TC_KB, B_STAGE = TC_N * TC_KB;   // 16 KB, 32 KB
constexpr uint32_t STAGE_BYTES = A_STAGE + B_STAGE;
constexpr int TC_THREADS = 192;

// ---- operand expansion: bit planes -> e4m3 tile images --------------------------------------------------------------------------
// one thread = 16 consecutive positions of one plane of one row = one 16-byte chunk of the image
struct ExpandArgs {
    const uint32_t* hi; const uint32_t* lo; const uint16_t* len; uint32_t n; uint32_t W; int tiled;   // tiled: DB layout [(tile*W + w)*32 + r]
    uint32_t TM; uint32_t n_tiles; uint8_t* img;
};

__device__ __forceinline__ uint32_t spread4(uint32_t bits4) { return (bits4 * 0x00204081u) & 0x01010101u; }  // bit j -> bit 0 of byte j

__global__ void __launch_bounds__(256) hamming_tc_expand_kernel(const ExpandArgs a) {
    const uint32_t PB = a.W / 4;                       // position blocks of 128 bases
    const uint32_t KB = 3 * PB;
    const uint64_t chunks_per_tile = static_cast<uint64_t>(KB) * a.TM * 8;
    const uint64_t g = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (g >= chunks_per_tile * a.n_tiles) return;
    const uint32_t tile = static_cast<uint32_t>(g / chunks_per_tile);
    uint32_t rem = static_cast<uint32_t>(g % chunks_per_tile);
    const uint32_t kb = rem / (a.TM * 8); rem %= a.TM * 8;
    const uint32_t r = rem >> 3, chunk = rem & 7u;
    const uint32_t pb = kb / 3, plane = kb % 3;
    const uint32_t row = tile * a.TM + r;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (row < a.n) {
        const uint32_t ln = a.len[row];
        if (!(ln & 0x8000u)) {
            const uint32_t p0 = pb * 128 + chunk * 16;            // first base of this chunk
            if (p0 < ln) {
                const uint32_t w = p0 >> 5, sh = p0 & 31u;
                const size_t idx = a.tiled ? (static_cast<size_t>(row >> 5) * a.W + w) * 32 + (row & 31u) : static_cast<size_t>(row) * a.W + w;
                const uint32_t h = (a.hi[idx] >> sh) & 0xffffu, l = (a.lo[idx] >> sh) & 0xffffu;
                const uint32_t neg = plane == 0 ? h : plane == 1 ? l : (h ^ l);   // s1 s0 is negative iff exactly one of b1, b0 is set
                const uint32_t nvalid = min(16u, ln - p0);
                uint32_t wds[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t bits = (neg >> (4 * j)) & 0xfu;
                    uint32_t v = 0x38383838u | (spread4(bits) << 7);
                    const int left = static_cast<int>(nvalid) - 4 * j;
                    if (left <= 0) v = 0; else if (left < 4) v &= (1u << (8 * left)) - 1u;
                    wds[j] = v;
                }
                out = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            }
        }
    }
    // swizzled image: (tile, kb) block of TM x 128 B; row r at r * 128 (8-row groups are 1024 B apart by construction), chunk ^ (r & 7)
    uint8_t* dst = a.img + (static_cast<size_t>(tile) * KB + kb) * (static_cast<size_t>(a.TM) * 128) + static_cast<size_t>(r) * 128 + ((chunk ^ (r & 7u)) << 4);
    *reinterpret_cast<uint4*>(dst) = out;
}

// per tile: the largest clean length (the K loop of a tile pair stops at the smaller of the two)
__global__ void __launch_bounds__(256) hamming_tc_maxlen_kernel(const uint16_t* len, uint32_t n, uint32_t TM, uint32_t n_tiles, uint32_t* maxlen) {
    const uint32_t tile = blockIdx.x;
    uint32_t m = 0;
    for (uint32_t r = threadIdx.x; r < TM; r += blockDim.x) {
        const uint32_t row = tile * TM + r;
        if (row < n) { const uint32_t l = len[row]; if (!(l & 0x8000u)) m = max(m, l); }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    __shared__ uint32_t sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(&sh, m);
    __syncthreads();
    if (threadIdx.x == 0 && tile < n_tiles) maxlen[tile] = sh;
}

// ---- tcgen05 wrappers -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on the mbarrier when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO), version 1 (Blackwell); start address in 16-byte units
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = E4M3 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(TC_N >> 3) << 17) | (static_cast<uint32_t>(TC_M >> 4) << 24);

struct TcArgs {
    const uint8_t* a_img; const uint8_t* b_img; const uint16_t* q_len; const uint16_t* row_len;
    const uint32_t* a_maxlen; const uint32_t* b_maxlen;
    uint32_t n_q, n_rows, KB, n_mt, n_nt, row_index_base;
    unsigned long long* best;
};

__global__ void __launch_bounds__(TC_THREADS, 1) hamming_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tfull = empty + TC_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint16_t* s_rowlen = reinterpret_cast<uint16_t*>(tmem_slot + 4);   // [2][TC_N]
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t n_tiles = a.n_mt * a.n_nt;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const uint32_t nt = t / a.n_mt, mt = t % a.n_mt;
                const uint32_t reach = min(a.a_maxlen[mt], a.b_maxlen[nt]);
                const uint32_t kbn = min(a.KB, 3u * ((reach + 127u) >> 7));
                for (uint32_t kb = 0; kb < kbn; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(full + stage, STAGE_BYTES);
                    bulk_g2s(sa, a.a_img + (static_cast<size_t>(mt) * a.KB + kb) * A_STAGE, A_STAGE, full + stage);
                    bulk_g2s(sa + A_STAGE, a.b_img + (static_cast<size_t>(nt) * a.KB + kb) * B_STAGE, B_STAGE, full + stage);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const uint32_t nt = t / a.n_mt, mt = t % a.n_mt;
                const uint32_t reach = min(a.a_maxlen[mt], a.b_maxlen[nt]);
                const uint32_t kbn = min(a.KB, 3u * ((reach + 127u) >> 7));
                mbar_wait(tempty + acc, acc_phase ^ 1u);          // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + acc * TC_N;
                for (uint32_t kb = 0; kb < kbn; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t ad = smem_desc_sw128(sa), bd = smem_desc_sw128(sa + A_STAGE);
#pragma unroll
                    for (uint32_t k = 0; k < TC_KB / 32; ++k) umma_f8(d, ad + 2 * k, bd + 2 * k, kIdesc, (kb | k) ? 1u : 0u);   // +32 B along K
                    umma_commit(empty + stage);                   // stage free once these MMAs have read it
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull + acc);                         // accumulator complete
                if (kbn == 0) { /* nothing to compare: the epilogue still runs on zeros */ }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const uint32_t quarter = warp & 3u;
        const uint32_t et = threadIdx.x - 64;                    // 0..127
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const uint32_t nt = t / a.n_mt, mt = t % a.n_mt;
            const uint32_t reach = min(a.a_maxlen[mt], a.b_maxlen[nt]);
            const uint32_t kbn = min(a.KB, 3u * ((reach + 127u) >> 7));
            uint16_t* rl = s_rowlen + acc * TC_N;
            for (uint32_t j = et; j < TC_N; j += 128) {
                const uint32_t row = nt * TC_N + j;
                rl[j] = row < a.n_rows ? a.row_len[row] : static_cast<uint16_t>(0xffffu);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps only
            const uint32_t q = mt * TC_M + quarter * 32 + lane;
            const uint32_t lq_raw = q < a.n_q ? a.q_len[q] : 0xffffu;
            const bool q_ok = !(lq_raw & 0x8000u);
            const uint32_t lq = lq_raw & 0x7fffu;
            mbar_wait(tfull + acc, acc_phase);
            tc_fence_after();
            unsigned long long best = ~0ull;
            const uint32_t taddr = tmem_base + acc * TC_N + ((quarter * 32u) << 16);
#pragma unroll 1
            for (uint32_t c = 0; c < TC_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t lr = rl[c * 32 + j];
                    if (lr & 0x8000u) continue;                   // flagged row / beyond the DB: the exact path's business
                    const int m = static_cast<int>(min(lq, lr));
                    const int dot = kbn ? __float2int_rn(__uint_as_float(v[j])) : 0;
                    const uint32_t dist = static_cast<uint32_t>(3 * m - dot) >> 2;
                    const unsigned long long key = (static_cast<unsigned long long>(dist) << 32) | (a.row_index_base + nt * TC_N + c * 32 + j);
                    best = key < best ? key : best;
                }
            }
            tc_fence_before();
            mbar_arrive(tempty + acc);                            // 128 arrivals free the accumulator
            if (q_ok && best != ~0ull) atomicMin(a.best + q, best);
            asm volatile("bar.sync 1, 128;" ::: "memory");       // rl[] of this buffer is rewritten two tiles from now; keep the warps together
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

constexpr size_t kTcSmem = 1024 + TC_STAGES * STAGE_BYTES + 16 * 8 + 16 + 2 * TC_N * 2 + 64;

}  // namespace

// bytes of the e4m3 tile image of n sequences of W plane words (tile_rows = 128 for queries, 256 for DB rows); 0 if W is not a multiple of 4
extern "C" size_t mmlst_hamming_tc_image_bytes(uint32_t n, uint32_t W, uint32_t tile_rows) {
    if (W == 0 || (W & 3u) || (tile_rows != TC_M && tile_rows != TC_N)) return 0;
    const size_t tiles = (static_cast<size_t>(n) + tile_rows - 1) / tile_rows;
    return tiles * (3 * (W / 4)) * static_cast<size_t>(tile_rows) * 128;
}

extern "C" int mmlst_hamming_tc_expand_dev(const uint32_t* hi, const uint32_t* lo, const uint16_t* len, uint32_t n, uint32_t W, int db_tiled_layout,
                                           uint32_t tile_rows, uint8_t* image, uint32_t* tile_maxlen, void* stream) {
    if (!hi || !lo || !len || !image || !tile_maxlen) { mmlst_set_error("mmlst_hamming_tc_expand_dev: null pointer"); return MMLST_E_ARG; }
    if (mmlst_hamming_tc_image_bytes(n, W, tile_rows) == 0 && n) { mmlst_set_error("mmlst_hamming_tc_expand_dev: W must be a multiple of 4 (128 bases), tile_rows 128 or 256"); return MMLST_E_ARG; }
    if (n == 0) return MMLST_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t n_tiles = (n + tile_rows - 1) / tile_rows;
    const ExpandArgs a{hi, lo, len, n, W, db_tiled_layout, tile_rows, n_tiles, image};
    const uint64_t chunks = static_cast<uint64_t>(n_tiles) * (3 * (W / 4)) * tile_rows * 8;
    if ((chunks + 255) / 256 > 0x7fffffffull) { mmlst_set_error("mmlst_hamming_tc_expand_dev: too many rows for one launch"); return MMLST_E_RANGE; }
    hamming_tc_expand_kernel<<<static_cast<unsigned>((chunks + 255) / 256), 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    hamming_tc_maxlen_kernel<<<n_tiles, 256, 0, st>>>(len, n, tile_rows, n_tiles, tile_maxlen);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}

extern "C" int mmlst_hamming_tc_search_dev(const uint8_t* q_image, const uint32_t* q_tile_maxlen, const uint16_t* q_len, uint32_t n_q,
                                           const uint8_t* db_image, const uint32_t* db_tile_maxlen, const uint16_t* row_len, uint32_t n_rows,
                                           uint32_t W, uint32_t row_index_base, unsigned long long* best, void* stream) {
    if (n_q == 0 || n_rows == 0) return MMLST_OK;
    if (!q_image || !q_tile_maxlen || !q_len || !db_image || !db_tile_maxlen || !row_len || !best) { mmlst_set_error("mmlst_hamming_tc_search_dev: null pointer"); return MMLST_E_ARG; }
    if (W == 0 || (W & 3u) || W > 32) { mmlst_set_error("mmlst_hamming_tc_search_dev: W must be a multiple of 4, at most 32"); return MMLST_E_ARG; }
    if ((reinterpret_cast<uintptr_t>(q_image) | reinterpret_cast<uintptr_t>(db_image)) & 15) { mmlst_set_error("mmlst_hamming_tc_search_dev: images must be 16-byte aligned"); return MMLST_E_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool configured_by_device[MMLST_MAX_DEVICES] = {false};
    bool& configured = configured_by_device[mmlst_current_device()];
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTcSmem)));
        configured = true;
    }
    TcArgs a;
    a.a_img = q_image; a.b_img = db_image; a.q_len = q_len; a.row_len = row_len; a.a_maxlen = q_tile_maxlen; a.b_maxlen = db_tile_maxlen;
    a.n_q = n_q; a.n_rows = n_rows; a.KB = 3 * (W / 4); a.n_mt = (n_q + TC_M - 1) / TC_M; a.n_nt = (n_rows + TC_N - 1) / TC_N;
    a.row_index_base = row_index_base; a.best = best;
    const uint64_t tiles = static_cast<uint64_t>(a.n_mt) * a.n_nt;
    if (tiles > 0xffffffffull) { mmlst_set_error("mmlst_hamming_tc_search_dev: too many tiles"); return MMLST_E_RANGE; }
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(tiles, static_cast<uint64_t>(mmlst_num_sms())));
    hamming_tc_kernel<<<grid, TC_THREADS, kTcSmem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}
