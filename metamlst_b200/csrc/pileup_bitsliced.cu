// Stage 2, implementation 2: pileup histogram WITHOUT per-base atomics -- bit-sliced vertical counters.
//
// Why: at config-2 depth (~1.5e5 x) the histogram is 1.5 G increments into ~1e4 x 5 bins.  Shared/global atomics
// retire ~0.5-1 increment/clk/SM (B300_MICROARCH "Atomics": ATOMS 2 cyc/lane, REDG 1.29 cyc/lane), i.e. ~1e11/s
// chip-wide, two orders of magnitude under what the 6.5 TB/s HBM stream delivers (~1e13 base-increments/s).
// So the bases are never expanded: a record stays three 32-column bit-planes per word (V, B1, B0) and 32 columns
// are counted per LOP3 with carry-save adders (Harley-Seal), the same trick as a vertical popcount.
//
// Mapping
//   CTA      : persistent over chunks (<= 63 tiles of one contig); tile = 512 consecutive coordinate-sorted records.
//              Their plane rows are ONE contiguous global range and their 16-byte records another, so a tile is two
//              cp.async.bulk (TMA 1-D, SASS UBLKCP) into a double-buffered shared-memory stage, completion on one
//              mbarrier; thread 0 issues tile t+1 before tile t is counted.  No per-thread global loads at all.
//   warp     : owns one 32-column word w of the contig (w mod 8 == warp id inside the sliding 8-word window).
//   lane     : takes records lane, lane+32, ... of the tile (16 per tile), funnel-shifts the record's planes to the
//              word's alignment and adds five 1-bit planes (counted, ok, ok&B0, ok&B1, ok&B0&B1) into private
//              bit-sliced counters: ones/twos/fours/eights by a 16-input Harley-Seal block (30 LOP3 per 16 inputs),
//              the sixteens plane rippled into 6 upper planes once per tile.
//   flush    : when the warp's word changes or the chunk ends: bit-sliced add across the 32 lanes (shuffle butterfly),
//              lane i extracts column i, converts to A/C/G/T/N and issues 5 coalesced RED.ADD.
#include "common.cuh"
#include "pileup.cuh"

namespace {

constexpr int TR = 512;         // records per tile
constexpr int NWARP = 8;        // column words in the window
constexpr int NTHREADS = NWARP * 32;
constexpr int LOWP = 4;         // ones, twos, fours, eights
constexpr int UPP = 6;          // 16s .. 512s  => < 1024 records per lane between flushes (63 tiles x 16)
constexpr int NP = LOWP + UPP;  // planes per lane counter
constexpr int NPR = NP + 5;     // planes after the 32-lane reduction
constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(addr), "r"(phase)
                     : "memory");
    } while (!done);
}

// full adder on 32 columns at once: 2 LOP3
__device__ __forceinline__ void csa(uint32_t& hi, uint32_t& lo, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    hi = (a & b) | (u & c);
    lo = u ^ c;
}

struct Counter {
    uint32_t p[NP];  // p[0..3] = ones, twos, fours, eights; p[4..9] = 16s..512s
};

__device__ __forceinline__ void ripple16(Counter& c, uint32_t x) {  // add the "sixteens" plane
#pragma unroll
    for (int i = LOWP; i < NP; ++i) {
        const uint32_t carry = c.p[i] & x;
        c.p[i] ^= x;
        x = carry;
    }
}

struct Contrib { uint32_t x[5]; };  // counted, ok, ok&B0, ok&B1, ok&B0&B1

struct TileCtx {
    const uint32_t* planes;  // stage planes
    uint32_t soff_base;      // plane_delta - a0 (mod 2^32): row_off + soff_base = word offset inside the stage
    uint32_t cnt;            // records in the tile
    int minscore, max_xm;
};

// planes of tile record i (16-byte mmlst_prec in shared memory) shifted to the alignment of the column word at w32
__device__ __forceinline__ Contrib load_contrib(const TileCtx& tc, const uint4* __restrict__ meta, uint32_t i, int w32) {
    const uint4 m = meta[i];
    const int pos = static_cast<int>(m.x);
    const uint32_t nw = ((m.z & 0xffffu) + 31u) >> 5;
    const int as = static_cast<int>(m.z) >> 16;
    const int xm = static_cast<int>(m.w & 0xffu);
    const uint32_t pm = (as >= tc.minscore && xm <= tc.max_xm) ? FULL : 0u;
    const int d = w32 - pos;
    const int j0 = d >> 5;
    const uint32_t s = static_cast<uint32_t>(d) & 31u;
    const bool valid = i < tc.cnt;
    const bool lo_ok = valid && (static_cast<uint32_t>(j0) < nw);
    const bool hi_ok = valid && (static_cast<uint32_t>(j0 + 1) < nw);
    const uint32_t* r = tc.planes + (m.y + tc.soff_base) + 3 * j0;
    uint32_t vl = 0, hl = 0, ll = 0, vh = 0, hh = 0, lh = 0;
    if (lo_ok) { vl = r[0]; hl = r[1]; ll = r[2]; }
    if (hi_ok) { vh = r[3]; hh = r[4]; lh = r[5]; }
    const uint32_t v = __funnelshift_r(vl, vh, s);
    const uint32_t b1 = __funnelshift_r(hl, hh, s);
    const uint32_t b0 = __funnelshift_r(ll, lh, s);
    Contrib c;
    c.x[0] = v | b0;
    c.x[1] = v & pm;
    c.x[2] = c.x[1] & b0;
    c.x[3] = c.x[1] & b1;
    c.x[4] = c.x[2] & b1;
    return c;
}

// bit-sliced sum over the 32 lanes: every lane ends with the warp total in NPR planes
__device__ __forceinline__ void reduce_lanes(const Counter& c, uint32_t (&t)[NPR]) {
#pragma unroll
    for (int i = 0; i < NPR; ++i) t[i] = (i < NP) ? c.p[i] : 0u;
#pragma unroll
    for (int step = 0; step < 5; ++step) {
        uint32_t carry = 0;
#pragma unroll
        for (int i = 0; i < NPR; ++i) {
            if (i <= NP + step) {  // planes that can be non-zero after this step
                const uint32_t o = __shfl_xor_sync(FULL, t[i], 1 << step);
                uint32_t hi, lo;
                csa(hi, lo, t[i], o, carry);
                t[i] = lo;
                carry = hi;
            }
        }
    }
}

__device__ __forceinline__ uint32_t extract(const uint32_t (&t)[NPR], uint32_t lane) {
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < NPR; ++i) v |= ((t[i] >> lane) & 1u) << i;
    return v;
}

struct WarpAcc {
    Counter c[5];  // counted, ok, ok&B0, ok&B1, ok&B0&B1
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int i = 0; i < NP; ++i) c[k].p[i] = 0u;
    }
};

__device__ __forceinline__ void flush_word(const WarpAcc& acc, int w, const mmlst_chunk& ck, uint32_t* __restrict__ counts) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t t[NPR];
    uint32_t n[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        reduce_lanes(acc.c[k], t);
        n[k] = extract(t, lane);
    }
    const uint32_t n_cnt = n[0], n_ok = n[1], n_p0 = n[2], n_p1 = n[3], n_p01 = n[4];
    const long long col = static_cast<long long>(w) * 32 + lane;
    if (col < 0 || col >= static_cast<long long>(ck.contig_len)) return;
    uint32_t* c = counts + (static_cast<size_t>(ck.col_base) + col) * 5;
    const uint32_t nT = n_p01, nG = n_p1 - n_p01, nC = n_p0 - n_p01, nA = n_ok - n_p0 - n_p1 + n_p01, nN = n_cnt - n_ok;
    if (nA) atomicAdd(c + 0, nA);
    if (nC) atomicAdd(c + 1, nC);
    if (nG) atomicAdd(c + 2, nG);
    if (nT) atomicAdd(c + 3, nT);
    if (nN) atomicAdd(c + 4, nN);
}

template <int NS>
__global__ void __launch_bounds__(NTHREADS, 2) pileup_bitsliced_kernel(const PileupArgs a, const uint32_t stage_words) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // stage s: planes [stage_words] u32, then TR x 16-byte records
    const size_t stage_bytes = static_cast<size_t>(stage_words) * 4 + TR * sizeof(mmlst_prec);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + NS * stage_bytes);

    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase_bits = 0;  // parity per stage
    const int maxspan = int(a.max_row_words / 3u) * 32;  // upper bound of any record's reference span

    const uint32_t n_chunks = pileup_n_chunks(a);
    for (uint32_t ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const mmlst_chunk ck = a.chunks[ci];
        const uint32_t nrec = ck.rec_end - ck.rec_begin;
        const uint32_t ntiles = (nrec + TR - 1) / TR;
        uint4 g_first = make_uint4(0, 0, 0, 0), g_last = g_first;  // thread 0: first / last record of the next tile to copy

        auto load_geo = [&](uint32_t t) {  // thread 0: two 16-byte loads, consumed one tile later
            const uint32_t r0 = ck.rec_begin + t * TR;
            const uint32_t last = min(r0 + uint32_t(TR), ck.rec_end) - 1u;
            g_first = __ldg(reinterpret_cast<const uint4*>(a.recs + r0));
            g_last = __ldg(reinterpret_cast<const uint4*>(a.recs + last));
        };
        auto issue_copy = [&](uint32_t t, int stage) {  // thread 0
            const uint32_t r0 = ck.rec_begin + t * TR;
            const uint32_t cnt = min(uint32_t(TR), ck.rec_end - r0);
            const uint32_t a0 = (g_first.y + ck.plane_delta) & ~3u;
            const uint32_t w_end = g_last.y + ck.plane_delta + mmlst_row_words(g_last.z & 0xffffu);
            const uint32_t pbytes = (((w_end + 3u) & ~3u) - a0) * 4u;
            const uint32_t mbytes = cnt * uint32_t(sizeof(mmlst_prec));
            uint8_t* st = smem_raw + stage * stage_bytes;
            mbar_expect_tx(bars + stage, pbytes + mbytes);
            if (pbytes) bulk_g2s(st, a.planes + a0, pbytes, bars + stage);
            bulk_g2s(st + static_cast<size_t>(stage_words) * 4, a.recs + r0, mbytes, bars + stage);
        };

        WarpAcc acc;
        acc.clear();
        int cur_w = INT_MIN;

        if (threadIdx.x == 0) load_geo(0);
        __syncthreads();  // every warp is done with the previous chunk's stages
        if (threadIdx.x == 0) { issue_copy(0, 0); if (ntiles > 1) load_geo(1); }

        for (uint32_t t = 0; t < ntiles; ++t) {
            const int stage = t % NS;
            if (threadIdx.x == 0 && t + 1 < ntiles) {  // stage (t+1)%NS was last read for tile t+1-NS: behind a barrier
                issue_copy(t + 1, (t + 1) % NS);
                if (t + 2 < ntiles) load_geo(t + 2);
            }
            mbar_wait(bars + stage, (phase_bits >> stage) & 1u);
            phase_bits ^= 1u << stage;

            const uint8_t* st = smem_raw + stage * stage_bytes;
            const uint4* mt = reinterpret_cast<const uint4*>(st + static_cast<size_t>(stage_words) * 4);
            TileCtx tc;
            tc.planes = reinterpret_cast<const uint32_t*>(st);
            tc.cnt = min(uint32_t(TR), ck.rec_end - (ck.rec_begin + t * TR));
            const uint4 first = mt[0], last = mt[tc.cnt - 1];
            tc.soff_base = ck.plane_delta - ((first.y + ck.plane_delta) & ~3u);
            tc.minscore = a.minscore; tc.max_xm = a.max_xm;
            const int wlo0 = static_cast<int>(first.x) >> 5;  // sorted: the first record has the smallest pos
            // last word any record of the tile can touch (span bound), clipped to the contig
            const int whi = min((static_cast<int>(last.x) + maxspan - 1) >> 5, (int(ck.contig_len) - 1) >> 5);
            for (int wlo = wlo0; wlo <= whi; wlo += NWARP) {
                // the word of this window that this warp owns: w == wid (mod NWARP)
                const int w = wlo + ((int(wid) - wlo) & (NWARP - 1));
                if (w > whi) continue;
                if (w != cur_w) {
                    if (cur_w != INT_MIN) flush_word(acc, cur_w, ck, a.counts);
                    acc.clear();
                    cur_w = w;
                }
                const int w32 = w * 32;
                uint32_t t2[5][2], t4[5][2], t8[5][2];
#pragma unroll
                for (int pr = 0; pr < 8; ++pr) {
                    const Contrib xa = load_contrib(tc, mt, (2 * pr) * 32 + lane, w32);
                    const Contrib xb = load_contrib(tc, mt, (2 * pr + 1) * 32 + lane, w32);
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        Counter& c = acc.c[k];
                        csa(t2[k][pr & 1], c.p[0], c.p[0], xa.x[k], xb.x[k]);
                        if (pr & 1) {
                            csa(t4[k][(pr >> 1) & 1], c.p[1], c.p[1], t2[k][0], t2[k][1]);
                            if ((pr & 3) == 3) {
                                csa(t8[k][(pr >> 2) & 1], c.p[2], c.p[2], t4[k][0], t4[k][1]);
                                if (pr == 7) {
                                    uint32_t t16;
                                    csa(t16, c.p[3], c.p[3], t8[k][0], t8[k][1]);
                                    ripple16(c, t16);
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();  // the stage may be overwritten by the copy issued at the top of the next iteration
        }
        if (cur_w != INT_MIN) flush_word(acc, cur_w, ck, a.counts);
    }
}

}  // namespace

int launch_pileup_bitsliced(const PileupArgs& a, cudaStream_t stream) {
    // stage = TR rows of the largest row (+ alignment slack) + TR 16-byte records; rows too long for shared memory
    // take the atomic path
    const uint32_t stage_words = ((TR * a.max_row_words + 8u) + 31u) & ~31u;
    const size_t stage_bytes = size_t(stage_words) * 4 + TR * sizeof(mmlst_prec);
    const size_t smem = 2 * stage_bytes + 2 * sizeof(uint64_t) + 128;
    if (a.max_row_words < 3 || smem > 220 * 1024) return launch_pileup_atomic(a, stream);
    const int sms = mmlst_num_sms();
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(pileup_bitsliced_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return mmlst_cuda_fail(e, "cudaFuncSetAttribute(pileup_bitsliced_kernel)");
        configured = smem;
    }
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const uint32_t grid = a.n_chunks_dev ? uint32_t(sms * per_sm) : min(a.n_chunks, uint32_t(sms * per_sm));
    pileup_bitsliced_kernel<2><<<grid, NTHREADS, smem, stream>>>(a, stage_words);
    return mmlst_cuda_fail(cudaGetLastError(), "pileup_bitsliced_kernel");
}
