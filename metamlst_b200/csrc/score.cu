// Stage 1: per-allele scoring as a streaming segmented reduction (replaces metamlst.py:101-151, integer half).
//
// HBM-bound: 9 B/record (tid u32, as0 i16, xm3 u8, qlen u16) read exactly once with 128-bit streaming loads; each
// warp owns a contiguous range of 128-record chunks and carries a running (allele, sum, hits, first index) run in
// registers, so a coordinate-sorted stream costs one atomic triple per allele run and warp, not per record.
// Unsorted (name-grouped) streams fall back to per-lane run aggregation + atomics: still exact, slower.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t FULL = 0xffffffffu;

struct ScoreArgs {
    const uint32_t* tid; const int16_t* as0; const uint8_t* xm3; const uint16_t* qlen; const uint32_t* orig_idx;
    uint64_t n_rec; uint64_t idx_base;
    const uint8_t* allow; const uint32_t* locus_of; uint32_t n_ref;
    int minscore, max_xm, min_read_len;
    long long* sum_as; uint32_t* n_hit; uint32_t* first_idx; unsigned long long* counters;
};

__device__ __forceinline__ void flush_run(const ScoreArgs& a, uint32_t key, long long s, uint32_t c, uint32_t mn) {
    if (c) {
        atomicAdd(reinterpret_cast<unsigned long long*>(a.sum_as + key), static_cast<unsigned long long>(s));
        atomicAdd(a.n_hit + key, c);
        atomicMin(a.first_idx + key, mn);
    }
}

// R = 8 records per lane: 2 x LDG.128 (tid) + LDG.128 (as0) + LDG.64 (xm3) + LDG.128 (qlen) in flight per lane,
// 256-record warp chunks.
__global__ void __launch_bounds__(kThreads) score_kernel(const ScoreArgs a) {
    pdl_launch_dependents();
    constexpr int R = 8;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t nchunks = a.n_rec >> 8;  // full 256-record chunks
    const uint64_t per = (nchunks + nwarps - 1) / nwarps;
    const uint64_t c0 = warp * per;
    const uint64_t c1 = (c0 + per < nchunks) ? c0 + per : nchunks;
    const bool packed_ok = a.minscore >= 0;  // passing scores are >= 0: hits and sum share one REDUX

    uint32_t run_key = 0xffffffffu, run_c = 0, run_min = 0xffffffffu;
    long long run_s = 0;
    uint32_t tot = 0, ign = 0;

    for (uint64_t ch = c0; ch < c1; ++ch) {
        const uint64_t base = (ch << 8) + (lane << 3);
        const uint4 ta = ld_stream_u4(a.tid + base);
        const uint4 tb = ld_stream_u4(a.tid + base + 4);
        const uint4 a8 = ld_stream_u4(a.as0 + base);
        const uint2 x8 = ld_stream_u2(a.xm3 + base);
        const uint4 q8 = ld_stream_u4(a.qlen + base);
        uint4 oa = make_uint4(0, 0, 0, 0), ob = oa;
        if (a.orig_idx) { oa = ld_stream_u4(a.orig_idx + base); ob = ld_stream_u4(a.orig_idx + base + 4); }
        const uint32_t t[R] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
        const uint32_t aw[4] = {a8.x, a8.y, a8.z, a8.w};
        const uint32_t qw[4] = {q8.x, q8.y, q8.z, q8.w};
        const uint32_t xw[2] = {x8.x, x8.y};
        const uint32_t oi[R] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
        const uint32_t idx0 = static_cast<uint32_t>(a.idx_base + base);
        int as[R];
        bool pass[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            as[k] = (k & 1) ? (static_cast<int>(aw[k >> 1]) >> 16) : static_cast<int>(static_cast<short>(aw[k >> 1] & 0xffffu));
            const int ql = (k & 1) ? int(qw[k >> 1] >> 16) : int(qw[k >> 1] & 0xffffu);
            const int xm = int((xw[k >> 2] >> (8 * (k & 3))) & 255u);
            pass[k] = (as[k] >= a.minscore) && (ql >= a.min_read_len) && (xm <= a.max_xm);
        }
        const uint32_t k0 = __shfl_sync(FULL, t[0], 0);
        bool uni = true;
#pragma unroll
        for (int k = 0; k < R; ++k) uni = uni && (t[k] == k0);
        if (__all_sync(FULL, uni)) {
            const bool al = (k0 < a.n_ref) && a.allow[k0];
            if (al) {
                int s = 0;
                uint32_t c = 0, mn = 0xffffffffu;
#pragma unroll
                for (int k = R - 1; k >= 0; --k)
                    if (pass[k]) { s += as[k]; ++c; mn = a.orig_idx ? min(mn, oi[k]) : (idx0 + k); }
                tot += R;
                ign += R - c;
                if (packed_ok) {
                    const uint32_t pk = __reduce_add_sync(FULL, (c << 22) + static_cast<uint32_t>(s));
                    c = pk >> 22; s = static_cast<int>(pk & 0x3fffffu);
                } else {
                    s = __reduce_add_sync(FULL, s);
                    c = __reduce_add_sync(FULL, c);
                }
                mn = __reduce_min_sync(FULL, mn);
                if (k0 != run_key) {
                    if (lane == 0) flush_run(a, run_key, run_s, run_c, run_min);
                    run_key = k0; run_s = 0; run_c = 0; run_min = 0xffffffffu;
                }
                run_s += s; run_c += c; run_min = min(run_min, mn);
            }
        } else {
            // mixed chunk: flush the warp run, then every lane aggregates runs inside its R consecutive records
            if (lane == 0) flush_run(a, run_key, run_s, run_c, run_min);
            run_key = 0xffffffffu; run_s = 0; run_c = 0; run_min = 0xffffffffu;
            uint32_t lk = 0xffffffffu, lc = 0, lmin = 0xffffffffu;
            long long ls = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const uint32_t key = t[k];
                if (!((key < a.n_ref) && a.allow[key])) continue;
                ++tot;
                if (!pass[k]) { ++ign; continue; }
                if (key != lk) { if (lk != 0xffffffffu) flush_run(a, lk, ls, lc, lmin); lk = key; ls = 0; lc = 0; lmin = 0xffffffffu; }
                ls += as[k]; ++lc; lmin = min(lmin, a.orig_idx ? oi[k] : (idx0 + k));
            }
            if (lk != 0xffffffffu) flush_run(a, lk, ls, lc, lmin);
        }
    }
    if (lane == 0) flush_run(a, run_key, run_s, run_c, run_min);

    // tail (< 256 records): last warp of the grid, one record per lane per step
    if (warp == nwarps - 1) {
        for (uint64_t i = (nchunks << 8) + lane; i < a.n_rec; i += 32) {
            const uint32_t key = a.tid[i];
            if (!((key < a.n_ref) && a.allow[key])) continue;
            ++tot;
            const int as = a.as0[i];
            if ((as >= a.minscore) && (int(a.qlen[i]) >= a.min_read_len) && (int(a.xm3[i]) <= a.max_xm)) {
                const uint32_t idx = a.orig_idx ? a.orig_idx[i] : static_cast<uint32_t>(a.idx_base + i);
                flush_run(a, key, as, 1u, idx);
            } else {
                ++ign;
            }
        }
    }
    tot = __reduce_add_sync(FULL, tot);
    ign = __reduce_add_sync(FULL, ign);
    if (lane == 0 && tot) {
        atomicAdd(a.counters + 0, static_cast<unsigned long long>(tot));
        atomicAdd(a.counters + 1, static_cast<unsigned long long>(ign));
    }
}

}  // namespace

extern "C" int mmlst_score_dev(const uint32_t* tid, const int16_t* as0, const uint8_t* xm3, const uint16_t* qlen,
                               const uint32_t* orig_idx, uint64_t n_rec, uint64_t idx_base, const uint8_t* allow,
                               const uint32_t* locus_of, uint32_t n_ref, int minscore, int max_xm, int min_read_len,
                               int64_t* sum_as, uint32_t* n_hit, uint32_t* first_idx, uint64_t* counters, void* stream) {
    if (n_rec == 0) return MMLST_OK;
    if (!tid || !as0 || !xm3 || !qlen || !allow || !sum_as || !n_hit || !first_idx || !counters) {
        mmlst_set_error("mmlst_score_dev: null pointer");
        return MMLST_E_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(tid) & 15) || (reinterpret_cast<uintptr_t>(as0) & 15) ||
        (reinterpret_cast<uintptr_t>(xm3) & 7) || (reinterpret_cast<uintptr_t>(qlen) & 15) ||
        (orig_idx && (reinterpret_cast<uintptr_t>(orig_idx) & 15))) {
        mmlst_set_error("mmlst_score_dev: record arrays must be 16-byte aligned (tid/orig_idx/as0/qlen), 8 (xm3)");
        return MMLST_E_ARG;
    }
    ScoreArgs a{tid, as0, xm3, qlen, orig_idx, n_rec, idx_base, allow, locus_of, n_ref, minscore, max_xm, min_read_len,
                reinterpret_cast<long long*>(sum_as), n_hit, first_idx, reinterpret_cast<unsigned long long*>(counters)};
    const uint64_t nchunks = n_rec >> 8;
    uint64_t want = (nchunks + 7) / 8;  // CTAs if every warp took one chunk
    static int resident_by_device[MMLST_MAX_DEVICES] = {0};  // one wave exactly: the blocked chunk distribution has no tail
    int& resident = resident_by_device[mmlst_current_device()];
    if (!resident) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, score_kernel, kThreads, 0) != cudaSuccess || resident < 1) resident = 4;
    }
    const uint64_t cap = static_cast<uint64_t>(mmlst_num_sms()) * resident;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    score_kernel<<<static_cast<unsigned>(want), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}
