// Stage 1 over the RUN-LENGTH score stream (replaces metamlst.py:101-151, integer half) -- the form a coordinate-sorted BAM
// is shipped in.  The allele id of a record is not stored per record: the stream is cut into runs of equal tid
// (run_tid[r], run_start[r] .. run_start[r+1]) and every 256-record chunk carries the index of the run its first record
// belongs to (chunk_run[c]), so any chunk can be entered without a search.  What crosses PCIe and HBM per record is
// as0 i16 + xm3 u8 + qlen u16 = 5 B (the explicit-tid form of score.cu moves 9 B); the run arrays add 8 B per run and
// 4 B per 256 records.
//
// HBM-bound streaming segmented reduction: each warp owns a contiguous range of chunks, issues the loads of TWO chunks
// (6 fully coalesced 128/64-bit streaming loads per lane) before reducing either, and carries the running (allele, sum,
// hits, first index) of the open run in registers -- one atomic triple per (run, warp).  A chunk that lies inside one run
// takes the uniform path (three REDUX); a chunk crossing run boundaries is reduced segment by segment.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t FULL = 0xffffffffu;

struct RunArgs {
    const uint32_t* run_tid; const uint32_t* run_start; const uint32_t* chunk_run; uint32_t n_runs;
    const int16_t* as0; const uint8_t* xm3; const uint16_t* qlen; const uint32_t* orig_idx;
    uint64_t n_rec; uint64_t idx_base;
    const uint8_t* allow; uint32_t n_ref;
    int minscore, max_xm, min_read_len;
    long long* sum_as; uint32_t* n_hit; uint32_t* first_idx; unsigned long long* counters;
};

__device__ __forceinline__ void flush_run(const RunArgs& a, uint32_t key, long long s, uint32_t c, uint32_t mn) {
    if (c) {
        atomicAdd(reinterpret_cast<unsigned long long*>(a.sum_as + key), static_cast<unsigned long long>(s));
        atomicAdd(a.n_hit + key, c);
        atomicMin(a.first_idx + key, mn);
    }
}

template <bool OIDX> struct Loaded;  // one lane's 8 consecutive records of a chunk
template <> struct Loaded<false> { uint4 a8, q8; uint2 x8; };
template <> struct Loaded<true> { uint4 a8, q8; uint2 x8; uint4 oa, ob; };

template <bool OIDX>
__device__ __forceinline__ Loaded<OIDX> load_chunk(const RunArgs& a, uint64_t base_lane) {
    Loaded<OIDX> L;
    L.a8 = ld_stream_u4(a.as0 + base_lane);
    L.x8 = ld_stream_u2(a.xm3 + base_lane);
    L.q8 = ld_stream_u4(a.qlen + base_lane);
    if constexpr (OIDX) { L.oa = ld_stream_u4(a.orig_idx + base_lane); L.ob = ld_stream_u4(a.orig_idx + base_lane + 4); }
    return L;
}

template <bool OIDX>
__device__ __forceinline__ uint32_t rec_index(const Loaded<OIDX>& L, uint32_t idx0, int k) {
    if constexpr (OIDX) {
        const uint32_t oi[8] = {L.oa.x, L.oa.y, L.oa.z, L.oa.w, L.ob.x, L.ob.y, L.ob.z, L.ob.w};
        return oi[k];
    } else {
        return idx0 + k;
    }
}

struct WarpRun {  // warp-uniform state of the open run
    uint32_t r, key, end;  // run index, its allele, one past its last record (0xffffffff past the last run)
    bool al;               // allele passes --filter
    long long s; uint32_t c, mn;
};

__device__ __forceinline__ void open_run(const RunArgs& a, WarpRun& w, uint32_t r) {
    w.r = r;
    if (r < a.n_runs) {
        w.key = __ldg(a.run_tid + r);
        w.end = __ldg(a.run_start + r + 1);
        w.al = (w.key < a.n_ref) && a.allow[w.key];
    } else {
        w.key = 0xffffffffu; w.end = 0xffffffffu; w.al = false;
    }
    w.s = 0; w.c = 0; w.mn = 0xffffffffu;
}

__device__ __forceinline__ void close_run(const RunArgs& a, WarpRun& w, uint32_t lane) {
    if (lane == 0 && w.al) flush_run(a, w.key, w.s, w.c, w.mn);
}

// sum / count / first index of the lane's records selected by mask m (bit k = record k of the lane's 8)
template <bool OIDX>
__device__ __forceinline__ void lane_sums(const Loaded<OIDX>& L, const int (&as)[8], uint32_t m, uint32_t idx0, int& s, uint32_t& c, uint32_t& mn) {
    s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += (m & (1u << k)) ? as[k] : 0;
    c = __popc(m);
    if constexpr (OIDX) {
        mn = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 8; ++k) mn = (m & (1u << k)) ? min(mn, rec_index<true>(L, idx0, k)) : mn;
    } else {
        mn = m ? idx0 + static_cast<uint32_t>(__ffs(m) - 1) : 0xffffffffu;
    }
}

// one 256-record chunk starting at record `base` (this lane: records base + 8 lane .. + 7)
template <bool OIDX>
__device__ __forceinline__ void reduce_chunk(const RunArgs& a, WarpRun& w, const Loaded<OIDX>& L, uint64_t base, uint32_t lane,
                                             uint32_t& tot, uint32_t& ign) {
    constexpr int R = 8;
    const uint32_t aw[4] = {L.a8.x, L.a8.y, L.a8.z, L.a8.w};
    const uint32_t qw[4] = {L.q8.x, L.q8.y, L.q8.z, L.q8.w};
    const uint32_t xw[2] = {L.x8.x, L.x8.y};
    const uint64_t rec0 = base + (lane << 3);
    const uint32_t idx0 = static_cast<uint32_t>(a.idx_base + rec0);
    int as[R];
    uint32_t pass = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        as[k] = (k & 1) ? (static_cast<int>(aw[k >> 1]) >> 16) : static_cast<int>(static_cast<short>(aw[k >> 1] & 0xffffu));
        const int ql = (k & 1) ? int(qw[k >> 1] >> 16) : int(qw[k >> 1] & 0xffffu);
        const int xm = int((xw[k >> 2] >> (8 * (k & 3))) & 255u);
        pass |= ((as[k] >= a.minscore) && (ql >= a.min_read_len) && (xm <= a.max_xm)) ? (1u << k) : 0u;
    }
    const uint64_t chunk_end = base + 256;
    if (chunk_end <= w.end) {  // the whole chunk lies inside the open run
        if (w.al) {
            int s; uint32_t c, mn;
            lane_sums<OIDX>(L, as, pass, idx0, s, c, mn);
            tot += R;
            ign += R - c;
            w.s += __reduce_add_sync(FULL, s);
            w.c += __reduce_add_sync(FULL, c);
            w.mn = min(w.mn, __reduce_min_sync(FULL, mn));
        }
        if (chunk_end == w.end) { close_run(a, w, lane); open_run(a, w, w.r + 1); }
        return;
    }
    uint64_t lo = base;
    while (lo < chunk_end) {  // segment [lo, hi) of the chunk belongs to the open run
        const uint64_t hi = (w.end < chunk_end) ? static_cast<uint64_t>(w.end) : chunk_end;
        if (w.al) {
            // the lane's records inside the segment: bits [klo, khi) of its 8
            const uint32_t klo = lo <= rec0 ? 0u : (lo - rec0 >= R ? R : static_cast<uint32_t>(lo - rec0));
            const uint32_t khi = hi <= rec0 ? 0u : (hi - rec0 >= R ? R : static_cast<uint32_t>(hi - rec0));
            const uint32_t inm = khi > klo ? (((1u << khi) - 1u) & ~((1u << klo) - 1u)) : 0u;
            int s; uint32_t c, mn;
            lane_sums<OIDX>(L, as, pass & inm, idx0, s, c, mn);
            const uint32_t in = __popc(inm);
            tot += in;
            ign += in - c;
            w.s += __reduce_add_sync(FULL, s);
            w.c += __reduce_add_sync(FULL, c);
            w.mn = min(w.mn, __reduce_min_sync(FULL, mn));
        }
        if (hi == w.end) { close_run(a, w, lane); open_run(a, w, w.r + 1); }
        lo = hi;
    }
}

template <bool OIDX>
__global__ void __launch_bounds__(kThreads, 4) score_runs_kernel(const RunArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t nchunks = a.n_rec >> 8;  // full 256-record chunks
    const uint64_t per = (nchunks + nwarps - 1) / nwarps;
    const uint64_t c0 = warp * per;
    const uint64_t c1 = (c0 + per < nchunks) ? c0 + per : nchunks;
    uint32_t tot = 0, ign = 0;

    if (c0 < c1) {
        WarpRun w;
        open_run(a, w, __ldg(a.chunk_run + c0));
        uint64_t ch = c0;
        for (; ch + 1 < c1; ch += 2) {  // two chunks' loads in flight before either is reduced
            const uint64_t base = ch << 8;
            const Loaded<OIDX> L0 = load_chunk<OIDX>(a, base + (lane << 3));
            const Loaded<OIDX> L1 = load_chunk<OIDX>(a, base + 256 + (lane << 3));
            reduce_chunk<OIDX>(a, w, L0, base, lane, tot, ign);
            reduce_chunk<OIDX>(a, w, L1, base + 256, lane, tot, ign);
        }
        if (ch < c1) {
            const uint64_t base = ch << 8;
            const Loaded<OIDX> L0 = load_chunk<OIDX>(a, base + (lane << 3));
            reduce_chunk<OIDX>(a, w, L0, base, lane, tot, ign);
        }
        close_run(a, w, lane);
    }

    // tail (< 256 records): last warp of the grid, one record per lane per step; the run is found by walking from the
    // tail chunk's entry
    if (warp == nwarps - 1 && (a.n_rec & 255u)) {
        uint32_t r = __ldg(a.chunk_run + nchunks);
        for (uint64_t i = (nchunks << 8) + lane; i < a.n_rec; i += 32) {
            while (r + 1 < a.n_runs && i >= a.run_start[r + 1]) ++r;
            const uint32_t key = a.run_tid[r];
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

// tid[i] of every record from the run arrays (the coverage kernel and tests want the explicit form)
__global__ void __launch_bounds__(256) expand_runs_kernel(const uint32_t* __restrict__ run_tid, const uint32_t* __restrict__ run_start,
                                                          const uint32_t* __restrict__ chunk_run, uint32_t n_runs, uint64_t n_rec,
                                                          uint32_t* __restrict__ tid) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_rec; i += stride) {
        uint32_t r = __ldg(chunk_run + (i >> 8));
        while (r + 1 < n_runs && i >= __ldg(run_start + r + 1)) ++r;
        tid[i] = __ldg(run_tid + r);
    }
}

}  // namespace

extern "C" int mmlst_score_runs_dev(const uint32_t* run_tid, const uint32_t* run_start, uint32_t n_runs, const uint32_t* chunk_run,
                                    const int16_t* as0, const uint8_t* xm3, const uint16_t* qlen, const uint32_t* orig_idx,
                                    uint64_t n_rec, uint64_t idx_base, const uint8_t* allow, uint32_t n_ref, int minscore,
                                    int max_xm, int min_read_len, int64_t* sum_as, uint32_t* n_hit, uint32_t* first_idx,
                                    uint64_t* counters, void* stream) {
    if (n_rec == 0) return MMLST_OK;
    if (!run_tid || !run_start || !chunk_run || !n_runs || !as0 || !xm3 || !qlen || !allow || !sum_as || !n_hit || !first_idx || !counters) {
        mmlst_set_error("mmlst_score_runs_dev: null pointer");
        return MMLST_E_ARG;
    }
    if (n_rec >= 0xffffff00ull) { mmlst_set_error("mmlst_score_runs_dev: %llu records do not fit 32-bit run offsets", (unsigned long long)n_rec); return MMLST_E_RANGE; }
    if ((reinterpret_cast<uintptr_t>(as0) & 15) || (reinterpret_cast<uintptr_t>(xm3) & 7) || (reinterpret_cast<uintptr_t>(qlen) & 15) ||
        (orig_idx && (reinterpret_cast<uintptr_t>(orig_idx) & 15))) {
        mmlst_set_error("mmlst_score_runs_dev: record arrays must be 16-byte aligned (as0/qlen/orig_idx), 8 (xm3)");
        return MMLST_E_ARG;
    }
    RunArgs a{run_tid, run_start, chunk_run, n_runs, as0, xm3, qlen, orig_idx, n_rec, idx_base, allow, n_ref, minscore, max_xm,
              min_read_len, reinterpret_cast<long long*>(sum_as), n_hit, first_idx, reinterpret_cast<unsigned long long*>(counters)};
    const uint64_t nchunks = n_rec >> 8;
    uint64_t want = (nchunks + 15) / 16;  // CTAs if every warp took two chunks
    static int resident[2] = {0, 0};  // one wave exactly: the blocked chunk distribution has no tail
    const int v = orig_idx ? 1 : 0;
    if (!resident[v]) {
        const cudaError_t e = v ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident[v], score_runs_kernel<true>, kThreads, 0)
                                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident[v], score_runs_kernel<false>, kThreads, 0);
        if (e != cudaSuccess || resident[v] < 1) resident[v] = 4;
    }
    const uint64_t cap = static_cast<uint64_t>(mmlst_num_sms()) * resident[v];
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    if (v) score_runs_kernel<true><<<static_cast<unsigned>(want), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    else score_runs_kernel<false><<<static_cast<unsigned>(want), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}

extern "C" int mmlst_expand_runs_dev(const uint32_t* run_tid, const uint32_t* run_start, uint32_t n_runs, const uint32_t* chunk_run,
                                     uint64_t n_rec, uint32_t* tid, void* stream) {
    if (n_rec == 0) return MMLST_OK;
    if (!run_tid || !run_start || !chunk_run || !n_runs || !tid) { mmlst_set_error("mmlst_expand_runs_dev: null pointer"); return MMLST_E_ARG; }
    uint64_t blocks = (n_rec + 255) / 256;
    const uint64_t cap = static_cast<uint64_t>(mmlst_num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    expand_runs_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(run_tid, run_start, chunk_run, n_runs, n_rec, tid);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}
