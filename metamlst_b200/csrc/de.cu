// see de.h
#include "de.h"

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace {

typedef CUresult (*decomp_fn)(CUmemDecompressParams*, size_t, unsigned int, size_t*, CUstream);
constexpr int kSlices = 32;   // upper bound; the count used is slices_wanted()
int slices_wanted(int asked) {   // slices a buffer is cut into; MMLST_DE_SLICES=<1..32> overrides the caller's choice (a tuning knob, read per call)
    const char* e = getenv("MMLST_DE_SLICES");
    const int n = e ? atoi(e) : asked;
    return n < 1 ? 1 : (n > kSlices ? kSlices : n);
}
struct Lane { std::mutex m; cudaStream_t s = nullptr; cudaEvent_t ev[kSlices + 2]; bool ready = false; decomp_fn fn = nullptr; int state = 0; };   // state: 0 unknown, 1 ok, -1 absent
Lane g_lane[MMLST_MAX_DEVICES];

int resolve(int device, Lane& lane) {
    if (lane.state == 1) return MMLST_OK;
    int mask = 0;
    CUdevice cudev;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuDeviceGet", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { cudaGetLastError(); mmlst_set_error("hardware decompression: no driver entry points"); return MMLST_E_CUDA; }
    reinterpret_cast<CUresult (*)(CUdevice*, int)>(fn)(&cudev, device);
    if (cudaGetDriverEntryPoint("cuDeviceGetAttribute", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { cudaGetLastError(); mmlst_set_error("hardware decompression: no driver entry points"); return MMLST_E_CUDA; }
    reinterpret_cast<CUresult (*)(int*, CUdevice_attribute, CUdevice)>(fn)(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, cudev);
    if (!(mask & CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE)) {
        mmlst_set_error("this device has no hardware DEFLATE decompression (algorithm mask %d)", mask);
        lane.state = -1;
        return MMLST_E_CUDA;
    }
    if (cudaGetDriverEntryPoint("cuMemBatchDecompressAsync", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        cudaGetLastError();
        mmlst_set_error("driver has no cuMemBatchDecompressAsync");
        lane.state = -1;
        return MMLST_E_CUDA;
    }
    lane.fn = reinterpret_cast<decomp_fn>(fn);
    lane.state = 1;
    return MMLST_OK;
}

}  // namespace

int mmlst_de_available(int device) {
    Lane& lane = g_lane[device % MMLST_MAX_DEVICES];
    std::lock_guard<std::mutex> g(lane.m);
    if (lane.state == -1) { mmlst_set_error("this device has no hardware DEFLATE decompression"); return MMLST_E_CUDA; }
    return resolve(device, lane);
}

int mmlst_h2d_inflate(int device, cudaStream_t st, uint8_t* d_comp, const uint8_t* h_comp, size_t n_bytes,
                      std::vector<CUmemDecompressParams>& prm, const std::vector<uint64_t>& src_off,
                      const std::vector<MmlstPlainCopy>* plain, int slices) {
    Lane& lane = g_lane[device % MMLST_MAX_DEVICES];
    std::lock_guard<std::mutex> g(lane.m);
    if (lane.state != 1) { const int rc = resolve(device, lane); if (rc != MMLST_OK) return rc; }
    if (!lane.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&lane.s, cudaStreamNonBlocking));
        for (auto& e : lane.ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        lane.ready = true;
    }
    const uint32_t nb = static_cast<uint32_t>(prm.size());
    const bool has_plain = plain && !plain->empty();
    if (nb == 0 && !has_plain) return MMLST_OK;
    CUDA_TRY(cudaEventRecord(lane.ev[kSlices], st));
    CUDA_TRY(cudaStreamWaitEvent(lane.s, lane.ev[kSlices], 0));   // d_comp may still be in use by earlier work of `st`
    const uint32_t n_slices = static_cast<uint32_t>(slices_wanted(slices));
    const uint32_t per = nb ? (nb + n_slices - 1) / n_slices : 1u;
    size_t lo = 0;
    for (uint32_t c = 0, b0 = 0; b0 < nb; ++c, b0 += per) {
        const uint32_t b1 = std::min(nb, b0 + per);
        const size_t hi = (b1 == nb) ? n_bytes : static_cast<size_t>(src_off[b1]);   // up to the next slice's first payload byte
        CUDA_TRY(cudaMemcpyAsync(d_comp + lo, h_comp + lo, hi - lo, cudaMemcpyHostToDevice, lane.s));
        CUDA_TRY(cudaEventRecord(lane.ev[c], lane.s));
        CUDA_TRY(cudaStreamWaitEvent(st, lane.ev[c], 0));
        lo = hi;
        for (uint32_t q0 = b0; q0 < b1; q0 += 1u << 16) {
            const size_t cnt = std::min<size_t>(1u << 16, b1 - q0);
            size_t bad = static_cast<size_t>(-1);
            const CUresult rc = lane.fn(prm.data() + q0, cnt, 0, &bad, reinterpret_cast<CUstream>(st));
            if (rc != CUDA_SUCCESS) {
                mmlst_set_error("cuMemBatchDecompressAsync failed (CUresult %d) at block %lld", static_cast<int>(rc),
                                bad == static_cast<size_t>(-1) ? -1ll : static_cast<long long>(q0 + bad));
                return MMLST_E_CUDA;
            }
        }
    }
    if (has_plain) {
        for (const MmlstPlainCopy& pc : *plain)
            if (pc.bytes) CUDA_TRY(cudaMemcpyAsync(pc.dst, pc.src, pc.bytes, cudaMemcpyHostToDevice, lane.s));
        CUDA_TRY(cudaEventRecord(lane.ev[kSlices + 1], lane.s));
        CUDA_TRY(cudaStreamWaitEvent(st, lane.ev[kSlices + 1], 0));
    }
    mmlst_trace_mark("slices_enqueued");
    CUDA_TRY(cudaStreamSynchronize(lane.s));   // the lane (and its events) is free for the next call; the host buffer has been read
    mmlst_trace_mark("copies_done");
    return MMLST_OK;
}


int mmlst_h2d_inflate_segments(int device, cudaStream_t st, const std::vector<MmlstSegment>& segs, std::vector<CUmemDecompressParams>& prm,
                               const std::vector<uint32_t>& first_block, int groups) {
    Lane& lane = g_lane[device % MMLST_MAX_DEVICES];
    std::lock_guard<std::mutex> g(lane.m);
    if (lane.state != 1) { const int rc = resolve(device, lane); if (rc != MMLST_OK) return rc; }
    if (!lane.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&lane.s, cudaStreamNonBlocking));
        for (auto& e : lane.ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        lane.ready = true;
    }
    const uint32_t ns = static_cast<uint32_t>(segs.size());
    if (ns == 0) return MMLST_OK;
    if (first_block.size() != segs.size() + 1 || first_block.back() != prm.size()) { mmlst_set_error("mmlst_h2d_inflate_segments: block ranges do not match"); return MMLST_E_ARG; }
    CUDA_TRY(cudaEventRecord(lane.ev[kSlices], st));
    CUDA_TRY(cudaStreamWaitEvent(lane.s, lane.ev[kSlices], 0));   // the destination buffers may still be in use by earlier work of `st`
    const uint32_t n_groups = static_cast<uint32_t>(std::min<int>(std::max(groups, 1), kSlices));
    // groups of about equal compressed size, whole segments each
    size_t total = 0;
    for (const MmlstSegment& sg : segs) total += sg.bytes;
    const size_t per = (total + n_groups - 1) / n_groups;
    uint32_t i = 0, ev = 0;
    while (i < ns) {
        const uint32_t i0 = i;
        size_t acc = 0;
        while (i < ns && (acc == 0 || acc + segs[i].bytes <= per || ev + 1 == n_groups)) {
            if (segs[i].bytes) CUDA_TRY(cudaMemcpyAsync(segs[i].d_dst, segs[i].h_src, segs[i].bytes, cudaMemcpyHostToDevice, lane.s));
            acc += segs[i].bytes;
            ++i;
        }
        CUDA_TRY(cudaEventRecord(lane.ev[ev], lane.s));
        CUDA_TRY(cudaStreamWaitEvent(st, lane.ev[ev], 0));
        ev = std::min(ev + 1, n_groups - 1);
        const uint32_t b0 = first_block[i0], b1 = first_block[i];
        for (uint32_t q0 = b0; q0 < b1; q0 += 1u << 16) {
            const size_t cnt = std::min<size_t>(1u << 16, b1 - q0);
            size_t bad = static_cast<size_t>(-1);
            const CUresult rc = lane.fn(prm.data() + q0, cnt, 0, &bad, reinterpret_cast<CUstream>(st));
            if (rc != CUDA_SUCCESS) {
                mmlst_set_error("cuMemBatchDecompressAsync failed (CUresult %d) at block %lld", static_cast<int>(rc),
                                bad == static_cast<size_t>(-1) ? -1ll : static_cast<long long>(q0 + bad));
                return MMLST_E_CUDA;
            }
        }
    }
    mmlst_trace_mark("segments_enqueued");
    CUDA_TRY(cudaStreamSynchronize(lane.s));   // the lane is free for the next call; the host buffer has been read
    mmlst_trace_mark("segment_copies_done");
    return MMLST_OK;
}
