// Owner-mode exchange over NVLink peer memory: every GPU ends a pass holding every GPU's result block.
//
// Replaces the NCCL all-gather of the pass (pipeline.py, exchange="gather") by two tiny kernels on the pass's own
// stream, working on a SYMMETRIC buffer (same layout on every GPU, each GPU holds the peer addresses of all the others):
//
//     [ half 0 : world slots x slot_stride ][ half 1 : world slots x slot_stride ][ flag[world] (u64) ]
//
//   publish : CTA p stores this GPU's block into slot[rank] of half (e & 1) of GPU p with 128-bit stores through the
//             NVLink peer mapping, fences at system scope and releases flag[rank] = e on GPU p.  Nothing is read from a
//             peer and nothing waits: the block leaves as soon as the consensus kernel has written it.
//   await   : CTA r acquires the LOCAL flag[r] >= e, then copies slot r of half (e & 1) into the contiguous `out_all`
//             the D2H copy reads.  The last CTA advances the pass counter e.
//
// e (the pass number) lives in device memory, so both kernels are CUDA-graph nodes with frozen arguments.  Halves
// alternate with e: a peer can run at most ONE pass ahead of this GPU's await (its publish(e+2) is ordered after its
// await(e+1), which needs this GPU's publish(e+1), which this GPU's stream orders after its own await(e)), so the half
// being read is never overwritten.  Flags only grow, a late reader never misses one.
// The wait is bounded (about 2 s of SM clocks): on expiry bit 2 is set in `status` and the pass ends with a flagged
// result instead of hanging the GPU.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct XchgArgs {
    const unsigned long long* peer_base;  // [world] device-visible address of every GPU's symmetric buffer
    uint32_t rank, world;
    unsigned long long half_stride, slot_stride, flag_off;
    uint32_t bytes;  // multiple of 16
};

__global__ void __launch_bounds__(kThreads) publish_kernel(XchgArgs a, const uint4* __restrict__ block, const unsigned long long* epoch) {
    const unsigned long long e = *epoch + 1ull;
    const uint32_t p = blockIdx.x;
    char* base = reinterpret_cast<char*>(a.peer_base[p]);
    uint4* dst = reinterpret_cast<uint4*>(base + (e & 1ull) * a.half_stride + a.rank * a.slot_stride);
    const uint32_t n16 = a.bytes >> 4;
    for (uint32_t i = threadIdx.x; i < n16; i += kThreads) dst[i] = block[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(reinterpret_cast<unsigned long long*>(base + a.flag_off) + a.rank, e);
}

__global__ void __launch_bounds__(kThreads) await_kernel(XchgArgs a, char* local_base, uint4* __restrict__ out_all, unsigned long long* epoch,
                                                         uint32_t* ticket, uint32_t* status) {
    __shared__ int s_ok;
    const unsigned long long e = *epoch + 1ull;
    const uint32_t r = blockIdx.x;
    if (threadIdx.x == 0) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(local_base + a.flag_off) + r;
        const long long t0 = clock64();
        int ok = 1;
        while (ld_acquire_sys(flag) < e) {
            if (clock64() - t0 > 4000000000ll) { ok = 0; break; }
            __nanosleep(64);
        }
        s_ok = ok;
    }
    __syncthreads();
    if (s_ok) {
        const uint4* src = reinterpret_cast<const uint4*>(local_base + (e & 1ull) * a.half_stride + r * a.slot_stride);
        uint4* dst = out_all + static_cast<size_t>(r) * (a.bytes >> 4);
        const uint32_t n16 = a.bytes >> 4;
        for (uint32_t i = threadIdx.x; i < n16; i += kThreads) dst[i] = __ldcg(src + i);  // L2 is where the peer's stores landed
    } else if (threadIdx.x == 0) {
        atomicOr(status, 4u);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(ticket, 1u) == a.world - 1u) {  // every CTA has read `epoch` by now
            *ticket = 0u;
            *epoch = e;
        }
    }
}

}  // namespace

extern "C" int mmlst_xchg_publish_dev(const void* block, uint32_t bytes, const uint64_t* peer_base, uint32_t rank, uint32_t world,
                                      uint64_t half_stride, uint64_t slot_stride, uint64_t flag_off, const uint64_t* epoch, void* stream) {
    if (!block || !peer_base || !epoch || world == 0 || rank >= world) { mmlst_set_error("mmlst_xchg_publish_dev: bad argument"); return MMLST_E_ARG; }
    if ((bytes & 15u) || bytes > slot_stride || (slot_stride & 15u) || (half_stride & 15u) || (flag_off & 7u)) {
        mmlst_set_error("mmlst_xchg_publish_dev: block size / strides must be 16-byte multiples and fit a slot");
        return MMLST_E_ARG;
    }
    XchgArgs a{reinterpret_cast<const unsigned long long*>(peer_base), rank, world, half_stride, slot_stride, flag_off, bytes};
    static bool carve_p[MMLST_MAX_DEVICES] = {false};
    mmlst_prefer_max_shared(publish_kernel, carve_p);
    publish_kernel<<<world, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, static_cast<const uint4*>(block),
                                                                              reinterpret_cast<const unsigned long long*>(epoch));
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}

extern "C" int mmlst_xchg_await_dev(void* local_base, uint32_t bytes, uint32_t world, uint64_t half_stride, uint64_t slot_stride,
                                    uint64_t flag_off, void* out_all, uint64_t* epoch, uint32_t* ticket, uint32_t* status, void* stream) {
    if (!local_base || !out_all || !epoch || !ticket || !status || world == 0) { mmlst_set_error("mmlst_xchg_await_dev: bad argument"); return MMLST_E_ARG; }
    if ((bytes & 15u) || bytes > slot_stride) { mmlst_set_error("mmlst_xchg_await_dev: block size must be a 16-byte multiple and fit a slot"); return MMLST_E_ARG; }
    XchgArgs a{nullptr, 0, world, half_stride, slot_stride, flag_off, bytes};
    static bool carve_a[MMLST_MAX_DEVICES] = {false};
    mmlst_prefer_max_shared(await_kernel, carve_a);
    await_kernel<<<world, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, static_cast<char*>(local_base), static_cast<uint4*>(out_all),
                                                                            reinterpret_cast<unsigned long long*>(epoch), ticket, status);
    CUDA_TRY(cudaGetLastError());
    return MMLST_OK;
}
