// Microbenchmark: how long after the last CTA of kernel A has finished does the first CTA of a dependent kernel B start, as a function of A's
// grid size, dynamic shared memory per CTA and block size?  (profiles/r3k_tail_timeline.json shows 6.3 us between the last pileup CTA and the
// consensus kernel's first instruction, against 0.9 us between selection and pileup.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/tools/exit_gap_probe profiles/tools/exit_gap_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>
#include <vector>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return t; }
__global__ void kernA(unsigned long long* endt, int spin_ns, int touch) {
    extern __shared__ unsigned char sm[];
    if (touch) for (int i = threadIdx.x; i < touch; i += blockDim.x) sm[i] = (unsigned char)i;
    __syncthreads();
    const unsigned long long t0 = gt();
    while (gt() - t0 < (unsigned long long)spin_ns) { }
    __syncthreads();
    if (threadIdx.x == 0) endt[blockIdx.x] = gt();
}
// the same with the CTA's work being `rounds` bulk copies (cp.async.bulk, completion on an mbarrier) of `bytes` each instead of a spin
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void kernT(unsigned long long* endt, const unsigned char* src, int bytes, int rounds, int inval, int stagger_ns = 0, int touch = 0, unsigned* sink = nullptr) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    unsigned phase = 0;
    for (int r = 0; r < rounds; ++r) {
        if (threadIdx.x == 0) {
            size_t off = ((size_t)blockIdx.x * rounds + r) * bytes;
            if (touch & 16) off += 16 + 16 * (blockIdx.x % 7);   // 16-byte aligned only, like a plane range that starts anywhere in the stream
            if (touch & 8) {   // geometry first: two read-only 16-byte loads whose values decide the source offset
                const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(src + off)), g1 = __ldg(reinterpret_cast<const uint4*>(src + off + bytes - 16));
                off += ((g0.y + g1.y) & 0u);
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(bytes) : "memory");
            if (touch & 4) {   // two copies on one barrier
                const int b0 = bytes / 6 * 5 / 16 * 16, b1 = bytes - b0;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(sm)), "l"(src + off), "r"(b0), "r"(su32(&bar)) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(sm + b0)), "l"(src + off + b0), "r"(b1), "r"(su32(&bar)) : "memory");
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(sm)), "l"(src + off), "r"(bytes), "r"(su32(&bar)) : "memory");
            }
        }
        unsigned done;
        do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(su32(&bar)), "r"(phase) : "memory"); } while (!done);
        phase ^= 1;
        if (touch & 1) {   // read what the copy brought with ordinary shared-memory loads, write a little back
            unsigned acc = 0;
            for (int i = threadIdx.x * 16; i + 16 <= bytes; i += blockDim.x * 16) { const uint4 v = *reinterpret_cast<const uint4*>(sm + i); acc += v.x ^ v.y ^ v.z ^ v.w; }
            reinterpret_cast<unsigned*>(sm + 100 * 1024)[threadIdx.x] = acc;
            if (acc == 0x12345678u && sink) sink[0] = acc;
        }
        __syncthreads();
    }
    if (stagger_ns) { const unsigned long long t0 = gt(); const unsigned long long d = (unsigned long long)(blockIdx.x % 16) * stagger_ns / 16; while (gt() - t0 < d) { } }
    if (inval && threadIdx.x == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(su32(&bar)) : "memory");
    if (threadIdx.x == 0) endt[blockIdx.x] = gt();
}
// kernT with ~128 live registers per thread (two such CTAs fill the register file of an SM, like the pileup kernel's)
__global__ void __launch_bounds__(256, 2) kernTR(unsigned long long* endt, const unsigned char* src, int bytes, int rounds, unsigned* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned r[96];
#pragma unroll
    for (int i = 0; i < 96; ++i) r[i] = threadIdx.x * 2654435761u + i;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    unsigned phase = 0;
    for (int k = 0; k < rounds; ++k) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(sm)),
                         "l"(src + ((size_t)blockIdx.x * rounds + k) * bytes), "r"(bytes), "r"(su32(&bar)) : "memory");
        }
        unsigned done;
        do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(su32(&bar)), "r"(phase) : "memory"); } while (!done);
        phase ^= 1;
        // bit-sliced style work on what arrived: every register stays live across the rounds
        for (int i = threadIdx.x * 4; i + 4 <= bytes; i += blockDim.x * 4 * 8) {
            const unsigned v = *reinterpret_cast<const unsigned*>(sm + i);
#pragma unroll
            for (int j = 0; j < 96; ++j) r[j] = (r[j] ^ v) + (r[(j + 1) % 96] & v);
        }
        __syncthreads();
    }
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 96; ++i) acc ^= r[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) endt[blockIdx.x] = gt();
}
__global__ void kernB(unsigned long long* startt) { if (threadIdx.x == 0) startt[blockIdx.x] = gt(); }
// one measurement: [marker kernel, kernel under test, marker kernel] replayed as a CUDA graph (kernel -> kernel edges without the host launch latency that
// hides completion latency), cold L2; prints how long the kernel's CTAs ran and how long after the last of them the next kernel's first CTA started
template <class Launch>
static void measure(const char* what, cudaStream_t s, unsigned long long* dA, unsigned long long* dB, unsigned char* flush, int grid, Launch launch) {
    cudaGraphExec_t ge = nullptr;
    std::vector<double> gaps, durs;
    for (int rep = 0; rep < 12; ++rep) {
        cudaMemsetAsync(dA, 0, 4096 * 8, s); cudaMemsetAsync(dB, 0, 4096 * 8, s);
        cudaMemsetAsync(flush, rep, 256u << 20, s);
        if (!ge) {
            cudaGraph_t g;
            cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
            kernB<<<21, 512, 0, s>>>(dB + 1024);
            launch();
            kernB<<<21, 512, 0, s>>>(dB);
            cudaStreamEndCapture(s, &g);
            cudaGraphInstantiate(&ge, g, 0);
        }
        cudaGraphLaunch(ge, s);
        cudaStreamSynchronize(s);
        std::vector<unsigned long long> a(grid), b(21), b0(21);
        cudaMemcpy(a.data(), dA, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), dB, 21 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b0.data(), dB + 1024, 21 * 8, cudaMemcpyDeviceToHost);
        const unsigned long long ea = *std::max_element(a.begin(), a.end()), sb = *std::min_element(b.begin(), b.end()), s0 = *std::min_element(b0.begin(), b0.end());
        if (rep >= 2) { gaps.push_back((double)(sb - ea) / 1e3); durs.push_back((double)(ea - s0) / 1e3); }
    }
    std::sort(gaps.begin(), gaps.end()); std::sort(durs.begin(), durs.end());
    printf("%-92s previous kernel's entry -> last CTA end %6.2f us | last CTA end -> next kernel's entry  min %.2f median %.2f max %.2f us\n", what, durs[durs.size() / 2],
           gaps.front(), gaps[gaps.size() / 2], gaps.back());
}

int main() {
    unsigned long long *dA, *dB;
    cudaMalloc(&dA, 4096 * 8); cudaMalloc(&dB, 4096 * 8);
    unsigned char *src, *flush;
    cudaMalloc(&src, (size_t)296 * 3 * 49152 + 4096); cudaMemset(src, 1, (size_t)296 * 3 * 49152 + 4096);
    cudaMalloc(&flush, 256u << 20);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaFuncSetAttribute(kernA, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (auto k : {(const void*)kernT, (const void*)kernTR}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    int occT = 0, occR = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occT, kernT, 256, 110 * 1024);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occR, kernTR, 256, 111 * 1024);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kernTR);
    printf("resident CTAs per SM: bulk-copy kernel %d, register-heavy kernel %d (%d registers)\n", occT, occR, fa.numRegs);
    unsigned* sink = (unsigned*)dB + 4000;
    // eager launches first: the host launch latency (~3 us) hides everything below it
    for (int grid : {21, 296}) for (int smem : {0, 110 * 1024}) {
        std::vector<double> gaps;
        for (int rep = 0; rep < 12; ++rep) {
            cudaMemsetAsync(dA, 0, 4096 * 8, s); cudaMemsetAsync(dB, 0, 4096 * 8, s);
            kernA<<<grid, 256, smem, s>>>(dA, 5000, smem);
            kernB<<<21, 512, 0, s>>>(dB);
            cudaStreamSynchronize(s);
            std::vector<unsigned long long> a(grid), b(21);
            cudaMemcpy(a.data(), dA, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), dB, 21 * 8, cudaMemcpyDeviceToHost);
            if (rep >= 2) gaps.push_back((double)(*std::min_element(b.begin(), b.end()) - *std::max_element(a.begin(), a.end())) / 1e3);
        }
        std::sort(gaps.begin(), gaps.end());
        printf("EAGER 5 us spin, grid %3d, %3d KB shared memory per CTA: last CTA end -> next kernel's entry median %.2f us\n", grid, smem >> 10, gaps[gaps.size() / 2]);
    }
    const int B = 49152;
    measure("GRAPH 296 CTAs x 3 bulk copies of 48 KB, waited one by one", s, dA, dB, flush, 296, [&] { kernT<<<296, 256, 110 * 1024, s>>>(dA, src, B, 3, 0, 0, 0, sink); });
    measure("GRAPH ... + the stage read with ordinary loads, two copies per barrier, geometry loads first", s, dA, dB, flush, 296, [&] { kernT<<<296, 256, 110 * 1024, s>>>(dA, src, B, 3, 0, 0, 13, sink); });
    measure("GRAPH ... + CTAs ending up to 8 us apart", s, dA, dB, flush, 296, [&] { kernT<<<296, 256, 110 * 1024, s>>>(dA, src, B, 3, 0, 8000, 13, sink); });
    measure("GRAPH ... + sources aligned to 16 bytes only, sizes not a multiple of 128", s, dA, dB, flush, 296, [&] { kernT<<<296, 256, 110 * 1024, s>>>(dA, src, B - 208, 3, 0, 0, 13 + 16, sink); });
    measure("GRAPH ... + mbarrier invalidated before exit", s, dA, dB, flush, 296, [&] { kernT<<<296, 256, 110 * 1024, s>>>(dA, src, B, 3, 1, 0, 13, sink); });
    measure("GRAPH 296 CTAs, 120 registers, 2 per SM, bit-sliced style work on what the copies bring", s, dA, dB, flush, 296, [&] { kernTR<<<296, 256, 111 * 1024, s>>>(dA, src, B, 3, sink); });
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
