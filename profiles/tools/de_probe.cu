// Probe of Blackwell's hardware decompression engine through the CUDA 12.8+ driver API (cuMemBatchDecompressAsync,
// CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE): is raw DEFLATE (the payload of a BGZF block) supported on this device, are the
// results byte-identical to zlib's, and how fast is a batch of <= 64 KB blocks?  Profiling aid only -- not linked into
// libmmlst.so.  Build: nvcc -O2 -o de_probe de_probe.cu -lz -lcuda (libcuda is resolved with dlopen, so the stub is enough).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*decomp_fn)(CUmemDecompressParams*, size_t, unsigned int, size_t*, CUstream);

int main(int argc, char** argv) {
    const int n_blocks = argc > 1 ? atoi(argv[1]) : 4096;
    const int src_align = argc > 2 ? atoi(argv[2]) : 8;    // 1: payloads packed as in a BGZF file (26 bytes of header / trailer between them)
    const int raw = argc > 3 ? atoi(argv[3]) : 65280;      // odd sizes make the destinations unaligned too
    cudaFree(0);
    int dev = 0;
    cudaGetDevice(&dev);
    void* h = dlopen("libcuda.so.1", RTLD_NOW);
    if (!h) { printf("{\"error\": \"no libcuda.so.1\"}\n"); return 0; }
    auto attr = (CUresult(*)(int*, CUdevice_attribute, CUdevice))dlsym(h, "cuDeviceGetAttribute");
    int mask = -1, maxlen = -1;
    CUresult r1 = attr(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, dev);
    CUresult r2 = attr(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, dev);
    decomp_fn fn = (decomp_fn)dlsym(h, "cuMemBatchDecompressAsync_ptsz");
    if (!fn) fn = (decomp_fn)dlsym(h, "cuMemBatchDecompressAsync");
    printf("{\"src_align\": %d, \"raw\": %d, \"attr_rc\": [%d, %d], \"algorithm_mask\": %d, \"max_length\": %d, \"have_entry_point\": %d", src_align, raw, (int)r1, (int)r2, mask, maxlen, fn ? 1 : 0);
    if (!fn || !(mask & 1)) { printf("}\n"); return 0; }

    // BAM-like payload: low-entropy structured bytes (compresses ~2.5x like a real BAM)
    std::vector<uint8_t> plain((size_t)n_blocks * raw);
    uint32_t s = 12345;
    for (size_t i = 0; i < plain.size(); ++i) {
        s = s * 1664525u + 1013904223u;
        const uint32_t k = (uint32_t)(i % 320);
        plain[i] = k < 48 ? (uint8_t)(k * 7 + (i / 320) % 5) : k < 123 ? (uint8_t)("\x11\x12\x14\x18\x21\x22\x24\x28"[(s >> 24) & 7]) : (uint8_t)(30 + ((s >> 20) & 7));
    }
    std::vector<uint8_t> comp;
    std::vector<size_t> coff(n_blocks + 1, 0);
    std::vector<uint8_t> tmp(compressBound(raw) + 64);
    std::vector<size_t> clen_tmp;
    for (int b = 0; b < n_blocks; ++b) {
        z_stream z; memset(&z, 0, sizeof z);
        deflateInit2(&z, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        z.next_in = plain.data() + (size_t)b * raw; z.avail_in = raw;
        z.next_out = tmp.data(); z.avail_out = (uInt)tmp.size();
        deflate(&z, Z_FINISH);
        size_t n = tmp.size() - z.avail_out;
        deflateEnd(&z);
        size_t pad = src_align > 1 ? ((comp.size() + src_align - 1) / src_align) * src_align : comp.size() + 18;
        comp.resize(pad);
        coff[b] = pad;
        comp.insert(comp.end(), tmp.begin(), tmp.begin() + n);
        clen_tmp.push_back(n);
        if (src_align <= 1) comp.resize(comp.size() + 8);   // CRC32 + ISIZE trailer
        coff[b + 1] = comp.size();
    }
    std::vector<size_t> clen(n_blocks);
    uint8_t *d_comp, *d_out; uint32_t* d_act;
    cudaMalloc(&d_comp, comp.size() + 64);
    cudaMalloc(&d_out, plain.size());
    cudaMalloc(&d_act, sizeof(uint32_t) * n_blocks);
    cudaMemcpy(d_comp, comp.data(), comp.size(), cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0, plain.size());
    std::vector<CUmemDecompressParams> prm(n_blocks);
    memset(prm.data(), 0, sizeof(CUmemDecompressParams) * n_blocks);
    for (int b = 0; b < n_blocks; ++b) {
        const size_t end = (b + 1 < n_blocks) ? coff[b + 1] : comp.size();
        size_t real_end = end;  // exact size: the next block's padding start is >= this block's end
        (void)real_end;
        prm[b].srcNumBytes = clen_tmp[b];
        prm[b].dstNumBytes = raw;
        prm[b].dstActBytes = d_act + b;
        prm[b].src = d_comp + coff[b];
        prm[b].dst = d_out + (size_t)b * raw;
        prm[b].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
    }
    cudaStream_t st; cudaStreamCreate(&st);
    size_t err_idx = (size_t)-1;
    CUresult rc = fn(prm.data(), n_blocks, 0, &err_idx, (CUstream)st);
    cudaError_t se = cudaStreamSynchronize(st);
    printf(", \"submit_rc\": %d, \"sync_rc\": %d, \"err_index\": %lld", (int)rc, (int)se, (long long)err_idx);
    if (rc == CUDA_SUCCESS && se == cudaSuccess) {
        std::vector<uint8_t> back(plain.size());
        std::vector<uint32_t> act(n_blocks);
        cudaMemcpy(back.data(), d_out, plain.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(act.data(), d_act, sizeof(uint32_t) * n_blocks, cudaMemcpyDeviceToHost);
        size_t bad = 0, badlen = 0;
        for (int b = 0; b < n_blocks; ++b) badlen += act[b] != (uint32_t)raw;
        for (size_t i = 0; i < plain.size(); ++i) bad += back[i] != plain[i];
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0, st);
            fn(prm.data(), n_blocks, 0, &err_idx, (CUstream)st);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf(", \"blocks\": %d, \"plain_bytes\": %zu, \"compressed_bytes\": %zu, \"mismatched_bytes\": %zu, \"blocks_with_wrong_length\": %zu, "
               "\"best_ms\": %.4f, \"out_GBps\": %.2f, \"in_GBps\": %.2f",
               n_blocks, plain.size(), comp.size(), bad, badlen, best, plain.size() / best / 1e6, comp.size() / best / 1e6);
    }
    printf("}\n");
    return 0;
}
