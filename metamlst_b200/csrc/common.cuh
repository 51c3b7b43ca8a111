// Shared helpers for libmmlst (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mmlst.h"

#define MMLST_NUM_SMS_DEFAULT 148

void mmlst_set_error(const char* fmt, ...);
int mmlst_cuda_fail(cudaError_t e, const char* what);

#define CUDA_TRY(expr)                                          \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return mmlst_cuda_fail(_e, #expr); \
    } while (0)

int mmlst_num_sms();

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// streaming 128-bit load: read once, do not pollute L1
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u1(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
