// CUDA vocabulary for compiling a kernel's text with g++ and running its warp program on the host: one std::thread per
// lane, warp collectives and barriers as real rendezvous between those threads, TMA bulk copies as memcpy completing an
// emulated mbarrier.  TEST INFRASTRUCTURE ONLY (tests/test_simt_score.py): it checks control flow and integer arithmetic of
// the kernels against numpy without a GPU; it says nothing about memory ordering or performance.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__
#define __shared__

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3e { unsigned x = 1, y = 1, z = 1; };
extern thread_local dim3e threadIdx, blockIdx;
extern dim3e blockDim, gridDim;

using std::max;
using std::min;

// ---- rendezvous of a fixed set of threads (reusable)
class Rendezvous {
  public:
    explicit Rendezvous(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned g = gen_;
        if (++cnt_ == n_) { cnt_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != g; });
    }
  private:
    std::mutex m_; std::condition_variable cv_; int n_, cnt_ = 0; unsigned gen_ = 0;
};
struct WarpEmu { Rendezvous rv{32}; unsigned long long slot[32]; };
struct BlockEmu { Rendezvous* rv; };
extern thread_local WarpEmu* t_warp;
extern thread_local BlockEmu* t_block;

inline void __syncthreads() { t_block->rv->wait(); }
inline void __syncwarp() { t_warp->rv.wait(); }
template <class T, class F>
inline T warp_fold(T v, F f) {
    WarpEmu* w = t_warp;
    w->slot[threadIdx.x & 31u] = static_cast<unsigned long long>(v);
    w->rv.wait();
    T r = static_cast<T>(w->slot[0]);
    for (int i = 1; i < 32; ++i) r = f(r, static_cast<T>(w->slot[i]));
    w->rv.wait();
    return r;
}
inline int __reduce_add_sync(uint32_t, int v) { return warp_fold<int>(v, [](int a, int b) { return a + b; }); }
inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) { return warp_fold<uint32_t>(v, [](uint32_t a, uint32_t b) { return a + b; }); }
inline uint32_t __reduce_min_sync(uint32_t, uint32_t v) { return warp_fold<uint32_t>(v, [](uint32_t a, uint32_t b) { return a < b ? a : b; }); }
inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) { return warp_fold<uint32_t>(v, [](uint32_t a, uint32_t b) { return a > b ? a : b; }); }

// ---- scalar intrinsics
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __ffs(uint32_t x) { return __builtin_ffs(static_cast<int>(x)); }
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    const unsigned long long v = (static_cast<unsigned long long>(b) << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= static_cast<uint32_t>((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
// dp2a: the two signed 16-bit halves of a times the two low (lo) / high (hi) signed bytes of b, plus c
inline int __dp2a_lo(int a, int b, int c) {
    return c + static_cast<int16_t>(a & 0xffff) * static_cast<int8_t>(b & 0xff) + static_cast<int16_t>((a >> 16) & 0xffff) * static_cast<int8_t>((b >> 8) & 0xff);
}
inline int __dp2a_hi(int a, int b, int c) {
    return c + static_cast<int16_t>(a & 0xffff) * static_cast<int8_t>((b >> 16) & 0xff) + static_cast<int16_t>((a >> 16) & 0xffff) * static_cast<int8_t>((b >> 24) & 0xff);
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ---- streaming loads
inline uint4 ld_stream_u4(const void* p) { uint4 r; memcpy(&r, p, 16); return r; }
inline uint2 ld_stream_u2(const void* p) { uint2 r; memcpy(&r, p, 8); return r; }
inline uint32_t ld_stream_u1(const void* p) { uint32_t r; memcpy(&r, p, 4); return r; }

inline uint64_t l2_policy(int kind) { return static_cast<uint64_t>(kind); }
inline uint4 ld_stream_u4_hint(const void* p, uint64_t) { return ld_stream_u4(p); }
inline uint2 ld_stream_u2_hint(const void* p, uint64_t) { return ld_stream_u2(p); }
inline uint32_t ld_table_u32(const uint32_t* p, uint64_t) { return *p; }
inline uint32_t ld_table_u16(const uint16_t* p, uint64_t) { return *p; }
inline uint32_t ld_table_u8(const uint8_t* p, uint64_t) { return *p; }
inline uint32_t __shfl_sync(uint32_t, uint32_t v, uint32_t src) {
    WarpEmu* w = t_warp;
    w->slot[threadIdx.x & 31u] = v;
    w->rv.wait();
    const uint32_t r = static_cast<uint32_t>(w->slot[src & 31u]);
    w->rv.wait();
    return r;
}

// ---- mbarrier + bulk copy.  The 64-bit barrier word holds: completed phases (low 32 bits are enough here) -- the
// arrival count is 1 in every kernel that uses these, so a phase completes when its expected bytes have landed.
struct MbarEmu { std::atomic<uint32_t> phases; std::atomic<int64_t> tx; std::atomic<uint32_t> arrived; uint32_t count; };
// the kernels reserve 8 bytes per barrier; the emulation keeps its state in a side table keyed by that address
MbarEmu* mbar_emu_of(uint64_t* bar);
void mbar_emu_reset();
inline void mbar_try_complete(MbarEmu* m) {
    if (m->arrived.load() >= m->count && m->tx.load() == 0) { m->arrived.store(0); m->phases.fetch_add(1); }
}
inline void mbar_init(uint64_t* bar, uint32_t count) { MbarEmu* m = mbar_emu_of(bar); m->phases = 0; m->tx = 0; m->arrived = 0; m->count = count; }
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    MbarEmu* m = mbar_emu_of(bar);
    m->tx.fetch_add(bytes);
    m->arrived.fetch_add(1);
    mbar_try_complete(m);
}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    // the hardware requires 16-byte aligned addresses and sizes: fail the test when a kernel breaks that
    if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15) || bytes == 0) abort();
    memcpy(dst, src, bytes);
    MbarEmu* m = mbar_emu_of(bar);
    m->tx.fetch_sub(bytes);
    mbar_try_complete(m);
}
inline void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t) { bulk_g2s(dst, src, bytes, bar); }
void mbar_wait(uint64_t* bar, uint32_t parity);  // blocks until the phase of that parity has completed; aborts on a deadlock
