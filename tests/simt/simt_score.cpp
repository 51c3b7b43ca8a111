// Host build of the run-length score kernels (metamlst_b200/csrc/score_runs_kernels.cuh, the text the GPU runs) on the
// emulation of simt_host_emul.h: every CTA is 256 std::threads, CTAs run one after another.  TEST INFRASTRUCTURE ONLY.
#define MMLST_HOST_EMUL 1
#include "simt_host_emul.h"

#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <map>
#include <memory>
#include <thread>
#include <vector>

thread_local dim3e threadIdx, blockIdx;
dim3e blockDim, gridDim;
thread_local WarpEmu* t_warp = nullptr;
thread_local BlockEmu* t_block = nullptr;

static std::mutex g_mbar_mu;
static std::map<uint64_t*, std::unique_ptr<MbarEmu>> g_mbars;
MbarEmu* mbar_emu_of(uint64_t* bar) {
    std::lock_guard<std::mutex> lk(g_mbar_mu);
    auto& p = g_mbars[bar];
    if (!p) { p.reset(new MbarEmu()); p->phases = 0; p->tx = 0; p->arrived = 0; p->count = 1; }
    return p.get();
}
void mbar_emu_reset() {
    std::lock_guard<std::mutex> lk(g_mbar_mu);
    g_mbars.clear();
}
void mbar_wait(uint64_t* bar, uint32_t parity) {
    MbarEmu* m = mbar_emu_of(bar);
    const auto t0 = std::chrono::steady_clock::now();
    while ((m->phases.load() & 1u) == parity) {
        std::this_thread::yield();
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) {
            fprintf(stderr, "simt emulation: mbarrier wait never completes (block %u thread %u, parity %u): the kernel would hang\n", blockIdx.x,
                    threadIdx.x, parity);
            abort();
        }
    }
}

namespace {
alignas(128) uint8_t ring_raw[232448];  // the dynamic shared memory of the CTA being run (227 KB)
}
#include "score_runs_kernels.cuh"

template <class K>
static void launch(K kern, unsigned grid, unsigned block, const RunArgs& a) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        mbar_emu_reset();
        Rendezvous rv(static_cast<int>(block));
        BlockEmu be{&rv};
        std::vector<std::unique_ptr<WarpEmu>> warps;
        for (unsigned w = 0; w < block / 32; ++w) warps.emplace_back(new WarpEmu());
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t] {
                threadIdx.x = t; blockIdx.x = b;
                t_warp = warps[t / 32].get(); t_block = &be;
                kern(a);
            });
        for (auto& x : th) x.join();
    }
}

// form: 0 registers, 1 registers + software pipeline, 2..5 shared-memory ring configurations (the product's launcher takes form 0 when a
// file-order index is present; here that combination is refused)
extern "C" int simt_score_runs(int form, unsigned grid, const uint32_t* run_tid, const uint32_t* run_start, uint32_t n_runs, const uint32_t* chunk_run,
                               const int16_t* as0, const uint8_t* xm3, const uint16_t* qlen, const uint16_t* chunk_qlen, const uint32_t* orig_idx,
                               uint64_t n_rec, uint64_t idx_base, const uint8_t* allow, uint32_t n_ref, int minscore, int max_xm, int min_read_len,
                               int64_t* sum_as, uint32_t* n_hit, uint32_t* first_idx, uint64_t* counters) {
    RunArgs a{run_tid, run_start, chunk_run, n_runs, as0, xm3, qlen, orig_idx, chunk_qlen, n_rec, idx_base, allow, n_ref, minscore, max_xm,
              min_read_len, reinterpret_cast<long long*>(sum_as), n_hit, first_idx, reinterpret_cast<unsigned long long*>(counters), 1};
    const bool qc = chunk_qlen != nullptr, oi = orig_idx != nullptr;
    if ((form >= 2 && form <= 5) || (form >= 12 && form <= 15)) {  // 12..15: the builds with the L2 residency hints
        if (oi) return -1;
        const bool hint = form >= 12;
        const int f = hint ? form - 10 : form;
        launch((qc ? (hint ? ring_config<true, true>(f) : ring_config<true, false>(f)) : (hint ? ring_config<false, true>(f) : ring_config<false, false>(f))).kern,
               grid, kThreads, a);
        return 0;
    }
    if (form == 6) {  // pair-fused ring: per-chunk len(SEQ) streams without a file-order index only
        if (oi || !qc) return -1;
        launch(score_runs_ring_pair_kernel, grid, kThreads, a);
        return 0;
    }
    if (form != 0 && form != 1) return -1;
    const bool pipe = form == 1;
    if (!oi && !pipe && !qc) launch(score_runs_kernel<false, false, false>, grid, kThreads, a);
    else if (oi && !pipe && !qc) launch(score_runs_kernel<true, false, false>, grid, kThreads, a);
    else if (!oi && pipe && !qc) launch(score_runs_kernel<false, true, false>, grid, kThreads, a);
    else if (oi && pipe && !qc) launch(score_runs_kernel<true, true, false>, grid, kThreads, a);
    else if (!oi && !pipe && qc) launch(score_runs_kernel<false, false, true>, grid, kThreads, a);
    else if (oi && !pipe && qc) launch(score_runs_kernel<true, false, true>, grid, kThreads, a);
    else if (!oi && pipe && qc) launch(score_runs_kernel<false, true, true>, grid, kThreads, a);
    else launch(score_runs_kernel<true, true, true>, grid, kThreads, a);
    return 0;
}
