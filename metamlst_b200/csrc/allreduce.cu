// The N-GPU exchange of the path behind the C-ABI (north_star: "per-GPU partial pileup-count and score tensors are allreduced
// with NCCL over NVLink"): a caller that is not a torch.distributed process (the reference's scripts are plain Python) joins one NCCL communicator
// per GPU through these four calls.  The reference is single-process and has no counterpart.  libnccl is resolved at run time (dlopen of
// "libnccl.so.2": when PyTorch is already in the process this is the very library it loaded, so there is ONE NCCL per process).
//   tables of a pass : [sum_as i64 | counters u64 x 2 | n_hit u32 (padded to 8 bytes)]   SUM   one all-reduce on 64-bit words (the two u32 halves of
//                      a word never carry into each other below 2^32 hits per allele)
//                      first_idx u32                                                      MIN
//                      counts u32 [columns][5]                                            SUM
// Integer reductions: order-independent, so N ranks give the tables of one rank bit for bit (tests/test_dist_gloo.py runs the same three reductions
// over gloo on the CPU; tests/run_c_allreduce.py runs these calls on two GPUs).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

// the handful of NCCL declarations used here (nccl.h of NCCL 2.x; the ABI of these entry points is stable across 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSumOp = 0, ncclMinOp = 3 };
enum { ncclUint32T = 3, ncclInt64T = 4 };

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl g_nccl;
std::mutex g_nccl_mutex;

int load_nccl() {
    std::lock_guard<std::mutex> g(g_nccl_mutex);
    if (g_nccl.h) return MMLST_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { mmlst_set_error("mmlst_comm: libnccl.so.2 not found (%s)", dlerror()); return MMLST_E_CUDA; }
    Nccl n;
    n.h = h;
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(h, "ncclAllReduce"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(dlsym(h, "ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllReduce || !n.GroupStart || !n.GroupEnd) {
        mmlst_set_error("mmlst_comm: libnccl.so.2 lacks an entry point");
        return MMLST_E_CUDA;
    }
    g_nccl = n;
    return MMLST_OK;
}

int nccl_fail(ncclResult_t r, const char* what) {
    if (r == 0) return MMLST_OK;
    mmlst_set_error("NCCL error in %s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return MMLST_E_CUDA;
}

}  // namespace

struct mmlst_comm { ncclComm_t comm = nullptr; int rank = 0, world = 1, device = 0; };

extern "C" int mmlst_comm_unique_id(uint8_t* id128) {
    if (!id128) { mmlst_set_error("mmlst_comm_unique_id: null pointer"); return MMLST_E_ARG; }
    if (int rc = load_nccl()) return rc;
    ncclUniqueId id;
    if (int rc = nccl_fail(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
    memcpy(id128, id.internal, 128);
    return MMLST_OK;
}

extern "C" int mmlst_comm_create(const uint8_t* id128, int rank, int world, int device, mmlst_comm** out) {
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) { mmlst_set_error("mmlst_comm_create: bad argument"); return MMLST_E_ARG; }
    if (int rc = load_nccl()) return rc;
    CUDA_TRY(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    mmlst_comm* c = new mmlst_comm();
    c->rank = rank; c->world = world; c->device = device;
    if (int rc = nccl_fail(g_nccl.CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank")) { delete c; return rc; }
    *out = c;
    return MMLST_OK;
}

extern "C" void mmlst_comm_destroy(mmlst_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

// see include/mmlst.h
extern "C" int mmlst_allreduce(mmlst_comm* c, int64_t* score_block, size_t score_words, uint32_t* first_idx, size_t n_ref, uint32_t* counts, size_t n_counts,
                               void* stream) {
    if (!c || !c->comm) { mmlst_set_error("mmlst_allreduce: null communicator"); return MMLST_E_ARG; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(c->device));
    if (int rc = nccl_fail(g_nccl.GroupStart(), "ncclGroupStart")) return rc;
    int rc = MMLST_OK;
    if (score_block && score_words) rc = nccl_fail(g_nccl.AllReduce(score_block, score_block, score_words, ncclInt64T, ncclSumOp, c->comm, s), "ncclAllReduce(score tables)");
    if (rc == MMLST_OK && first_idx && n_ref) rc = nccl_fail(g_nccl.AllReduce(first_idx, first_idx, n_ref, ncclUint32T, ncclMinOp, c->comm, s), "ncclAllReduce(first_idx)");
    if (rc == MMLST_OK && counts && n_counts) rc = nccl_fail(g_nccl.AllReduce(counts, counts, n_counts, ncclUint32T, ncclSumOp, c->comm, s), "ncclAllReduce(counts)");
    const int rc2 = nccl_fail(g_nccl.GroupEnd(), "ncclGroupEnd");
    return rc != MMLST_OK ? rc : rc2;
}
