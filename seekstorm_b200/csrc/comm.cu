// comm.cu — NCCL plumbing of the sharded index (see comm.h).  Reference counterpart: the in-process fan-out over shards and
// the concatenation of their results (search.rs:1637-1743, 1875-1928) — here one process per GPU, one ncclAllGather per batch.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "common.cuh"

namespace ssb {

namespace {
struct Nccl {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
Nccl g_nccl;
std::once_flag g_once;

void load_nccl() {
    const char* env = getenv("SSB_NCCL_LIB");
    void* h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy this process already loaded (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_nccl.h = h;
#define SSB_SYM(name) g_nccl.name = (decltype(g_nccl.name))dlsym(h, "nccl" #name)
    SSB_SYM(GetUniqueId); SSB_SYM(CommInitRank); SSB_SYM(CommDestroy); SSB_SYM(AllGather); SSB_SYM(AllReduce); SSB_SYM(GetErrorString);
#undef SSB_SYM
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.AllReduce && g_nccl.GetErrorString;
}

const Nccl* nccl() {
    std::call_once(g_once, load_nccl);
    if (!g_nccl.ok) { set_error("NCCL is not available (libnccl.so.2 not loadable; set SSB_NCCL_LIB)"); return nullptr; }
    return &g_nccl;
}

#define SSB_NCCL_TRY(expr)                                                                     \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                              \
        if (_r != ncclSuccess) { set_error("NCCL error %d (%s) in %s", (int)_r, n->GetErrorString(_r), #expr); return SSB_E_CUDA; } \
    } while (0)
}  // namespace

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in every NCCL 2.x release");

int32_t comm_unique_id(uint8_t id[128]) {
    const Nccl* n = nccl(); if (!n) return SSB_E_UNSUPPORTED;
    ncclUniqueId u;
    SSB_NCCL_TRY(n->GetUniqueId(&u));
    memcpy(id, &u, 128);
    return SSB_OK;
}

int32_t comm_init(ShardComm& c, const uint8_t id[128], uint32_t rank, uint32_t world) {
    const Nccl* n = nccl(); if (!n) return SSB_E_UNSUPPORTED;
    if (world == 0 || rank >= world) { set_error("comm_init: rank %u / world %u", rank, world); return SSB_E_INVALID; }
    ncclUniqueId u; memcpy(&u, id, 128);
    ncclComm_t comm = nullptr;
    SSB_NCCL_TRY(n->CommInitRank(&comm, (int)world, u, (int)rank));
    c.comm = comm; c.rank = rank; c.world = world; c.owned = true;
    return SSB_OK;
}

void comm_destroy(ShardComm& c) {
    if (c.comm && c.owned && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)c.comm);
    c = ShardComm{};
}

int32_t comm_all_gather_u64(const ShardComm& c, const uint64_t* send, uint64_t* recv, size_t count, cudaStream_t st) {
    const Nccl* n = nccl(); if (!n) return SSB_E_UNSUPPORTED;
    SSB_NCCL_TRY(n->AllGather(send, recv, count, ncclUint64, (ncclComm_t)c.comm, st));
    return SSB_OK;
}
int32_t comm_all_reduce_sum_u64(const ShardComm& c, uint64_t* buf, size_t count, cudaStream_t st) {
    const Nccl* n = nccl(); if (!n) return SSB_E_UNSUPPORTED;
    SSB_NCCL_TRY(n->AllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)c.comm, st));
    return SSB_OK;
}
int32_t comm_all_reduce_max_u64(const ShardComm& c, uint64_t* buf, size_t count, cudaStream_t st) {
    const Nccl* n = nccl(); if (!n) return SSB_E_UNSUPPORTED;
    SSB_NCCL_TRY(n->AllReduce(buf, buf, count, ncclUint64, ncclMax, (ncclComm_t)c.comm, st));
    return SSB_OK;
}

}  // namespace ssb
