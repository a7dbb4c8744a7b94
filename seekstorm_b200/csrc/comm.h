// comm.h — the one exchange step of the sharded index: per-query top-k keys all-gathered over NCCL
// (NVLink / NVSwitch) and merged; match counts all-reduced.  NCCL is resolved at run time (dlopen of the copy already
// loaded in the process — e.g. torch's — or libnccl.so.2 on the loader path, or $SSB_NCCL_LIB): the library has no link-time
// dependency on it and single-GPU users never touch it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssb {

struct ShardComm {
    void* comm = nullptr;      // ncclComm_t
    uint32_t rank = 0, world = 1;
    bool owned = false;        // created by ssb_comm_init (destroyed with the index) vs borrowed through ssb_comm_attach
    bool active() const { return comm != nullptr && world > 1; }
};

int32_t comm_unique_id(uint8_t id[128]);
int32_t comm_init(ShardComm& c, const uint8_t id[128], uint32_t rank, uint32_t world);
void comm_destroy(ShardComm& c);
// send [count] u64 per rank -> recv [world][count]
int32_t comm_all_gather_u64(const ShardComm& c, const uint64_t* send, uint64_t* recv, size_t count, cudaStream_t st);
int32_t comm_all_reduce_sum_u64(const ShardComm& c, uint64_t* buf, size_t count, cudaStream_t st);
int32_t comm_all_reduce_max_u64(const ShardComm& c, uint64_t* buf, size_t count, cudaStream_t st);

}  // namespace ssb
