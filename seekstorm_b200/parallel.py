"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for
the plumbing.

Partition: contiguous ranges of 64K-doc levels per GPU.  Unlike the reference's `doc_id % S`
shards (index.rs:5284), which make idf/avgdl shard-local (search.rs:3225, commit.rs:318-319), every GPU uses the
GLOBAL N, avgdl and per-term df, so scores equal the 1-shard reference exactly.  The only exchange step is the
per-query top-k merge (the reference's in-process `Vec::append` + sort, search.rs:1875-1928, 2097-2106): an
all-gather of k packed (score, doc) keys per query per rank — 256 B/query/rank — followed by a G·k -> k merge.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def level_range(n_levels: int, rank: int, world: int) -> range:
    """Contiguous block range of `rank`: levels [floor(r*L/W), floor((r+1)*L/W))."""
    return range(n_levels * rank // world, n_levels * (rank + 1) // world)


def allreduce_global_df(index, group=None, device=None) -> int:
    """Sum per-term document frequencies over all ranks and install them (idf must use the global df).
    Works with NCCL (device tensors) and gloo (CPU).  Returns the global number of distinct terms."""
    world = dist.get_world_size(group)
    keys, dfs = index.dict_export()
    if world == 1:
        return len(keys)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    n_local = torch.tensor([len(keys)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    n_max = int(max(int(s.item()) for s in sizes))
    pad_k = torch.zeros(max(n_max, 1), dtype=torch.int64, device=dev)
    pad_d = torch.zeros(max(n_max, 1), dtype=torch.int64, device=dev)
    if len(keys):
        pad_k[:len(keys)] = torch.from_numpy(keys.view(np.int64)).to(dev)
        pad_d[:len(keys)] = torch.from_numpy(dfs.astype(np.int64)).to(dev)
    all_k = [torch.zeros_like(pad_k) for _ in range(world)]
    all_d = [torch.zeros_like(pad_d) for _ in range(world)]
    dist.all_gather(all_k, pad_k, group=group)
    dist.all_gather(all_d, pad_d, group=group)
    ks = torch.cat([all_k[r][:int(sizes[r].item())] for r in range(world)])
    ds = torch.cat([all_d[r][:int(sizes[r].item())] for r in range(world)])
    uk, inv = torch.unique(ks, return_inverse=True)
    tot = torch.zeros(uk.numel(), dtype=torch.int64, device=dev).index_add_(0, inv, ds)
    index.set_global_df(uk.cpu().numpy().view(np.uint64), tot.cpu().numpy().astype(np.uint32))
    return int(uk.numel())


def init_shard_comm(index, group=None, device=None):
    """Give the index its own NCCL communicator (behind the C-ABI: ssb_comm_unique_id on rank 0, the 128 id bytes broadcast
    with torch.distributed, ssb_comm_init on every rank).  Afterwards every `index.search_*` call is a collective that returns
    the global result: the all-gather of the per-shard top-k keys, the merge and the count all-reduce are enqueued by the
    library on its search stream — no Python in the exchange."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    ident = index.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(ident.astype(np.uint8)).to(dev)
    dist.broadcast(t, src=0, group=group)
    index.comm_init(t.cpu().numpy(), rank, world)


class ShardedSearcher:
    """Host-side reference of the exchange step with torch.distributed collectives (kept for the gloo CPU tests and as the
    cross-check of the in-library NCCL path; the product path is `init_shard_comm` + plain `index.search_*`).
    Per-rank searcher over a block-range shard; every rank issues the same query batch."""

    def __init__(self, index, group=None, merge_fn=None):
        self.index = index
        # the collectives below are ordered against torch's current stream only: make the index launch on it
        if torch.cuda.is_available() and hasattr(index, "set_stream") and dist.is_initialized() and dist.get_backend(group) == "nccl":
            index.set_stream(torch.cuda.current_stream().cuda_stream)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.merge_fn = merge_fn or (lambda keys_all, n_lists, nq, k: index.merge_keys(keys_all, n_lists, nq, k))
        self._bufs = {}

    def _buf(self, name, shape, device):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.device != torch.device(device):
            t = torch.zeros(shape, dtype=torch.int64, device=device)
            self._bufs[name] = t
        return t

    def gather_keys(self, keys_local: torch.Tensor) -> torch.Tensor:
        """[nq, 32] int64 packed keys on this rank -> [world, nq, 32] on every rank (NCCL all-gather)."""
        if self.world == 1:
            return keys_local.unsqueeze(0)
        out = self._buf("all", (self.world,) + tuple(keys_local.shape), keys_local.device)
        dist.all_gather_into_tensor(out.view(-1, keys_local.shape[-1]), keys_local.contiguous(), group=self.group)
        return out

    def search_vector(self, queries_dev: torch.Tensor, k: int, raw_out=None):
        """raw_out = (hits structured array [nq*k], n_hits u32 [nq]) writes results there instead of building Python lists."""
        nq = int(queries_dev.shape[0])
        keys = self._buf("vec", (nq, 32), queries_dev.device)
        self.index.search_vector_keys(queries_dev, k, keys)
        allk = self.gather_keys(keys)
        if raw_out is not None:
            return self.index.merge_keys_raw(allk, self.world, nq, k, raw_out[0], raw_out[1])
        return self.merge_fn(allk, self.world, nq, k)

    def search_lexical(self, batch_struct, nq: int, k: int, result_type, device="cuda", raw_out=None):
        keys = self._buf("lex", (nq, 32), device)
        counts = self._buf("cnt", (nq,), device)
        self.index.search_lexical_keys(batch_struct, k, result_type, keys, counts)
        allk = self.gather_keys(keys)
        if self.world > 1:
            dist.all_reduce(counts, group=self.group)   # result_count_total = Σ shards (search.rs:1875-1940)
        if raw_out is not None:
            return self.index.merge_keys_raw(allk, self.world, nq, k, raw_out[0], raw_out[1]), counts
        return self.merge_fn(allk, self.world, nq, k), counts
