"""CPU: the N>1 host logic with world_size-2 gloo — block-range sharding, global-df all-reduce, packed-key
all-gather + merge.  The GPU merge kernel is replaced by a numpy merge here (test code only)."""
import os
import struct

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seekstorm_b200.parallel import ShardedSearcher, allreduce_global_df, init_shard_comm, level_range


def test_level_range_partitions():
    for n_levels in (1, 2, 15, 16, 153, 1526):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += list(level_range(n_levels, r, world))
            assert seen == list(range(n_levels))
            sizes = [len(level_range(n_levels, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _ord(f):
    u = struct.unpack("<I", struct.pack("<f", f))[0]
    return (~u & 0xFFFFFFFF) if u & 0x80000000 else (u | 0x80000000)


def _pack(score, doc):
    k = (_ord(score) << 32) | (0xFFFFFFFF - doc)
    return k - (1 << 64) if k >= (1 << 63) else k


def _np_merge(keys_all, n_lists, nq, k):
    """numpy stand-in for ssb_merge_keys: top-k of the union of the per-rank descending lists (unsigned order)."""
    a = keys_all.cpu().numpy().view(np.uint64).reshape(n_lists, nq, 32)
    out = []
    for q in range(nq):
        allk = np.sort(a[:, q, :].reshape(-1))[::-1]
        allk = allk[allk != 0][:k]
        out.append([int(x) for x in allk])
    return out


class FakeIndex:
    """Stands in for seekstorm_b200.Index (which needs a GPU): each rank owns half of a tiny scored corpus."""

    def __init__(self, rank, world):
        rng = np.random.default_rng(123)
        self.scores = rng.normal(size=(4, 200)).astype(np.float32)   # 4 queries x 200 docs
        self.mine = [d for d in range(200) if level_range(200, rank, world).start <= d < level_range(200, rank, world).stop]
        self.keys = np.array([1000 + rank, 5, 7 + 10 * rank], dtype=np.uint64)
        self.dfs = np.array([3, 10 + rank, 2], dtype=np.uint32)
        self.global_df = None

    # the C-ABI communicator bootstrap (ssb_comm_unique_id / ssb_comm_init) seen from the host side
    def comm_unique_id(self):
        return (np.arange(128) * 7 % 251).astype(np.uint8)

    def comm_init(self, ident, rank, world):
        self.comm = (bytes(ident.tobytes()), rank, world)

    def dict_export(self):
        o = np.argsort(self.keys)
        return self.keys[o], self.dfs[o]

    def set_global_df(self, keys, dfs):
        self.global_df = dict(zip([int(k) for k in keys], [int(d) for d in dfs]))

    def search_vector_keys(self, queries, k, keys_out):
        for q in range(keys_out.shape[0]):
            ks = sorted((_pack(float(self.scores[q, d]), d) for d in self.mine), key=lambda x: x % (1 << 64), reverse=True)[:k]
            row = ks + [0] * (32 - len(ks))
            keys_out[q] = torch.tensor(row, dtype=torch.int64)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ix = FakeIndex(rank, world)
        n = allreduce_global_df(ix, device="cpu")
        assert n == 5
        assert ix.global_df == {5: 21, 7: 2, 17: 2, 1000: 3, 1001: 3}
        assert ix.global_df[5] == 10 + 11 and ix.global_df[1000] == 3 and ix.global_df[1001] == 3
        init_shard_comm(ix, device="cpu")      # rank 0's id bytes reach every rank, each calls comm_init(rank, world)
        assert ix.comm == (bytes(((np.arange(128) * 7) % 251).astype(np.uint8).tobytes()), rank, world)
        sh = ShardedSearcher(ix, merge_fn=_np_merge)
        got = sh.search_vector(torch.zeros((4, 8)), 10)
        full = FakeIndex(0, 1)
        for q in range(4):
            want = sorted((_pack(float(full.scores[q, d]), d) % (1 << 64) for d in range(200)), reverse=True)[:10]
            assert got[q] == want
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allgather_merge_and_df():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 1, 1: 1}
