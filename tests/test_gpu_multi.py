"""GPU (needs >= 2 GPUs, otherwise skipped): block-range sharding over NCCL returns exactly the 1-GPU results."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_DOCS, VOCAB, DIMS, N_VEC = 200000, 20000, 64, 150000


def _build(rank, world, dev):
    from seekstorm_b200 import Index, VectorSimilarity, synth
    from seekstorm_b200.parallel import init_shard_comm, level_range
    ix = Index(dev, vector_dims=DIMS, vector_similarity=VectorSimilarity.Cosine)
    n_levels = (N_DOCS + 65535) // 65536
    ls = 0
    for lv in synth.gen_lexical_corpus(N_DOCS, VOCAB, 5, "cpu"):
        ls += lv.len_sum_normalized                       # global statistic (every rank sees every level's stats)
        if lv.level_id in level_range(n_levels, rank, world):
            ix.add_synth_level(lv)
    ix.commit(N_DOCS, ls)
    if world > 1:
        init_shard_comm(ix)        # NCCL communicator owned by the library (ssb_comm_init)
        ix.sync_df()               # ssb_lexical_sync_df: index-wide df on every shard
    rows = synth.gen_vectors(N_VEC, DIMS, 6, "cpu").numpy()
    nvl = (N_VEC + 65535) // 65536
    for l in level_range(nvl, rank, world):
        ix.add_vector_level(l, rows[l * 65536: min(N_VEC, (l + 1) * 65536)])
    return ix


def _queries():
    from seekstorm_b200 import synth
    qs = synth.gen_queries(64, 7, 5, 15000, (2, 3), (0.5, 0.5))
    qk = [[int(k) for k in synth.term_keys_np(np.array(q, dtype=np.int64))] for q in qs]
    qv = synth.gen_vectors(64, DIMS, 8, "cpu").numpy()
    return qk, qv


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from seekstorm_b200 import QueryType, ResultType
        ix = _build(rank, world, rank)
        qk, qv = _queries()
        # plain C-ABI calls with host buffers: the exchange (ncclAllGather + merge, count all-reduce, RRF after the merge) is the library's
        vec = ix.search_vector_batch(qv, 10)
        lex, counts = ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount)
        land, acounts = ix.search_lexical_batch(qk, QueryType.Intersection, 10, ResultType.TopkCount)
        hyb = ix.search_hybrid_batch(qk, QueryType.Union, qv, 10)
        deep = ix.search_vector_batch(qv[:4], 50)          # paging beyond 32 across shards
        ret[f"vec{rank}"] = vec
        if rank == 0:
            ret["vec"] = vec
            ret["lex"] = lex
            ret["counts"] = [int(c) for c in counts]
            ret["and"] = land
            ret["acounts"] = [int(c) for c in acounts]
            ret["hyb"] = hyb
            ret["deep"] = deep
        ix.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharding_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from seekstorm_b200 import QueryType, ResultType
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 300, ret), nprocs=world, join=True)
    ix = _build(0, 1, 0)
    qk, qv = _queries()
    want_vec = ix.search_vector_batch(qv, 10)
    want_lex, want_counts = ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount)
    want_and, want_acounts = ix.search_lexical_batch(qk, QueryType.Intersection, 10, ResultType.TopkCount)
    assert ret["lex"] == want_lex                      # bit-exact: global N / avgdl / df on every shard
    assert ret["counts"] == [int(c) for c in want_counts]
    assert ret["and"] == want_and and ret["acounts"] == [int(c) for c in want_acounts]
    assert ret["vec0"] == ret["vec1"]                  # every rank returns the global result
    want_hyb = ix.search_hybrid_batch(qk, QueryType.Union, qv, 10)
    assert [[d for d, _ in h] for h in ret["hyb"]] == [[d for d, _ in h] for h in want_hyb]
    want_deep = ix.search_vector_batch(qv[:4], 50)
    assert [[d for d, _ in h] for h in ret["deep"]] == [[d for d, _ in h] for h in want_deep]
    for g, w in zip(ret["vec"], want_vec):
        assert [d for d, _ in g] == [d for d, _ in w]
        assert np.allclose([s for _, s in g], [s for _, s in w], rtol=1e-5)
    ix.close()
