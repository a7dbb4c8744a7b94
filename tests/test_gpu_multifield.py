"""GPU: BM25F over several indexed fields (get_bm25f_multiterm_multifield, add_result.rs:1171-1426) through ssb_lexical_set_field_boosts +
per-field tfs / doc lengths, vs the oracle: ids, ranks, scores and counts `==` (every f32 operation individually rounded on both sides,
accumulation order = query order x ascending field)."""
import numpy as np
import pytest

from oracle import oracle as O
from seekstorm_b200 import synth
from helpers_mf import multifield_levels

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_fields,boosts", [(2, (2.0, 1.0)), (3, (3.0, 1.0, 0.5)), (4, (1.0, 1.0, 1.0, 1.0))])
def test_multifield_bm25f_parity(n_fields, boosts):
    from seekstorm_b200 import Index, QueryType, ResultType
    n_docs, vocab = 70000, 300
    levels, len_sum = multifield_levels(n_docs, vocab, n_fields, seed=40 + n_fields)
    ix = Index(0)
    ix.set_field_boosts(boosts)
    orc = O.OracleIndex()
    orc.set_fields(boosts)
    for lv in levels:
        ix.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"])
        orc.add_level(lv)
    ix.commit(n_docs, len_sum)
    orc.commit(n_docs, len_sum)
    rng = np.random.default_rng(9)
    queries = []
    for nt in (1, 2, 2, 3, 4, 6, 9):
        for _ in range(6):
            ranks = rng.choice(np.arange(2, vocab), size=nt, replace=False)
            queries.append([int(k) for k in synth.term_keys_np(np.array(ranks, dtype=np.int64))])
    queries.append([int(synth.term_keys_np(np.array([5], dtype=np.int64))[0]), 0x1234567 << 3])     # a term that is not in the dictionary
    for qt, oqt in ((QueryType.Union, O.QUERY_UNION), (QueryType.Intersection, O.QUERY_INTERSECTION)):
        for rt, ort in ((ResultType.TopkCount, O.RESULT_TOPKCOUNT), (ResultType.Topk, O.RESULT_TOPK)):
            got, counts = ix.search_lexical_batch(queries, qt, 10, rt)
            for i, kq in enumerate(queries):
                want, tot = orc.search(kq, oqt, 10, ort)
                assert got[i] == want, (n_fields, qt, rt, i, got[i][:3], want[:3])
                if rt == ResultType.TopkCount:
                    assert int(counts[i]) == tot
    # NOT terms and the delete set go through the same general kernel
    nk = [[queries[3][0]]] + [None] * (len(queries) - 1)
    got, counts = ix.search_lexical_batch(queries, QueryType.Union, 10, ResultType.TopkCount, [n or [] for n in nk])
    want, tot = orc.search(queries[0], O.QUERY_UNION, 10, O.RESULT_TOPKCOUNT, not_keys=nk[0])
    assert got[0] == want and int(counts[0]) == tot
    ix.close()


def test_multifield_contract_errors():
    from seekstorm_b200 import Index
    levels, len_sum = multifield_levels(500, 20, 2, seed=1)
    lv = levels[0]
    ix = Index(0)
    with pytest.raises(Exception):       # level with 2 fields into a single-field index
        ix._n_fields = 2
        ix.add_lexical_level(0, lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"])
    ix._n_fields = 1
    with pytest.raises(Exception):
        ix.set_field_boosts([1.0] * 5)
    ix.set_field_boosts([1.0, 2.0])
    bad = lv["tfs"].copy(); bad[3] = 0          # a posting whose term occurs in no field
    with pytest.raises(Exception):
        ix.add_lexical_level(0, lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], bad, lv["doc_len_bytes"])
    ix.add_lexical_level(0, lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"])
    with pytest.raises(Exception):       # too late
        ix.set_field_boosts([1.0, 1.0])
    ix.close()
