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


def test_field_filter_and_facet_filter_on_a_multifield_index():
    """field_filter (field_filter_set, add_result.rs:3124-3137): a doc is dropped when a query term it contains occurs in none of the filter's
    fields (tested only when term fields + filter fields <= indexed fields); scores still sum every field.  Alone, with facet filters, NOT
    terms and a delete set: ids, scores and counts == the oracle."""
    from seekstorm_b200 import FacetFilter, Index, QueryType, ResultType
    from helpers_facets import abi_filters
    n_docs, vocab, n_fields = 70000, 200, 3
    boosts = (2.0, 1.0, 0.5)
    levels, len_sum = multifield_levels(n_docs, vocab, n_fields, seed=77)
    ix = Index(0); ix.set_field_boosts(boosts)
    orc = O.OracleIndex(); orc.set_fields(boosts)
    for lv in levels:
        ix.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"])
        orc.add_level(lv)
    ix.commit(n_docs, len_sum); orc.commit(n_docs, len_sum)
    rng = np.random.default_rng(78)
    price = rng.integers(0, 1000, n_docs, dtype=np.uint32)
    ix.set_facets({"price": price})
    rows, fields, first, nd, rb = ix._facet_rows
    orc.set_facets(rows, [(fields[0].type, fields[0].offset)], first, nd, rb)
    queries, masks, filters = [], [], []
    for nt in (1, 2, 3, 4, 6):
        for _ in range(8):
            ranks = rng.choice(np.arange(2, vocab), size=nt, replace=False)
            queries.append([int(k) for k in synth.term_keys_np(np.array(ranks, dtype=np.int64))])
            masks.append(int(rng.choice([0, 1, 2, 4, 3, 5, 6, 7])))
            filters.append([FacetFilter("price", 100, 700)] if rng.random() < 0.4 else [])
    nk = [[queries[(i + 5) % len(queries)][0]] if i % 4 == 0 else [] for i in range(len(queries))]
    errs = []
    for deleted in ([], [int(x) for x in rng.integers(0, n_docs, 1500)]):
        ix.set_deleted(deleted); orc.set_deleted(deleted)
        for qt, oqt in ((QueryType.Union, O.QUERY_UNION), (QueryType.Intersection, O.QUERY_INTERSECTION)):
            got, cnt = ix.search_lexical_batch(queries, qt, 10, ResultType.TopkCount, not_keys=nk, filters=filters, field_masks=masks)
            got_t, _ = ix.search_lexical_batch(queries, qt, 10, ResultType.Topk, not_keys=nk, filters=filters, field_masks=masks)
            for i, kq in enumerate(queries):
                tup, sv = abi_filters(ix, filters[i])
                want, tot = orc.search(kq, oqt, 10, O.RESULT_TOPKCOUNT, not_keys=nk[i], filters=tup, set_values=sv, field_mask=masks[i])
                if got[i] != want or got_t[i] != want or int(cnt[i]) != tot:
                    errs.append((bool(deleted), int(qt), i, len(kq), masks[i], bool(filters[i]), got[i][:2], want[:2], int(cnt[i]), tot))
    assert not errs, (len(errs), errs[:5])
    # the filter really drops docs for narrow masks, and the full mask drops nothing
    base, bc = ix.search_lexical_batch(queries, QueryType.Union, 10, ResultType.TopkCount)
    one, oc = ix.search_lexical_batch(queries, QueryType.Union, 10, ResultType.TopkCount, field_masks=[4] * len(queries))
    full, fc = ix.search_lexical_batch(queries, QueryType.Union, 10, ResultType.TopkCount, field_masks=[7] * len(queries))
    assert full == base and list(fc) == list(bc) and sum(int(a) > int(b) for a, b in zip(bc, oc)) > len(queries) // 2
    ix.close()
