"""GPU: facet filters on the scoring path (is_facet_filter add_result.rs:340-478, called at :3498-3500) —
every FieldType, range / set filters, OR / AND, 1..8 terms, exact counts, Topk / TopkCount / Count, combined with NOT terms and the
delete set, paging beyond 32, hybrid, the mirrored Search::search: ids, scores and counts == the oracle."""
import numpy as np
import pytest

from oracle import oracle as O
from seekstorm_b200 import synth
from helpers import gpu_index, oracle_index, query_keys, synth_levels
from helpers_facets import abi_filters, facet_columns, random_filters

pytestmark = pytest.mark.gpu


def _setup(n, vocab, seed, **kw):
    lvs, ls = synth_levels(n, vocab, seed)
    levels = [l.to_numpy() for l in lvs]
    orc = oracle_index(levels, n, ls)
    ix = gpu_index(levels, n, ls, **kw)
    cols, kinds = facet_columns(n, seed + 1)
    ix.set_facets(cols, **kinds)
    rows, fields, first, nd, rb = ix._facet_rows
    orc.set_facets(rows, [(fields[i].type, fields[i].offset) for i in range(len(cols))], first, nd, rb)
    return ix, orc, cols


def test_facet_filters_lexical_parity():
    from seekstorm_b200 import QueryType, ResultType
    n = 150000
    ix, orc, cols = _setup(n, 2500, 71)
    qs = synth.gen_queries(96, 73, 2, 2000, (1, 2, 3, 4, 6, 8), (0.1, 0.3, 0.25, 0.15, 0.1, 0.1))
    qk = query_keys(qs)
    filters = random_filters(cols, 74, len(qk))
    filters[0] = []                                                   # an unfiltered query inside a filtered batch (record path)
    rng = np.random.default_rng(75)
    nots = [[int(x) for x in rng.integers(0, 50, int(rng.integers(0, 2)))] for _ in qs]
    nots = [[t for t in ns if t not in q] for ns, q in zip(nots, qs)]
    nk = [query_keys([ns])[0] if ns else [] for ns in nots]
    errs = []
    for deleted in ([], [int(x) for x in rng.integers(0, n, 3000)]):
        ix.set_deleted(deleted); orc.set_deleted(deleted)
        for qt, oqt in ((QueryType.Union, O.QUERY_UNION), (QueryType.Intersection, O.QUERY_INTERSECTION)):
            got, cnt = ix.search_lexical_batch(qk, qt, 10, ResultType.TopkCount, not_keys=nk, filters=filters)
            got_t, _ = ix.search_lexical_batch(qk, qt, 10, ResultType.Topk, not_keys=nk, filters=filters)
            _, cnt_c = ix.search_lexical_batch(qk, qt, 0, ResultType.Count, not_keys=nk, filters=filters)
            for i, k in enumerate(qk):
                tup, sv = abi_filters(ix, filters[i])
                want, tot = orc.search(k, oqt, 10, O.RESULT_TOPKCOUNT, not_keys=nk[i], filters=tup, set_values=sv) if tup else \
                    orc.search(k, oqt, 10, O.RESULT_TOPKCOUNT, not_keys=nk[i])
                if got[i] != want or got_t[i] != want or int(cnt[i]) != tot or int(cnt_c[i]) != tot:
                    errs.append((bool(deleted), int(qt), i, len(k), filters[i], got[i][:3], want[:3], int(cnt[i]), int(cnt_c[i]), tot))
    assert not errs, (len(errs), errs[:5])
    # how selective the filters were: some queries must lose hits, some must keep some
    ix.set_deleted([]); orc.set_deleted([])
    base, bc = ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount)
    flt, fc = ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount, filters=filters)
    assert sum(int(a) > int(b) > 0 for a, b in zip(bc, fc)) > 20 and base[0] == flt[0]
    ix.close()


def test_facet_filters_paging_hybrid_and_search_mirror():
    from seekstorm_b200 import FacetFilter, QueryType, ResultType, SearchMode, VectorSimilarity
    n, dims = 100000, 32
    ix, orc, cols = _setup(n, 1500, 81, vector_dims=dims, vector_similarity=VectorSimilarity.Cosine)
    rows = synth.gen_vectors(n, dims, 83, "cpu").numpy()
    ix.add_vectors(rows)
    qk = query_keys(synth.gen_queries(12, 84, 2, 1000, (2, 3), (0.5, 0.5)))
    fl = [[FacetFilter("u8", 32, 200), FacetFilter("f32", -5.0, 12.5)] for _ in qk]
    tup, sv = abi_filters(ix, fl[0])
    # paging beyond SSB_K_MAX = 32 hits (internal pages with key ceilings)
    got, cnt = ix.search_lexical_batch(qk, QueryType.Union, 100, ResultType.TopkCount, filters=fl)
    for i, k in enumerate(qk):
        want, tot = orc.search(k, O.QUERY_UNION, 100, O.RESULT_TOPKCOUNT, filters=tup, set_values=sv)
        assert got[i] == want and int(cnt[i]) == tot, (i, k)
    # hybrid: the filter applies to the lexical half only (search_vector_shard takes no facet filter, vector.rs:1105-1115)
    qv = synth.gen_vectors(len(qk), dims, 85, "cpu").numpy()
    nq = len(qk)
    import ctypes as C
    from seekstorm_b200._lib import check, lib
    from seekstorm_b200.index import _hits_array
    b, keep = ix._lex_batch(qk, QueryType.Union, None, fl)
    hits = _hits_array(nq * 10); nh = np.zeros(nq, dtype=np.uint32)
    check(lib().ssb_search_hybrid(ix._h, C.byref(b), qv.ctypes.data, 10, hits.ctypes.data, nh.ctypes.data))
    nrows = np.stack([O.normalize(r) for r in rows])
    for i in range(nq):
        lex, _ = orc.search(qk[i], O.QUERY_UNION, 10, O.RESULT_TOPK, filters=tup, set_values=sv)
        vec = O.search_vector(nrows, O.normalize(qv[i]), 10, O.SIM_COSINE)
        h = hits[i * 10: i * 10 + int(nh[i])]
        assert [int(d) for d in h["doc_id"]] == [d for d, _ in O.rrf(lex, vec)[:10]], i
    # the reference's public call with facet_filter
    ro = ix.search("t40 t300", None, QueryType.Union, SearchMode.Lexical(), False, 0, 10, ResultType.TopkCount, facet_filter=fl[0])
    k2 = query_keys([[40, 300]])[0]
    want, tot = orc.search(k2, O.QUERY_UNION, 10, O.RESULT_TOPKCOUNT, filters=tup, set_values=sv)
    assert [(r.doc_id, np.float32(r.score)) for r in ro.results] == [(d, np.float32(s)) for d, s in want] and ro.result_count_total == tot
    ix.close()


def test_facet_filter_errors_and_coverage():
    from seekstorm_b200 import FacetFilter, QueryType, ResultType, SsbError
    n = 70000
    lvs, ls = synth_levels(n, 800, 91)
    levels = [l.to_numpy() for l in lvs]
    ix = gpu_index(levels, n, ls)
    qk = query_keys([[5, 60]])
    ix._facet_schema = {"x": (0, 2)}
    with pytest.raises(SsbError):                                     # filters without ssb_set_facets
        ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount, filters=[[FacetFilter("x", 0, 5)]])
    # facet rows for the first level only: docs of the second level have no row and fail every filter
    x = np.arange(65536, dtype=np.uint32)
    ix.set_facets({"x": x})
    got, cnt = ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount, filters=[[FacetFilter("x", 0, 2**32 - 1)]])
    full, fc = ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount)
    assert all(d < 65536 for d, _ in got[0]) and 0 < int(cnt[0]) < int(fc[0])
    with pytest.raises(SsbError):                                     # facet index out of range
        ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount, filters=[[FacetFilter(3, 0, 5)]])
    ix.set_facets({})                                                 # cleared
    with pytest.raises(SsbError):
        ix.search_lexical_batch(qk, QueryType.Union, 10, ResultType.TopkCount, filters=[[FacetFilter(0, 0, 5)]])
    ix.close()
