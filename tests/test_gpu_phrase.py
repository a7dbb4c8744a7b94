"""GPU: QueryType::Phrase (add_result.rs:3586-3684) — intersection of the phrase's unique terms + the position check, through
SSB_QUERY_PHRASE on levels loaded with positions: ids, scores and counts == the oracle (whose matches equal a substring search over the
token sequences, tests/test_filters_cpu.py); with a delete set, a facet filter, paging, the mirrored Search::search."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import oracle_index, query_keys
from helpers_phrase import contains_phrase, phrase_queries, sequence_corpus

pytestmark = pytest.mark.gpu


def _gpu_index(levels, n, ls):
    from seekstorm_b200 import Index
    ix = Index(0)
    for lv in levels:
        ix.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"],
                             lv["positions"])
    ix.commit(n, ls)
    return ix


def test_phrase_parity():
    from seekstorm_b200 import FacetFilter, QueryType, ResultType, SearchMode
    n, vocab = 90000, 250
    docs, levels, ls = sequence_corpus(n, vocab, 21)
    orc = oracle_index(levels, n, ls)
    ix = _gpu_index(levels, n, ls)
    phrases = phrase_queries(docs, 22, 160, vocab)
    phrases.append([3, 100000])                                       # a term that is not in the dictionary -> no hit
    phrases.append([7])                                               # one token: a plain term query
    qk = query_keys(phrases)
    qk[-2][1] = 0xDEAD0008
    errs, n_hit = [], 0
    rng = np.random.default_rng(23)
    for deleted in ([], [int(x) for x in rng.integers(0, n, 2500)]):
        ix.set_deleted(deleted); orc.set_deleted(deleted)
        got, cnt = ix.search_lexical_batch(qk, QueryType.Phrase, 10, ResultType.TopkCount)
        got_t, _ = ix.search_lexical_batch(qk, QueryType.Phrase, 10, ResultType.Topk)
        _, cnt_c = ix.search_lexical_batch(qk, QueryType.Phrase, 0, ResultType.Count)
        for i, k in enumerate(qk):
            want, tot = orc.search_phrase(k, 10, O.RESULT_TOPKCOUNT)
            n_hit += tot > 0
            if got[i] != want or got_t[i] != want or int(cnt[i]) != tot or int(cnt_c[i]) != tot:
                errs.append((bool(deleted), i, phrases[i], got[i][:2], want[:2], int(cnt[i]), int(cnt_c[i]), tot))
    assert not errs, (len(errs), errs[:5])
    assert n_hit > 150
    ix.set_deleted([]); orc.set_deleted([])
    # ground truth once more, straight from the token sequences
    i = 0
    got, cnt = ix.search_lexical_batch(qk[:1], QueryType.Phrase, 32, ResultType.TopkCount)
    assert int(cnt[0]) == sum(contains_phrase(d, phrases[0]) for d in docs)
    # paging beyond 32 hits: a frequent bigram
    big = query_keys([[0, 1]])
    got, cnt = ix.search_lexical_batch(big, QueryType.Phrase, 100, ResultType.TopkCount)
    want, tot = orc.search_phrase(big[0], 100, O.RESULT_TOPKCOUNT)
    assert got[0] == want and int(cnt[0]) == tot and tot > 100
    # with a facet filter (both predicates on the same candidate)
    price = rng.integers(0, 100, n, dtype=np.uint32)
    ix.set_facets({"price": price})
    got, cnt = ix.search_lexical_batch(big, QueryType.Phrase, 20, ResultType.TopkCount, filters=[[FacetFilter("price", 10, 40)]])
    allw, _ = orc.search_phrase(big[0], n, O.RESULT_TOPKCOUNT)
    keep = [(d, s) for d, s in allw if 10 <= price[(d >> 16) * 65536 + (d & 0xFFFF)] < 40]
    assert got[0] == keep[:20] and int(cnt[0]) == len(keep)
    # the reference's public call: a quoted query string, and QueryType::Phrase as the default type
    ro = ix.search('"t0 t1"', None, QueryType.Union, SearchMode.Lexical(), False, 0, 10, ResultType.TopkCount)
    assert [(r.doc_id, np.float32(r.score)) for r in ro.results] == [(d, np.float32(s)) for d, s in want[:10]] and ro.result_count_total == tot
    ro2 = ix.search("t0 t1", None, QueryType.Phrase, SearchMode.Lexical(), False, 0, 10, ResultType.TopkCount)
    assert [r.doc_id for r in ro2.results] == [r.doc_id for r in ro.results]
    ix.close()


def test_phrase_needs_positions_and_rejects_mixed_levels():
    from seekstorm_b200 import Index, QueryType, ResultType, SsbError
    docs, levels, ls = sequence_corpus(70000, 100, 31)
    ix = Index(0)
    lv = levels[0]
    ix.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"])
    lv = levels[1]
    with pytest.raises(SsbError):                                      # second level WITH positions after one without
        ix.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"], lv["positions"])
    ix.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"])
    ix.commit(70000, ls)
    with pytest.raises(SsbError):                                      # phrase query without positions
        ix.search_lexical_batch(query_keys([[0, 1]]), QueryType.Phrase, 10, ResultType.TopkCount)
    bad = levels[0]["positions"].copy(); bad[1], bad[0] = bad[0], bad[1]
    ix2 = Index(0)
    lv = levels[0]
    tf0 = int(lv["tfs"][0])
    if tf0 >= 2:
        with pytest.raises(SsbError):                                  # positions of a posting must ascend
            ix2.add_lexical_level(lv["level_id"], lv["n_docs"], lv["term_keys"], lv["posting_offsets"], lv["doc_ids"], lv["tfs"], lv["doc_len_bytes"], bad)
    ix.close(); ix2.close()


def test_phrase_on_the_synthetic_zipf_corpus():
    """the bench's own corpus law (Zipf tokens, 64K-doc levels generated on the GPU) with positions: 2- and 3-token phrases of frequent terms,
    ids / scores / counts == the oracle; lists here are long (bitmap-backed) and span 4 levels."""
    import torch
    from seekstorm_b200 import Index, QueryType, ResultType, synth
    n, vocab = 200000, 50000
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ix = Index(0)
    orc = O.OracleIndex()
    ls = 0
    for lv in synth.gen_lexical_corpus(n, vocab, 41, dev, with_positions=True):
        ix.add_synth_level(lv)
        orc.add_level(lv.to_numpy())
        ls += lv.len_sum_normalized
    ix.commit(n, ls); orc.commit(n, ls)
    rng = np.random.default_rng(42)
    phrases = [[int(x) for x in np.floor(np.exp(rng.uniform(0, np.log(120), int(rng.integers(2, 4)))))] for _ in range(96)]
    qk = query_keys(phrases)
    got, cnt = ix.search_lexical_batch(qk, QueryType.Phrase, 10, ResultType.TopkCount)
    n_hit = 0
    for i, k in enumerate(qk):
        want, tot = orc.search_phrase(k, 10, O.RESULT_TOPKCOUNT)
        assert got[i] == want and int(cnt[i]) == tot, (i, phrases[i], got[i][:2], want[:2], int(cnt[i]), tot)
        n_hit += tot > 0
    assert n_hit > 60
    ix.close()
