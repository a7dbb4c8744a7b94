"""GPU: an index.bin written with real token positions (tests/refwriter.py, the restated writer) loaded through ssb_load_index_bin with
decode_positions answers QueryType::Phrase like the oracle.  (Its own file, sorted last: the newest test of the round.)"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_phrase_on_an_index_loaded_with_positions():
    """index.bin written with real token positions (restated writer) -> ssb_load_index_bin(decode_positions) -> QueryType::Phrase:
    ids / scores / counts == the oracle built from the same token sequences."""
    from seekstorm_b200 import Index, QueryType, ResultType
    from helpers import oracle_index, query_keys
    from helpers_phrase import phrase_queries, sequence_corpus
    import refwriter
    n, vocab = 4000, 80
    docs, lvs, ls = sequence_corpus(n, vocab, 61, docs_per_level=65536, mean_len=20)
    data, len_sum = refwriter.write_index_bin(lvs, n, seed=3)
    assert len_sum == ls
    ix = Index(0)
    assert ix.load_index_bin(data, decode_positions=True) == n
    orc = oracle_index(lvs, n, ls)
    qk = query_keys(phrase_queries(docs, 62, 40, vocab))
    got, cnt = ix.search_lexical_batch(qk, QueryType.Phrase, 10, ResultType.TopkCount)
    hits = 0
    for i, k in enumerate(qk):
        want, tot = orc.search_phrase(k, 10, O.RESULT_TOPKCOUNT)
        assert got[i] == want and int(cnt[i]) == tot, (i, got[i][:2], want[:2], int(cnt[i]), tot)
        hits += tot > 0
    assert hits > 20
    ix.close()
