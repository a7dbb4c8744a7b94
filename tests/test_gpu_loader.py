"""GPU: ssb_load_index_bin / ssb_load_vector_bin — an index loaded from the reference's file format answers exactly like the
same corpus loaded through the neutral layout, and like the oracle.  Fixtures come from tests/refwriter.py (a restatement of the
reference's writer: no Rust toolchain here, so loader parity is unpinned by the reference itself)."""
import numpy as np
import pytest

from oracle import oracle as O
from seekstorm_b200 import synth
from helpers import gpu_index, oracle_index
import refwriter
from test_loader_cpu import _levels_with_special_lists

pytestmark = pytest.mark.gpu


def test_load_index_bin_matches_neutral_load_and_oracle():
    from seekstorm_b200 import Index, QueryType, ResultType
    lvs, n_docs = _levels_with_special_lists()
    data, len_sum = refwriter.write_index_bin(lvs, n_docs, seed=5)
    ix_file = Index(0)
    assert ix_file.load_index_bin(data) == n_docs
    ix_neutral = gpu_index(lvs, n_docs, len_sum)
    orc = oracle_index(lvs, n_docs, len_sum)
    rng = np.random.default_rng(3)
    special = [0x1000 << 3, 0x2000 << 3, 0x3000 << 3, 0x4000 << 3]
    keys0 = [int(k) for k in lvs[0]["term_keys"][:400]]
    queries = [[special[0], special[1]], [special[2]], [special[1], special[3]], [special[3], special[2], special[0]]]
    for _ in range(60):
        nt = int(rng.integers(1, 5))
        queries.append([keys0[int(i)] for i in rng.choice(len(keys0), nt, replace=False)] + ([special[int(rng.integers(0, 4))]] if rng.random() < 0.3 else []))
    for qt, oqt in ((QueryType.Union, O.QUERY_UNION), (QueryType.Intersection, O.QUERY_INTERSECTION)):
        a, ca = ix_file.search_lexical_batch(queries, qt, 10, ResultType.TopkCount)
        b, cb = ix_neutral.search_lexical_batch(queries, qt, 10, ResultType.TopkCount)
        assert a == b and [int(x) for x in ca] == [int(x) for x in cb]
        for i, q in enumerate(queries):
            want, tot = orc.search(q, oqt, 10, O.RESULT_TOPKCOUNT)
            assert a[i] == want and int(ca[i]) == tot, (i, q)
    ix_file.close(); ix_neutral.close()


def test_load_vector_bin_matches_add_vectors():
    from seekstorm_b200 import Index, VectorSimilarity
    dims = 48
    rows = synth.gen_vectors(70000, dims, 77, "cpu").numpy()
    levels = []
    for s in range(0, 70000, 65536):
        n = min(65536, 70000 - s)
        levels.append((np.arange(n, dtype=np.uint16), rows[s:s + n]))
    # a second vector (chunk) for a few docs of level 1: doc ids repeat -> best chunk per doc
    extra_ids = np.array([5, 6, 7], dtype=np.uint16)
    extra_rows = rows[65536 + 5: 65536 + 8] * 0.5 + rows[10:13] * 0.5
    levels[1] = (np.concatenate([levels[1][0], extra_ids]), np.concatenate([levels[1][1], extra_rows]))
    data = refwriter.write_vector_bin(levels)
    ix = Index(0, vector_dims=dims, vector_similarity=VectorSimilarity.Cosine)
    assert ix.load_vector_bin(data) == 70003
    ref = Index(0, vector_dims=dims, vector_similarity=VectorSimilarity.Cosine)
    ref.add_vector_level(0, levels[0][1]); ref.add_vector_level(1, levels[1][1], levels[1][0])
    qs = synth.gen_vectors(20, dims, 78, "cpu").numpy()
    qs[0] = rows[65536 + 6]                      # a doc with two chunks: must come back once, with the better chunk's score
    a, b = ix.search_vector_batch(qs, 10), ref.search_vector_batch(qs, 10)
    assert a == b
    assert a[0][0][0] == (1 << 16) | 6 and len({d for d, _ in a[0]}) == len(a[0])
    ix.close(); ref.close()

