"""CPU: the oracle's facet filter (is_facet_filter, add_result.rs:340-478) against a numpy restatement on typed columns — every
FieldType, NaN / +-inf / +-0.0 / integer extremes, empty ranges, value sets — and the host-side filter encoding."""
import numpy as np

from oracle import oracle as O
from helpers import level_from_postings, oracle_index, key_of
from helpers_facets import facet_columns, numpy_pass, random_filters


class _Enc:
    """the host mirror's filter encoding without a GPU index (Index._encode_filters only needs the schema)"""
    def __init__(self, cols, kinds):
        from seekstorm_b200 import _lib
        from seekstorm_b200.index import Index
        m = {"uint8": _lib.FACET_U8, "uint16": _lib.FACET_U16, "uint32": _lib.FACET_U32, "uint64": _lib.FACET_U64, "int8": _lib.FACET_I8,
             "int16": _lib.FACET_I16, "int32": _lib.FACET_I32, "int64": _lib.FACET_I64, "float32": _lib.FACET_F32, "float64": _lib.FACET_F64}
        self._facet_schema, self.fields, off = {}, [], 0
        for i, (name, a) in enumerate(cols.items()):
            t = m[a.dtype.name]
            if name in kinds["string_facets"]:
                t = _lib.FACET_STRING16 if a.dtype == np.uint16 else _lib.FACET_STRING32
            if name in kinds["timestamp_facets"]:
                t = _lib.FACET_TIMESTAMP
            self._facet_schema[name] = (i, t)
            self.fields.append((t, off))
            off += a.dtype.itemsize
        self.row_bytes = off
        self._encode_filters = Index._encode_filters.__get__(self)

    def rows(self, cols, n):
        rows = np.zeros((n, self.row_bytes), dtype=np.uint8)
        for (t, off), a in zip(self.fields, cols.values()):
            a = np.ascontiguousarray(a)
            rows[:, off:off + a.dtype.itemsize] = a.view(np.uint8).reshape(n, a.dtype.itemsize)
        return rows


def test_oracle_facet_filter_matches_numpy_restatement():
    n = 3000
    rng = np.random.default_rng(5)
    post = {f"t{t}": sorted((int(d), int(rng.integers(1, 5))) for d in rng.choice(n, int(n / (t + 2)), replace=False)) for t in range(8)}
    lens = [O.lib().orc_int_to_byte4(int(x)) for x in rng.integers(5, 200, n)]
    lv = level_from_postings(0, n, post, lens)
    orc = oracle_index([lv], n, int(sum(O.lib().orc_byte4_to_int(b) for b in lens)))
    cols, kinds = facet_columns(n, 6)
    enc = _Enc(cols, kinds)
    orc.set_facets(enc.rows(cols, n), enc.fields, 0, n, enc.row_bytes)
    filters = random_filters(cols, 7, 120)
    n_checked = 0
    for qi, fl in enumerate(filters):
        terms = [f"t{t}" for t in rng.choice(8, int(rng.integers(1, 4)), replace=False)]
        keys = [key_of(t) for t in terms]
        for qt in (O.QUERY_UNION, O.QUERY_INTERSECTION):
            base, _ = orc.search(keys, qt, n, O.RESULT_TOPKCOUNT)                       # every match, best first
            offs, arr, sv = enc._encode_filters([fl])
            tup = [(arr[i].facet, arr[i].kind, arr[i].start, arr[i].end, arr[i].set_first, arr[i].set_count) for i in range(int(offs[1]))]
            got, tot = orc.search(keys, qt, 10, O.RESULT_TOPKCOUNT, filters=tup, set_values=[int(x) for x in sv])
            want = [(d, s) for d, s in base if numpy_pass(cols, fl, d)]
            if not fl:
                got, tot = orc.search(keys, qt, 10, O.RESULT_TOPKCOUNT)
            assert got == want[:10], (qi, terms, fl)
            assert tot == len(want), (qi, terms, fl, tot, len(want))
            n_checked += bool(fl)
    assert n_checked > 100


def test_filter_encoding_of_bounds():
    """signed bounds travel as two's complement, float bounds as f64 bits, unsigned as they are"""
    cols, kinds = facet_columns(10, 1)
    enc = _Enc(cols, kinds)
    from seekstorm_b200 import FacetFilter, _lib
    offs, arr, sv = enc._encode_filters([[FacetFilter("i8", -5, 7), FacetFilter("f32", -1.5, np.inf), FacetFilter("u64", 3, 2**64 - 1),
                                          FacetFilter("s16", values=[4, 9])], []])
    assert list(offs) == [0, 4, 4]
    assert arr[0].start == 2**64 - 5 and arr[0].end == 7 and arr[0].kind == _lib.FILTER_RANGE
    assert arr[1].start == int(np.float64(-1.5).view(np.uint64)) and arr[1].end == int(np.float64(np.inf).view(np.uint64))
    assert arr[2].start == 3 and arr[2].end == 2**64 - 1
    assert arr[3].kind == _lib.FILTER_SET and arr[3].set_first == 0 and arr[3].set_count == 2 and list(sv[:2]) == [4, 9]


def test_oracle_field_filter_on_a_multifield_corpus():
    """field_filter_set (add_result.rs:3124-3137) in the oracle against a direct numpy restatement over the per-field tfs"""
    from helpers_mf import multifield_levels
    from seekstorm_b200 import synth
    n_docs, vocab, nf = 3000, 40, 3
    levels, len_sum = multifield_levels(n_docs, vocab, nf, seed=3)
    orc = O.OracleIndex(); orc.set_fields((1.5, 1.0, 0.5))
    for lv in levels:
        orc.add_level(lv)
    orc.commit(n_docs, len_sum)
    lv = levels[0]
    post = {}                                  # term key -> {doc: mask of fields the term occurs in}
    for ti, key in enumerate(lv["term_keys"]):
        a, b = int(lv["posting_offsets"][ti]), int(lv["posting_offsets"][ti + 1])
        post[int(key)] = {int(d): sum(1 << f for f in range(nf) if lv["tfs"][j, f]) for j, d in zip(range(a, b), lv["doc_ids"][a:b])}
    rng = np.random.default_rng(4)
    keys_all = [int(k) for k in lv["term_keys"]]
    checked = 0
    for _ in range(60):
        keys = [keys_all[i] for i in rng.choice(len(keys_all), int(rng.integers(1, 4)), replace=False)]
        mask = int(rng.integers(1, 8))
        for qt in (O.QUERY_UNION, O.QUERY_INTERSECTION):
            base, _ = orc.search(keys, qt, n_docs, O.RESULT_TOPKCOUNT)
            got, tot = orc.search(keys, qt, n_docs, O.RESULT_TOPKCOUNT, field_mask=mask)
            def ok(d):
                for k in keys:
                    pm = post[k].get(d)
                    if pm is None:
                        continue
                    if bin(pm).count("1") + bin(mask).count("1") <= nf and not (pm & mask):
                        return False
                return True
            want = [(d, s) for d, s in base if ok(d)]
            assert got == want and tot == len(want), (keys, mask, qt)
            checked += len(want) != len(base)
    assert checked > 10


def test_oracle_phrase_search_against_substring_search():
    """orc_search_lexical_phrase (add_result.rs:3586-3684) on a corpus of real token sequences: the matching docs are exactly those whose
    token sequence contains the phrase; scores are the intersection's scores of the unique terms."""
    from helpers_phrase import contains_phrase, phrase_queries, sequence_corpus
    from helpers import query_keys
    n, vocab = 3000, 60
    docs, levels, ls = sequence_corpus(n, vocab, 11, docs_per_level=2000)
    orc = oracle_index(levels, n, ls)
    n_match = 0
    for ph in phrase_queries(docs, 12, 150, vocab):
        keys = query_keys([ph])[0]
        got, tot = orc.search_phrase(keys, n, O.RESULT_TOPKCOUNT)
        want_docs = {d for d in range(n) if contains_phrase(docs[d], ph)}
        # doc id = level << 16 | local, 2000 docs per level
        got_docs = {(d >> 16) * 2000 + (d & 0xFFFF) for d, _ in got}
        assert got_docs == want_docs and tot == len(want_docs), (ph, len(got_docs), len(want_docs))
        uniq = list(dict.fromkeys(keys))
        base, _ = orc.search(uniq, O.QUERY_INTERSECTION, n, O.RESULT_TOPKCOUNT)
        assert got == [(d, s) for d, s in base if ((d >> 16) * 2000 + (d & 0xFFFF)) in want_docs]
        n_match += bool(want_docs)
    assert n_match > 80
