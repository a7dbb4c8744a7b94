"""CPU: the oracle against the golden vectors (tests/golden/golden.json) and the reference's own fixtures."""
import numpy as np
import pytest

from oracle import oracle as O
from seekstorm_b200 import synth
from helpers import key_of, level_from_postings, oracle_index, query_keys, synth_levels


def test_byte4_table(golden):
    L = O.lib()
    tab = golden["byte4_to_int"]
    for b in range(256):
        assert L.orc_byte4_to_int(b) == tab[b] == synth.byte4_to_int(b)
    # index.rs:4232 NUM_FREE_VALUES: the first 24 lengths are exact
    for i in range(24):
        assert L.orc_int_to_byte4(i) == i
    for s, v in golden["int_to_byte4_samples"].items():
        assert L.orc_int_to_byte4(int(s)) == v == synth.int_to_byte4(int(s))
    # round trip is the identity on codes and monotone, never over-estimates
    prev = -1
    for b in range(256):
        v = L.orc_byte4_to_int(b)
        assert L.orc_int_to_byte4(v) == b
        assert v > prev
        prev = v
    for x in (24, 25, 100, 1000, 2000, 65535):
        assert L.orc_byte4_to_int(L.orc_int_to_byte4(x)) <= x


def test_idf_cases(golden):
    for n, df, want in golden["idf_cases"]:
        got = float(np.float32(O.lib().orc_idf(n, df)))
        assert abs(got - want) <= abs(want) * 2e-7, (n, df, got, want)


def _fixture_levels(fx):
    post = {t: [(d, tf) for d, tf in p] for t, p in fx["postings"].items()}
    return [level_from_postings(0, fx["n_docs"], post, fx["len_bytes"])]


def test_reference_fixture_lexical(golden):
    """tests/test.rs:150-208: AND '+body2 +test' -> 1/1/1; Union Count 'test' -> 0 results, total 2."""
    fx = golden["ref_fixture_lexical"]
    ix = oracle_index(_fixture_levels(fx), fx["n_docs"], fx["len_sum"])
    cache = O.bm25_cache(fx["n_docs"], fx["len_sum"])
    for b, v in fx["cache_at_len"].items():
        assert cache[int(b)] == np.float32(v)
    for pruned in (False, True):
        hits, total = ix.search([key_of("body2"), key_of("test")], O.QUERY_INTERSECTION, 10, O.RESULT_TOPKCOUNT, pruned)
        assert len(hits) == 1 and total == 1
        assert hits[0][0] == 2 and hits[0][1] == np.float32(fx["and_body2_test"]["results"][0][1])
        hits, total = ix.search([key_of("test")], O.QUERY_UNION, 10, O.RESULT_COUNT, pruned)
        assert hits == [] and total == 2
        hits, total = ix.search([key_of("body2"), key_of("test")], O.QUERY_UNION, 10, O.RESULT_TOPKCOUNT, pruned)
        assert [(d, np.float32(s)) for d, s in fx["or_body2_test"]["results"]] == [(d, np.float32(s)) for d, s in hits]
        assert total == 2


def test_hand_corpus(golden):
    h = golden["hand_corpus"]
    post = {t: [(d, tf) for d, tf in p] for t, p in h["postings"].items()}
    ix = oracle_index([level_from_postings(0, h["n_docs"], post, h["len_bytes"])], h["n_docs"], h["len_sum"])
    for q in h["queries"]:
        qt = O.QUERY_INTERSECTION if q["type"] == "and" else O.QUERY_UNION
        for pruned in (False, True):
            hits, total = ix.search([key_of(t) for t in q["terms"]], qt, 3, O.RESULT_TOPKCOUNT, pruned)
            assert total == q["count_total"], q
            assert [d for d, _ in hits] == [d for d, _ in q["top3"]], q
            for (_, s), (_, w) in zip(hits, q["top3"]):
                assert abs(s - w) <= 2e-7 * abs(w), q


def test_rrf(golden):
    r = golden["rrf"]
    got = O.rrf([tuple(x) for x in r["lex"]], [tuple(x) for x in r["vec"]])
    assert [d for d, _ in got] == [d for d, _ in r["fused"]]
    for (_, s), (_, w) in zip(got, r["fused"]):
        assert s == np.float32(w)


def test_neon_parity_vectors(golden):
    """vector_similarity.rs:3024-3062: SIMD kernel vs scalar within 1e-3."""
    a = np.array(golden["neon_vec"]["make_f32_128"], dtype=np.float32)
    L = O.lib()
    s = L.orc_dot_f32(a.ctypes.data, a.ctypes.data, 128)
    v = L.orc_dot_f32_lanes8(a.ctypes.data, a.ctypes.data, 128)
    assert s == np.float32(golden["neon_vec"]["dot_self_scalar"])
    assert abs(s - v) < 1e-3
    assert L.orc_euclidean_f32(a.ctypes.data, a.ctypes.data, 128) == 0.0
    n = O.normalize(a)
    assert abs(float(np.dot(n.astype(np.float64), n.astype(np.float64))) - 1.0) < 1e-6


def test_reference_fixture_vector(golden):
    """tests/test.rs:693-745: 3 x 128-d f32 Euclidean, AnnMode::All, length 10 -> 3 results."""
    fx = golden["ref_fixture_vector"]
    rows = np.array([[(128 * j + i + 1) / 1000.0 for i in range(128)] for j in range(3)], dtype=np.float32)
    hits = O.search_vector(rows, rows[0], 10, O.SIM_EUCLIDEAN)
    assert len(hits) == fx["result_count"]
    assert [d for d, _ in hits] == [0, 1, 2]
    for (d, s), (wd, ws) in zip(hits, fx["results"]):
        assert d == wd and abs(s - ws) <= 1e-6 * max(1.0, abs(ws))


@pytest.mark.parametrize("seed", [1, 2])
def test_exhaustive_equals_pruned(seed):
    """The reference-shaped block-max / MAXSCORE control flow returns the exhaustive top-k (canonical ties)."""
    lvs, ls = synth_levels(70000, 3000, seed)
    ix = oracle_index([l.to_numpy() for l in lvs], 70000, ls)
    qs = synth.gen_queries(60, 100 + seed, 2, 2500, (1, 2, 3, 4), (0.1, 0.4, 0.3, 0.2))
    for q, keys in zip(qs, query_keys(qs)):
        for qt in (O.QUERY_UNION, O.QUERY_INTERSECTION):
            for rt in (O.RESULT_TOPK, O.RESULT_TOPKCOUNT):
                a = ix.search(keys, qt, 10, rt)
                b = ix.search(keys, qt, 10, rt, pruned=True)
                assert a[0] == b[0], (q, qt, rt)
                if rt == O.RESULT_TOPKCOUNT:
                    assert a[1] == b[1], (q, qt, rt)


def test_vector_topk_canonical_ties():
    rows = np.zeros((10, 8), dtype=np.float32)
    rows[:, 0] = [1, 2, 2, 3, 3, 3, 0, 5, 5, 1]
    q = np.zeros(8, dtype=np.float32); q[0] = 1
    hits = O.search_vector(rows, q, 4, O.SIM_DOT)
    assert hits == [(7, 5.0), (8, 5.0), (3, 3.0), (4, 3.0)]
    hits_mt = O.search_vector(rows, q, 4, O.SIM_DOT, n_threads=3)
    assert hits_mt == hits


def test_int8_quantisation_known_answers():
    """quantize_f32_to_i8 (vector_similarity.rs:1226-1232): round half AWAY from zero, clamp to [-127, 127]; dot_i8 exact."""
    v = np.array([0.5 / 127, -0.5 / 127, 1.5 / 127, 2.5 / 127, 2.0, -3.0, 0.4999 / 127, 1e-9], dtype=np.float32)
    out = np.zeros(8, dtype=np.int8)
    O.lib().orc_quantize_f32_to_i8(v.ctypes.data, 8, out.ctypes.data)
    assert out.tolist() == [1, -1, 2, 3, 127, -127, 0, 0]
    # unit vector along one axis -> 127 on that axis; [3,4]/5 -> round(76.2), round(101.6)
    assert O.quantize_i8(np.array([0, 0, 9.0, 0], dtype=np.float32)).tolist() == [0, 0, 127, 0]
    assert O.quantize_i8(np.array([3.0, 4.0], dtype=np.float32)).tolist() == [76, 102]
    assert O.quantize_i8(np.zeros(4, dtype=np.float32)).tolist() == [0, 0, 0, 0]          # NaN as i8 = 0
    rng = np.random.default_rng(5)
    a = rng.integers(-127, 128, 300).astype(np.int8)
    b = rng.integers(-127, 128, 300).astype(np.int8)
    assert O.lib().orc_dot_i8(a.ctypes.data, b.ctypes.data, 300) == int(a.astype(np.int64) @ b.astype(np.int64))
    # worst case fits f32 exactly: 768 * 127 * 127 < 2^24
    assert 768 * 127 * 127 < 2 ** 24


def test_int8_search_matches_numpy():
    rng = np.random.default_rng(6)
    rows = rng.normal(size=(500, 96)).astype(np.float32)
    r8 = O.quantize_rows_i8(rows)
    assert (r8 == np.stack([O.quantize_i8(r) for r in rows])).all()
    q8 = O.quantize_i8(rows[17] + 0.1 * rng.normal(size=96).astype(np.float32))
    sc = r8.astype(np.int64) @ q8.astype(np.int64)
    order = np.lexsort((np.arange(500), -sc))[:10]
    got = O.search_vector_i8(r8, q8, 10)
    assert [d for d, _ in got] == order.tolist()
    assert [s for _, s in got] == [float(sc[i]) for i in order]
    assert got[0][0] == 17


def test_int8_reference_generators(golden):
    """The reference's own int8 test vector (make_i8, vector_similarity.rs:3018-3047: dot_i8 of it with itself) and the
    Cosine + ScalarQuantizationI8 codes of its make_f32(128) vector, against an independent numpy/python restatement."""
    g = golden["int8"]
    a = np.array(g["make_i8_128"], dtype=np.int8)
    assert O.lib().orc_dot_i8(a.ctypes.data, a.ctypes.data, 128) == g["dot_i8_self"]
    f = np.array(golden["neon_vec"]["make_f32_128"], dtype=np.float32)
    codes = O.quantize_i8(f)
    assert codes.tolist() == g["quant_of_make_f32_128"]
    assert O.lib().orc_dot_i8(codes.ctypes.data, codes.ctypes.data, 128) == g["dot_codes_self"]
    # a unit vector quantises to a self-product close to 127^2
    assert abs(g["dot_codes_self"] - 127 * 127) < 200


def test_oracle_ivf_probe_semantics():
    """oracle.search_vector_ivf (vector.rs:1300-1467): probing every cluster equals the exhaustive scan; Nprobe(1) scans exactly the cluster
    of the best medoid; a cluster threshold above every medoid score selects nothing; observed = vectors of the selected clusters."""
    from helpers_ivf import clustered_levels
    levels = clustered_levels(16, [(300, 5), (50, 1)], seed=3)
    olevels = [(lid, rows, None, counts) for lid, rows, counts in levels]
    allrows = np.concatenate([lv[1] for lv in levels])
    ids = np.concatenate([np.arange(len(lv[1]), dtype=np.uint32) | np.uint32(lv[0] << 16) for lv in levels])
    q = levels[0][1][120] + 0.05
    full = O.search_vector(allrows, q, 10, O.SIM_DOT, doc_ids=ids)
    got, obs = O.search_vector_ivf(olevels, q, 10, O.SIM_DOT, 1, 1000)
    assert got == full and obs == 350
    assert O.search_vector_ivf(olevels, q, 10, O.SIM_DOT, 0) == (full, 350)
    got1, obs1 = O.search_vector_ivf(olevels, q, 10, O.SIM_DOT, 1, 1)
    counts = [int(c) for c in levels[0][2]]
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    best = int(np.argmax([float(np.float32(levels[0][1][s] @ q)) for s in starts]))
    assert obs1 == counts[best] + 50                       # the best cluster of level 0 + the single cluster of level 1
    lo, hi = (levels[0][0] << 16) + int(starts[best]), (levels[0][0] << 16) + int(starts[best]) + counts[best]
    assert all(lo <= d < hi or (d >> 16) == levels[1][0] for d, _ in got1)
    assert O.search_vector_ivf(olevels, q, 10, O.SIM_DOT, 2, 0, 1.0e6) == ([], 0)
    assert float(O.ivf_premap_threshold(0.5, O.SIM_DOT)) == 0.0 and float(O.ivf_premap_threshold(3.0, O.SIM_EUCLIDEAN)) == -3.0


def test_oracle_multifield_bm25f_known_answer():
    """get_bm25f_multiterm_multifield (add_result.rs:1226-1262) by hand: 2 fields with boosts (2, 1), one level of 3 docs, two terms.
    score(doc) = sum over terms (query order), over the fields the term occurs in (ascending): boost * idf * (tf * 2.2 / (tf + cache[len_byte]))."""
    f32 = np.float32
    orc = O.OracleIndex()
    orc.set_fields([2.0, 1.0])
    lens = np.array([[3, 5, 9], [12, 20, 7]], dtype=np.uint8)                        # byte4 codes < 24 are the lengths themselves
    lv = dict(level_id=0, n_docs=3, term_keys=np.array([8, 16], dtype=np.uint64), posting_offsets=np.array([0, 2, 4], dtype=np.uint32),
              doc_ids=np.array([0, 2, 1, 2], dtype=np.uint16), tfs=np.array([[1, 0], [2, 3], [0, 4], [1, 1]], dtype=np.uint16), doc_len_bytes=lens)
    orc.add_level(lv)
    len_sum = int(lens.sum())
    orc.commit(3, len_sum)
    cache = O.bm25_cache(3, len_sum)
    idf = [f32(O.lib().orc_idf(3, 2))] * 2

    def part(boost, idf_t, tf, lb):
        tf = f32(tf)
        return f32(f32(f32(boost) * idf_t) * f32(f32(tf * f32(2.2)) / f32(tf + cache[lb])))
    want2 = f32(0)
    for x in (part(2, idf[0], 2, 9), part(1, idf[0], 3, 7), part(2, idf[1], 1, 9), part(1, idf[1], 1, 7)):
        want2 = f32(want2 + x)
    got, tot = orc.search([8, 16], O.QUERY_UNION, 10, O.RESULT_TOPKCOUNT)
    assert tot == 3
    assert dict(got)[2] == float(want2)
    assert dict(got)[0] == float(f32(f32(0) + part(2, idf[0], 1, 3)))
    assert dict(got)[1] == float(f32(f32(0) + part(1, idf[1], 4, 20)))
    got_and, tot_and = orc.search([8, 16], O.QUERY_INTERSECTION, 10, O.RESULT_TOPKCOUNT)
    assert tot_and == 1 and got_and == [(2, float(want2))]


def test_round2_golden_turboquant_and_affine(golden):
    """oracle vs the committed known-answer vectors of make_golden.py (numpy-f32 restatements): TurboQuantI8 on an 8-d vector, the affine
    Euclidean quantiser through three vectors of its running state"""
    g = golden["turboquant"]
    c, s, n = O.turboquant_rows_i8(np.array([g["v"]], dtype=np.float32), np.array(g["mask"], dtype=np.float32))
    assert [int(x) for x in c[0]] == g["codes"] and s[0] == np.float32(g["scale"]) and n[0] == np.float32(g["norm"])
    rows = np.array([a["v"] for a in golden["affine_sq"]], dtype=np.float32)
    c, s, n, zp, sq, st = O.quantize_affine_rows_i8(rows)
    for i, a in enumerate(golden["affine_sq"]):
        assert [int(x) for x in c[i]] == a["codes"] and s[i] == np.float32(a["scale"]) and int(zp[i]) == a["zero_point"], (i, a)
        assert int(sq[i]) == a["sum_q"] and n[i] == np.float32(a["norm"])
    assert list(st) == golden["affine_sq"][-1]["state"]


def test_round2_golden_facet_filter_and_phrase(golden):
    from helpers import key_of, level_from_postings, oracle_index
    from seekstorm_b200 import _lib
    types = {"U8": (_lib.FACET_U8, np.uint8), "I8": (_lib.FACET_I8, np.int8), "I32": (_lib.FACET_I32, np.int32), "I64": (_lib.FACET_I64, np.int64),
             "U64": (_lib.FACET_U64, np.uint64), "F32": (_lib.FACET_F32, np.float32), "F64": (_lib.FACET_F64, np.float64),
             "TIMESTAMP": (_lib.FACET_TIMESTAMP, np.int64), "STRING16": (_lib.FACET_STRING16, np.uint16), "STRING32": (_lib.FACET_STRING32, np.uint32)}
    lv = level_from_postings(0, 1, {"t": [(0, 1)]}, [5])
    for case in golden["facet_filter"]:
        t, dt = types[case["type"]]
        orc = oracle_index([lv], 1, 5)
        val = np.array([case["value"]], dtype=dt)
        orc.set_facets(val.view(np.uint8).reshape(1, -1), [(t, 0)], 0, 1, val.dtype.itemsize)
        if "values" in case:
            flt, sv = [(0, 1, 0, 0, 0, len(case["values"]))], case["values"]
        else:
            if case["type"] in ("F32", "F64"):
                enc = lambda x: int(np.float64(x).view(np.uint64))
            elif case["type"] in ("I8", "I32", "I64", "TIMESTAMP"):
                enc = lambda x: int(np.int64(x).view(np.uint64))
            else:
                enc = lambda x: int(np.uint64(x))
            flt, sv = [(0, 0, enc(case["start"]), enc(case["end"]), 0, 0)], None
        hits, tot = orc.search([key_of("t")], O.QUERY_UNION, 10, O.RESULT_TOPKCOUNT, filters=flt, set_values=sv)
        assert (tot == 1) == case["pass"] and (len(hits) == 1) == case["pass"], case
    # phrase containment on token sequences
    docs = golden["phrase"]["docs"]
    post, pos_by = {}, {}
    for d, seq in enumerate(docs):
        for p, tok in enumerate(seq):
            pos_by.setdefault((tok, d), []).append(p)
    for (tok, d), ps in sorted(pos_by.items()):
        post.setdefault(f"w{tok}", []).append((d, len(ps)))
    lvp = level_from_postings(0, len(docs), post, [len(s) for s in docs])
    order = sorted(post.keys())          # level_from_postings lays the terms out in this order
    lvp["positions"] = np.array([p for t in order for d, _ in post[t] for p in pos_by[(int(t[1:]), d)]], dtype=np.uint16)
    orc = oracle_index([lvp], len(docs), sum(len(s) for s in docs))
    for case in golden["phrase"]["cases"]:
        hits, tot = orc.search_phrase([key_of(f"w{t}") for t in case["phrase"]], 10, O.RESULT_TOPKCOUNT)
        assert sorted(d for d, _ in hits) == case["docs"] and tot == len(case["docs"]), case
