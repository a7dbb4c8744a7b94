"""Generates tests/golden/golden.json: known-answer vectors for the oracle and the CUDA path.

Independent restatement in numpy float32 scalars (NOT calling oracle/ or the library) of the reference
formulas, evaluated on tiny hand-checkable inputs:
  * the reference's own fixtures: tests/test.rs:64-86 (4 docs, field `body`), :150-208 (AND '+body2 +test' -> 1/1/1,
    Union Count 'test' -> total 2), :675-745 (3 x 128-d Euclidean, AnnMode::All -> 3 results)
  * byte4 codec table (index.rs:4237-4279), idf edge cases (search.rs:3225-3230), bm25 cache (commit.rs:318-325),
    BM25 scores (add_result.rs:1450-1452), RRF (search.rs:1962-2035), NEON-test vectors
    (vector_similarity.rs:3011-3021 make_f32).
  * round 2: TurboQuantI8 on an 8-d vector (vector_similarity.rs:1929-1958), the affine Euclidean quantiser with its running state
    (:1414-1472), typed facet-filter edge cases (add_result.rs:340-478), phrase containment (:3586-3684).
Run:  python tests/golden/make_golden.py   (writes golden.json next to this file)
"""
import json
import math
import os

import numpy as np

f32 = np.float32


def int_to_byte4(i):
    if i < 24:
        return i
    ii = i - 24
    nb = ii.bit_length()
    if nb < 4:
        return 24 + ii
    sh = nb - 4
    return 24 + (((ii >> sh) & 7) | ((sh + 1) << 3))


def byte4_to_int(b):
    if b < 24:
        return b
    i = b - 24
    bits, sh = i & 7, i >> 3
    return 24 + bits if sh == 0 else 24 + ((bits | 8) << (sh - 1))


def cache(n_docs, len_sum):
    K, B = f32(1.2), f32(0.75)
    avgdl = f32(len_sum) / f32(n_docs)
    return [K * (f32(1.0) - B + B * (f32(byte4_to_int(b)) / avgdl)) for b in range(256)]


def idf(n, df):
    return f32(math.log(float(((f32(n) - f32(df) + f32(0.5)) / (f32(df) + f32(0.5))) + f32(1.0))))  # double log, rounded


def term(idf_, tf, comp):
    tf = f32(tf)
    return idf_ * ((tf * (f32(1.2) + f32(1.0)) / (tf + comp)) + f32(0.0))


def main():
    out = {}
    out["byte4_to_int"] = [byte4_to_int(b) for b in range(256)]
    out["int_to_byte4_samples"] = {str(i): int_to_byte4(i) for i in
                                   [0, 1, 23, 24, 25, 31, 32, 39, 40, 41, 55, 56, 80, 96, 100, 255, 256, 1000, 2000, 65535, 1 << 20]}

    # ---- the reference's 4-doc fixture, field body: "body1", "body1", "body2 test", "body3 test"
    docs = [["body1"], ["body1"], ["body2", "test"], ["body3", "test"]]
    lens = [len(d) for d in docs]
    len_bytes = [int_to_byte4(l) for l in lens]
    len_sum = sum(byte4_to_int(b) for b in len_bytes)
    c = cache(4, len_sum)
    postings = {}
    for di, d in enumerate(docs):
        for t in d:
            postings.setdefault(t, {}).setdefault(di, 0)
            postings[t][di] += 1
    fx = {"docs": docs, "len_bytes": len_bytes, "len_sum": len_sum, "n_docs": 4,
          "postings": {t: sorted(p.items()) for t, p in postings.items()}}
    idf_body2, idf_test = idf(4, 1), idf(4, 2)
    s_and = f32(0.0)
    s_and = s_and + term(idf_body2, 1, c[len_bytes[2]])
    s_and = s_and + term(idf_test, 1, c[len_bytes[2]])
    fx["and_body2_test"] = {"results": [[2, float(s_and)]], "count_total": 1}
    fx["union_count_test"] = {"count_total": 2}
    # OR "body2 test": doc2 has both, doc3 only test
    s3 = f32(0.0) + term(idf_test, 1, c[len_bytes[3]])
    fx["or_body2_test"] = {"results": [[2, float(s_and)], [3, float(s3)]], "count_total": 2}
    fx["idf"] = {"body2": float(idf_body2), "test": float(idf_test)}
    fx["cache_at_len"] = {str(b): float(c[b]) for b in sorted(set(len_bytes))}
    out["ref_fixture_lexical"] = fx

    # ---- idf edge cases
    out["idf_cases"] = [[n, df, float(idf(n, df))] for n, df in [(4, 1), (4, 2), (4, 4), (100000, 1), (100000, 50000), (100000, 100000), (10000000, 123)]]

    # ---- a 6-doc hand corpus with varied tf / lengths
    docs2 = ["a a a b", "a b b b b b b b", "b c", "a c c c c c c c c c c c c c c c c c c c c c c c c c c c c c",
             "c", "a b c a b c a b c"]
    docs2 = [d.split() for d in docs2]
    lb2 = [int_to_byte4(len(d)) for d in docs2]
    ls2 = sum(byte4_to_int(b) for b in lb2)
    c2 = cache(6, ls2)
    post2 = {}
    for di, d in enumerate(docs2):
        for t in d:
            post2.setdefault(t, {}).setdefault(di, 0)
            post2[t][di] += 1
    df2 = {t: len(p) for t, p in post2.items()}
    def score(di, terms):
        s = f32(0.0)
        for t in terms:
            if di in post2[t]:
                s = s + term(idf(6, df2[t]), post2[t][di], c2[lb2[di]])
        return s
    def rank(cands, terms, k):
        sc = sorted(((-float(score(d, terms)), d) for d in cands))
        return [[d, -s] for s, d in sc[:k]]
    hand = {"docs": docs2, "len_bytes": lb2, "len_sum": ls2, "n_docs": 6,
            "postings": {t: sorted(p.items()) for t, p in post2.items()}, "queries": []}
    for terms, qt in [(["a", "b"], "and"), (["a", "b"], "or"), (["a", "b", "c"], "or"), (["c", "a"], "and"), (["b"], "or"), (["a", "zzz"], "and"), (["a", "zzz"], "or")]:
        live = [t for t in terms if t in post2]
        if qt == "and":
            cands = [] if len(live) < len(terms) else [d for d in range(6) if all(d in post2[t] for t in live)]
        else:
            cands = [d for d in range(6) if any(d in post2[t] for t in live)]
        hand["queries"].append({"terms": terms, "type": qt, "top3": rank(cands, live, 3), "count_total": len(cands)})
    out["hand_corpus"] = hand

    # ---- RRF: two 3-item lists, k = 0.6, rank from 0
    lex = [[10, 9.0], [11, 5.0], [12, 1.0]]
    vec = [[12, 0.9], [10, 0.8], [13, 0.7]]
    r = {}
    for i, (d, _) in enumerate(lex):
        r[d] = f32(1.0) / (f32(0.6) + f32(i))
    for i, (d, _) in enumerate(vec):
        s = f32(1.0) / (f32(0.6) + f32(i))
        r[d] = r[d] + s if d in r else s
    fused = sorted(((-float(s), d) for d, s in r.items()))
    out["rrf"] = {"lex": lex, "vec": vec, "fused": [[d, -s] for s, d in fused]}

    # ---- vectors: the reference's NEON-parity generator (vector_similarity.rs:3011-3015)
    def make_f32(n):
        return [float(f32(math.sin(float(f32(i) * f32(0.137)))) * f32(0.5) + f32(math.cos(float(f32(i) * f32(0.013)))) * f32(0.5)) for i in range(n)]
    a = np.array(make_f32(128), dtype=np.float32)
    dot = f32(0.0)
    for x in a:
        dot = dot + x * x
    out["neon_vec"] = {"make_f32_128": [float(x) for x in a], "dot_self_scalar": float(dot), "euclid_self": 0.0}

    # ---- int8: the reference's make_i8 generator and its dot_i8 self-product (vector_similarity.rs:3018-3047), and the
    # Cosine + ScalarQuantizationI8 codes of make_f32(128): normalize_f32 (:70-74, sequential f32 sum) then
    # quantize_f32_to_i8 (:1226-1232, round half away from zero, clamp) — restated here with numpy scalars / python ints
    mi8 = [((i * 17 + 5) % 251) - 125 for i in range(128)]
    assert all(-128 <= x <= 127 for x in mi8)
    norm2 = f32(0.0)
    for x in a:
        norm2 = norm2 + x * x
    fac = f32(1.0) / f32(math.sqrt(float(norm2)))          # f32 sqrt: sqrt in f64 of an f32, rounded once = correctly rounded
    codes = []
    for x in a:
        y = float(f32(f32(x * fac) * f32(127.0)))           # two individually rounded f32 multiplies
        r = math.floor(abs(y) + 0.5) * (1 if y >= 0 else -1)  # exact in f64: half away from zero
        codes.append(int(max(-127, min(127, r))))
    out["int8"] = {"make_i8_128": mi8, "dot_i8_self": sum(x * x for x in mi8),
                   "quant_of_make_f32_128": codes, "dot_codes_self": sum(c * c for c in codes)}

    # ---- the reference's 3-vector Euclidean fixture (tests/test.rs:675-745): v_j[i] = 0.001*(128*j + i + 1)
    vecs = [[(128 * j + i + 1) / 1000.0 for i in range(128)] for j in range(3)]
    q = np.array(vecs[0], dtype=np.float32)
    res = []
    for j in range(3):
        v = np.array(vecs[j], dtype=np.float32)
        s = f32(0.0)
        for x, y in zip(q, v):
            d = x - y
            s = s + d * d
        res.append([j, -float(s)])
    out["ref_fixture_vector"] = {"n": 3, "dims": 128, "results": res, "result_count": 3}

    # ---- round 2: TurboQuantI8 (vector_similarity.rs:1929-1958) on an 8-d vector, every step a float32 scalar operation
    tq_v = [f32(x) for x in (0.5, -1.25, 2.0, 0.125, -0.75, 3.5)]                    # 6 dims -> padded to 8
    tq_mask = [f32(x) for x in (1, -1, -1, 1, 1, -1, 1, -1)]
    a = [tq_v[i] * tq_mask[i] if i < 6 else f32(0.0) * tq_mask[i] for i in range(8)]
    h = 1
    while h < 8:
        for i in range(0, 8, 2 * h):
            for j in range(i, i + h):
                x, y = a[j], a[j + h]
                a[j], a[j + h] = x + y, x - y
        h *= 2
    nrm = f32(math.sqrt(8.0))
    a = [x / nrm for x in a]
    ss = f32(0.0)
    for x in a:
        ss = ss + x * x
    sigma = f32(math.sqrt(float(ss))) / nrm
    tq_scale = max(sigma / f32(32.0), f32(1e-8))
    codes = []
    for x in a:
        r = float(x / tq_scale)
        r = math.floor(abs(r) + 0.5) * (1 if r >= 0 else -1)                         # f32::round: half away from zero
        codes.append(int(max(-127.0, min(127.0, r))))
    sq = sum(c * c for c in codes)
    out["turboquant"] = {"v": [float(x) for x in tq_v], "mask": [float(x) for x in tq_mask], "codes": codes, "scale": float(tq_scale),
                         "norm": float(f32(sq) * tq_scale * tq_scale)}

    # ---- round 2: affine Euclidean SQ (vector_similarity.rs:1414-1472) — three integer vectors through the running (min, max) state
    def raster(r):
        if not (r > 1.0):
            return f32(r)
        v, p = int(r) + 1, 1
        while p < v:
            p <<= 1
        return f32(p - 1)
    smin, smax = f32(np.finfo(np.float32).max), f32(-np.finfo(np.float32).max)
    aff = []
    for vec in ([3, 10, 90, 40], [0, 200, 7, 7], [5, 255, 100, 1]):
        mn, mx = f32(min(vec)), f32(max(vec))
        if mn < smin:
            smin = mn
        else:
            mn = smin
        if mx > smax:
            smax = raster(mx - mn)
        else:
            mx = smax
        scale = raster(mx - mn) / f32(255.0)
        zf = float(f32(-128.0) - mn / scale)
        zp = int(max(-128.0, min(127.0, math.floor(abs(zf) + 0.5) * (1 if zf >= 0 else -1))))
        cd = []
        for x in vec:
            r = float(f32(x) / scale)
            r = math.floor(abs(r) + 0.5) * (1 if r >= 0 else -1)
            cd.append(int(max(-128, min(127, int(r) + zp))))
        norm_i = sum(c * c for c in cd) - 2 * zp * sum(cd) + len(cd) * zp * zp
        aff.append({"v": vec, "scale": float(scale), "zero_point": zp, "codes": cd, "sum_q": sum(cd), "norm": float(f32(norm_i) * scale * scale),
                    "state": [float(smin), float(smax)]})
    out["affine_sq"] = aff

    # ---- round 2: facet filters — Rust Range<T>::contains / Vec::contains on typed values (is_facet_filter, add_result.rs:340-478)
    nan = float("nan")
    out["facet_filter"] = [
        {"type": "U8", "value": 255, "start": 0, "end": 255, "pass": False}, {"type": "U8", "value": 254, "start": 0, "end": 255, "pass": True},
        {"type": "I8", "value": -128, "start": -128, "end": -127, "pass": True}, {"type": "I32", "value": -5, "start": -4, "end": 10, "pass": False},
        {"type": "I64", "value": -(2 ** 63), "start": -(2 ** 63), "end": 0, "pass": True}, {"type": "U64", "value": 2 ** 64 - 1, "start": 0, "end": 2 ** 64 - 1, "pass": False},
        {"type": "F32", "value": -0.0, "start": 0.0, "end": 1.0, "pass": True}, {"type": "F32", "value": nan, "start": -1e30, "end": 1e30, "pass": False},
        {"type": "F64", "value": 1.5, "start": nan, "end": 2.0, "pass": False}, {"type": "F64", "value": float("-inf"), "start": float("-inf"), "end": 0.0, "pass": True},
        {"type": "F32", "value": 3.0, "start": 5.0, "end": 1.0, "pass": False}, {"type": "TIMESTAMP", "value": 1700000000, "start": 1600000000, "end": 1800000000, "pass": True},
        {"type": "STRING16", "value": 7, "values": [1, 7, 9], "pass": True}, {"type": "STRING32", "value": 8, "values": [1, 7, 9], "pass": False},
    ]

    # ---- round 2: phrase (add_result.rs:3586-3684) — docs as token sequences; a doc matches iff the phrase is a contiguous subsequence
    pdocs = [[1, 2, 3, 1, 2], [2, 1, 3], [1, 2, 1, 2, 3], [3, 3, 3], [1, 3, 2]]
    pcases = []
    for ph in ([1, 2], [1, 2, 3], [2, 1], [3, 3], [1, 2, 1, 2], [2, 3, 1, 2], [3, 1]):
        m = len(ph)
        pcases.append({"phrase": ph, "docs": [d for d, seq in enumerate(pdocs) if any(seq[i:i + m] == ph for i in range(len(seq) - m + 1))]})
    out["phrase"] = {"docs": pdocs, "cases": pcases}

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
