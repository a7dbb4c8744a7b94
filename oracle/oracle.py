"""ctypes binding of the CPU oracle (oracle/libssb_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssb_oracle.so")

QUERY_UNION, QUERY_INTERSECTION, QUERY_PHRASE = 0, 1, 2
RESULT_COUNT, RESULT_TOPK, RESULT_TOPKCOUNT = 0, 1, 2
SIM_DOT, SIM_COSINE, SIM_EUCLIDEAN = 0, 1, 2


class OrcHit(C.Structure):
    _fields_ = [("doc_id", C.c_uint64), ("score", C.c_float), ("pad", C.c_uint32)]


class OrcFacetField(C.Structure):
    _fields_ = [("type", C.c_uint32), ("offset", C.c_uint32)]


class OrcFacetFilter(C.Structure):
    _fields_ = [("facet", C.c_uint32), ("kind", C.c_uint32), ("start", C.c_uint64), ("end", C.c_uint64),
                ("set_first", C.c_uint32), ("set_count", C.c_uint32)]


class OrcLevel(C.Structure):
    _fields_ = [("level_id", C.c_uint32), ("n_docs", C.c_uint32), ("n_terms", C.c_uint32),
                ("reserved", C.c_uint32), ("term_keys", C.c_void_p), ("posting_offsets", C.c_void_p),
                ("doc_ids", C.c_void_p), ("tfs", C.c_void_p), ("doc_len_bytes", C.c_void_p)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ssb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_int_to_byte4.restype = C.c_uint8
        L.orc_int_to_byte4.argtypes = [C.c_uint32]
        L.orc_byte4_to_int.restype = C.c_uint32
        L.orc_byte4_to_int.argtypes = [C.c_uint8]
        L.orc_bm25_cache.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_idf.restype = C.c_float
        L.orc_idf.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_bm25_term.restype = C.c_float
        L.orc_bm25_term.argtypes = [C.c_float, C.c_uint32, C.c_float]
        L.orc_index_new.restype = C.c_void_p
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_index_add_level.argtypes = [C.c_void_p, C.POINTER(OrcLevel)]
        L.orc_index_commit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_index_set_deleted.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_index_df.restype = C.c_uint32
        L.orc_index_df.argtypes = [C.c_void_p, C.c_uint64]
        for f in (L.orc_search_lexical, L.orc_search_lexical_pruned):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                          C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_search_lexical_not.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_index_set_facets.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_search_lexical_filtered.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                                  C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_normalize_f32.argtypes = [C.c_void_p, C.c_uint32]
        for f in (L.orc_dot_f32, L.orc_dot_f32_lanes8, L.orc_euclidean_f32):
            f.restype = C.c_float
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_vector_score_postmap.restype = C.c_float
        L.orc_vector_score_postmap.argtypes = [C.c_float, C.c_uint32]
        L.orc_search_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                        C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_quantize_f32_to_i8.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_quantize_rows_i8.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_dot_i8.restype = C.c_int32
        L.orc_dot_i8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_search_vector_i8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                                           C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_quantize_scale_i8.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_search_vector_i8_scaled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                                  C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_rrf.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                              C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _hits_to_list(buf, n):
    return [(int(buf[i].doc_id), float(np.float32(buf[i].score))) for i in range(n)]


class OracleIndex:
    """Lexical oracle index.  Levels are dicts of numpy arrays as produced by synth.Level.to_numpy()."""

    def __init__(self):
        self._h = lib().orc_index_new()
        self.n_docs = 0
        self.len_sum = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_index_free(self._h)
            self._h = None

    def set_fields(self, boosts):
        """several indexed fields (before the first level): levels then carry tfs [n_postings, n_fields], doc_len_bytes [n_fields, n_docs]"""
        b = np.ascontiguousarray(np.asarray(boosts, dtype=np.float32))
        lib().orc_index_set_fields.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        assert lib().orc_index_set_fields(self._h, len(b), _ptr(b)) == 0

    def add_level(self, lv: dict):
        keep = [np.ascontiguousarray(lv["term_keys"], dtype=np.uint64),
                np.ascontiguousarray(lv["posting_offsets"], dtype=np.uint32),
                np.ascontiguousarray(lv["doc_ids"], dtype=np.uint16),
                np.ascontiguousarray(lv["tfs"], dtype=np.uint16),
                np.ascontiguousarray(lv["doc_len_bytes"], dtype=np.uint8)]
        d = OrcLevel(lv["level_id"], lv["n_docs"], len(keep[0]), 0, *[_ptr(a) for a in keep])
        rc = lib().orc_index_add_level(self._h, C.byref(d))
        assert rc == 0
        if lv.get("positions") is not None:      # term positions (phrase queries)
            pos = np.ascontiguousarray(lv["positions"], dtype=np.uint16)
            lib().orc_index_set_last_level_positions.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
            assert lib().orc_index_set_last_level_positions(self._h, _ptr(pos), pos.size) == 0

    def search_phrase(self, seq_keys, k, result_type):
        """QueryType::Phrase: seq_keys = the phrase's term keys in order (repeats included)"""
        keys = np.ascontiguousarray(np.array(seq_keys, dtype=np.uint64))
        buf = (OrcHit * max(k, 1))()
        n = C.c_uint32(0)
        tot = C.c_uint64(0)
        f = lib().orc_search_lexical_phrase
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        rc = f(self._h, _ptr(keys), len(keys), k, result_type, buf, C.byref(n), C.byref(tot))
        assert rc == 0, rc
        return _hits_to_list(buf, n.value), int(tot.value)

    def commit(self, n_docs: int, len_sum: int):
        self.n_docs, self.len_sum = n_docs, len_sum
        assert lib().orc_index_commit(self._h, n_docs, len_sum) == 0

    def set_deleted(self, doc_ids):
        a = np.ascontiguousarray(np.asarray(list(doc_ids), dtype=np.uint64))
        assert lib().orc_index_set_deleted(self._h, _ptr(a) if a.size else None, a.size) == 0

    def df(self, key: int) -> int:
        return lib().orc_index_df(self._h, C.c_uint64(key))

    def set_facets(self, rows: np.ndarray, fields, first_doc_id: int, n_docs: int, row_bytes: int):
        """the shard's facet file: rows [n_docs, row_bytes] u8, fields = [(type, offset), ...] (same values as the C-ABI)"""
        fa = (OrcFacetField * max(len(fields), 1))(*[OrcFacetField(int(t), int(o)) for t, o in fields])
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        assert lib().orc_index_set_facets(self._h, _ptr(rows), first_doc_id, n_docs, row_bytes, fa, len(fields)) == 0

    def search(self, term_keys, query_type, k, result_type, pruned=False, not_keys=None, filters=None, set_values=None, field_mask=0):
        """filters: [(facet, kind, start_u64, end_u64, set_first, set_count), ...] in the C-ABI's encoding; set_values: the SET filters' ids;
        field_mask: bit f = indexed field f is in the query's field filter (0 = none)"""
        keys = np.ascontiguousarray(np.array(term_keys, dtype=np.uint64))
        buf = (OrcHit * max(k, 1))()
        n = C.c_uint32(0)
        tot = C.c_uint64(0)
        if filters or field_mask:
            filters = filters or []
            nk = np.ascontiguousarray(np.array(not_keys if not_keys else [0], dtype=np.uint64))
            fa = (OrcFacetFilter * max(len(filters), 1))(*[OrcFacetFilter(*[int(x) for x in f]) for f in filters])
            sv = np.ascontiguousarray(np.array(set_values if set_values is not None and len(set_values) else [0], dtype=np.uint64))
            lib().orc_search_lexical_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                                    C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
            rc = lib().orc_search_lexical_ex(self._h, _ptr(keys), len(keys), _ptr(nk), len(not_keys) if not_keys else 0, fa, len(filters), _ptr(sv), int(field_mask),
                                             query_type, k, result_type, buf, C.byref(n), C.byref(tot))
            assert rc == 0, rc
            return _hits_to_list(buf, n.value), int(tot.value)
        if not_keys:
            nk = np.ascontiguousarray(np.array(not_keys, dtype=np.uint64))
            rc = lib().orc_search_lexical_not(self._h, _ptr(keys), len(keys), _ptr(nk), len(nk), query_type, k, result_type, buf, C.byref(n), C.byref(tot))
            assert rc == 0, rc
            return _hits_to_list(buf, n.value), int(tot.value)
        f = lib().orc_search_lexical_pruned if pruned else lib().orc_search_lexical
        rc = f(self._h, _ptr(keys), len(keys), query_type, k, result_type, buf, C.byref(n), C.byref(tot))
        assert rc == 0, rc
        return _hits_to_list(buf, n.value), int(tot.value)


def bm25_cache(n_docs: int, len_sum: int) -> np.ndarray:
    out = np.zeros(256, dtype=np.float32)
    lib().orc_bm25_cache(n_docs, len_sum, _ptr(out))
    return out


def search_vector(rows: np.ndarray, query: np.ndarray, k: int, similarity: int, doc_ids=None,
                  lanes8: bool = False, n_threads: int = 1):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    q = np.ascontiguousarray(query, dtype=np.float32)
    ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint32)
    buf = (OrcHit * max(k, 1))()
    n = C.c_uint32(0)
    rc = lib().orc_search_vector(_ptr(rows), None if ids is None else _ptr(ids), rows.shape[0], rows.shape[1],
                                 rows.strides[0] // 4, _ptr(q), similarity, k, int(lanes8), n_threads,
                                 buf, C.byref(n))
    assert rc == 0
    return _hits_to_list(buf, n.value)


def ivf_premap_threshold(t: float, similarity: int) -> np.float32:
    """TopK::new (vector.rs:388-399): (2t - 1) * 16129 for Dot / Cosine, -t for Euclidean."""
    t = np.float32(t)
    if similarity == SIM_EUCLIDEAN:
        return np.float32(-t)
    return np.float32((t * np.float32(2.0) - np.float32(1.0)) / np.float32(1.0 / 16129.0))


def search_vector_ivf(levels, query: np.ndarray, k: int, similarity: int, ann_mode: int, n_probe: int = 0, cluster_threshold: float = 0.0):
    """search_vector_shard with AnnMode::Nprobe / Similaritythreshold / NprobeSimilaritythreshold (vector.rs:1300-1467).
    levels: list of (level_id, rows [n, d] f32 — already normalised for Cosine, local ids [n] or None, cluster child counts); a cluster is a
    contiguous row range, its medoid its first row (:1316-1320).  Per level: score the query against every medoid, keep the n_probe best
    (score desc, earlier cluster first) that are not below the pre-mapped threshold (:1300-1309, TopK::push :421), scan only their rows.
    ann_mode: 0 All, 1 Nprobe, 2 Similaritythreshold, 3 NprobeSimilaritythreshold.  Returns (hits, observed_vector_count)."""
    q = np.ascontiguousarray(query, dtype=np.float32)
    sel_rows, sel_ids, observed = [], [], 0
    thr = ivf_premap_threshold(cluster_threshold, similarity) if ann_mode in (2, 3) else None
    for level_id, rows, local_ids, counts in levels:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        n = rows.shape[0]
        ids = (np.arange(n, dtype=np.uint32) if local_ids is None else np.asarray(local_ids, dtype=np.uint32)) | np.uint32(level_id << 16)
        counts = [n] if counts is None else [int(c) for c in counts]
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        if ann_mode == 0:
            chosen = list(range(len(counts)))
        else:
            scored = []
            for c, st in enumerate(starts):
                m = np.ascontiguousarray(rows[st])
                s = -lib().orc_euclidean_f32(_ptr(q), _ptr(m), q.size) if similarity == SIM_EUCLIDEAN else lib().orc_dot_f32(_ptr(q), _ptr(m), q.size)
                s = np.float32(s)
                if thr is not None and s < thr:
                    continue
                scored.append((-float(s), c))
            scored.sort()
            np_eff = min(n_probe, len(counts)) if ann_mode in (1, 3) else len(counts)
            chosen = [c for _, c in scored[:np_eff]]
        for c in chosen:
            sel_rows.append(rows[starts[c]: starts[c] + counts[c]]); sel_ids.append(ids[starts[c]: starts[c] + counts[c]])
            observed += counts[c]
    if not sel_rows:
        return [], 0
    return search_vector(np.concatenate(sel_rows), q, k, similarity, doc_ids=np.concatenate(sel_ids)), observed


def normalize(v: np.ndarray) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.float32).copy()
    lib().orc_normalize_f32(_ptr(v), v.size)
    return v


def quantize_i8(v: np.ndarray) -> np.ndarray:
    """normalize_f32 (scalar order) then quantize_f32_to_i8, as the reference does at index / query time for Cosine + SQ-I8."""
    vn = normalize(v)
    out = np.zeros(vn.size, dtype=np.int8)
    lib().orc_quantize_f32_to_i8(_ptr(vn), vn.size, _ptr(out))
    return out


def quantize_rows_i8(rows: np.ndarray) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    out = np.zeros(rows.shape, dtype=np.int8)
    lib().orc_quantize_rows_i8(_ptr(rows), rows.shape[0], rows.shape[1], rows.shape[1], _ptr(out), rows.shape[1])
    return out


def search_vector_i8(rows_i8: np.ndarray, query_i8: np.ndarray, k: int, doc_ids=None):
    rows_i8 = np.ascontiguousarray(rows_i8, dtype=np.int8)
    q = np.ascontiguousarray(query_i8, dtype=np.int8)
    ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint32)
    buf = (OrcHit * max(k, 1))()
    n = C.c_uint32(0)
    rc = lib().orc_search_vector_i8(_ptr(rows_i8), None if ids is None else _ptr(ids), rows_i8.shape[0], rows_i8.shape[1],
                                    rows_i8.strides[0], _ptr(q), k, buf, C.byref(n))
    assert rc == 0
    return _hits_to_list(buf, n.value)


def quantize_scale_rows_i8(rows: np.ndarray, want_norm: bool):
    """QuantizedVector::new_scale / new_scale_norm per row -> (codes int8 [n, d], scale f32 [n], norm f32 [n])."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n, d = rows.shape
    out = np.zeros((n, d), dtype=np.int8); scale = np.zeros(n, dtype=np.float32); norm = np.zeros(n, dtype=np.float32)
    s, nn = C.c_float(0), C.c_float(0)
    for i in range(n):
        lib().orc_quantize_scale_i8(_ptr(rows[i]), d, int(want_norm), _ptr(out[i]), C.byref(s), C.byref(nn))
        scale[i], norm[i] = s.value, nn.value
    return out, scale, norm


def search_vector_i8_scaled(rows_i8, row_scale, row_norm, query_i8, q_scale, q_norm, similarity, k, doc_ids=None):
    rows_i8 = np.ascontiguousarray(rows_i8, dtype=np.int8)
    rs = np.ascontiguousarray(row_scale, dtype=np.float32); rn = np.ascontiguousarray(row_norm, dtype=np.float32)
    q = np.ascontiguousarray(query_i8, dtype=np.int8)
    ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint32)
    buf = (OrcHit * max(k, 1))()
    n = C.c_uint32(0)
    rc = lib().orc_search_vector_i8_scaled(_ptr(rows_i8), _ptr(rs), _ptr(rn), None if ids is None else _ptr(ids), rows_i8.shape[0], rows_i8.shape[1],
                                           rows_i8.strides[0], _ptr(q), C.c_float(q_scale), C.c_float(q_norm), similarity, k, buf, C.byref(n))
    assert rc == 0
    return _hits_to_list(buf, n.value)


def quantize_affine_rows_i8(rows: np.ndarray, state=None, update_state: bool = True):
    """QuantizedVector::new_scale_norm_affine row after row with the shard's running (min, max) state (start: f32::MAX / f32::MIN).
    update_state False = every row sees a COPY of the state (what the query side does, search.rs:1514-1530).
    -> (codes int8 [n, d], scale, norm, zero_point int32, sum_q int32, state (min, max))"""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n, d = rows.shape
    out = np.zeros((n, d), dtype=np.int8); scale = np.zeros(n, dtype=np.float32); norm = np.zeros(n, dtype=np.float32)
    zp = np.zeros(n, dtype=np.int32); sq = np.zeros(n, dtype=np.int32)
    f = lib().orc_quantize_affine_i8
    f.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    fmax = float(np.finfo(np.float32).max)
    smin, smax = (C.c_float(fmax), C.c_float(-fmax)) if state is None else (C.c_float(state[0]), C.c_float(state[1]))
    s, nn, z, su = C.c_float(0), C.c_float(0), C.c_int32(0), C.c_int32(0)
    for i in range(n):
        a, b = (smin, smax) if update_state else (C.c_float(smin.value), C.c_float(smax.value))
        f(_ptr(rows[i]), d, C.byref(a), C.byref(b), _ptr(out[i]), C.byref(s), C.byref(nn), C.byref(z), C.byref(su))
        scale[i], norm[i], zp[i], sq[i] = s.value, nn.value, z.value, su.value
    return out, scale, norm, zp, sq, (smin.value, smax.value)


def search_vector_i8_affine(rows_i8, row_scale, row_norm, row_zp, row_sum, query_i8, q_scale, q_norm, q_zp, q_sum, k, doc_ids=None):
    rows_i8 = np.ascontiguousarray(rows_i8, dtype=np.int8)
    rs = np.ascontiguousarray(row_scale, dtype=np.float32); rn = np.ascontiguousarray(row_norm, dtype=np.float32)
    rz = np.ascontiguousarray(row_zp, dtype=np.int32); ru = np.ascontiguousarray(row_sum, dtype=np.int32)
    q = np.ascontiguousarray(query_i8, dtype=np.int8)
    ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint32)
    buf = (OrcHit * max(k, 1))()
    n = C.c_uint32(0)
    f = lib().orc_search_vector_i8_affine
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                  C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    rc = f(_ptr(rows_i8), _ptr(rs), _ptr(rn), _ptr(rz), _ptr(ru), None if ids is None else _ptr(ids), rows_i8.shape[0], rows_i8.shape[1], rows_i8.strides[0],
           _ptr(q), C.c_float(q_scale), C.c_float(q_norm), int(q_zp), int(q_sum), k, buf, C.byref(n))
    assert rc == 0
    return _hits_to_list(buf, n.value)


def turboquant_rows_i8(rows: np.ndarray, seed_mask: np.ndarray, normalize_first: bool = False):
    """TurboQuant::quantize_f32_i8 per row (Cosine: normalize_f32 first, vector.rs:585-596) -> (codes int8 [n, dim], scale [n], norm [n]); dim = len(seed_mask)"""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    mask = np.ascontiguousarray(seed_mask, dtype=np.float32)
    n, d = rows.shape
    dim = mask.size
    out = np.zeros((n, dim), dtype=np.int8); scale = np.zeros(n, dtype=np.float32); norm = np.zeros(n, dtype=np.float32)
    s, nn = C.c_float(0), C.c_float(0)
    lib().orc_turboquant_i8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for i in range(n):
        r = normalize(rows[i]) if normalize_first else rows[i]
        lib().orc_turboquant_i8(_ptr(r), d, dim, _ptr(mask), _ptr(out[i]), C.byref(s), C.byref(nn))
        scale[i], norm[i] = s.value, nn.value
    return out, scale, norm


def search_vector_i8_turbo(rows_i8, row_scale, row_norm, query_i8, q_scale, q_norm, similarity, k, doc_ids=None):
    rows_i8 = np.ascontiguousarray(rows_i8, dtype=np.int8)
    rs = np.ascontiguousarray(row_scale, dtype=np.float32); rn = np.ascontiguousarray(row_norm, dtype=np.float32)
    q = np.ascontiguousarray(query_i8, dtype=np.int8)
    ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint32)
    buf = (OrcHit * max(k, 1))()
    n = C.c_uint32(0)
    f = lib().orc_search_vector_i8_turbo
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                  C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    rc = f(_ptr(rows_i8), _ptr(rs), _ptr(rn), None if ids is None else _ptr(ids), rows_i8.shape[0], rows_i8.shape[1],
           rows_i8.strides[0], _ptr(q), C.c_float(q_scale), C.c_float(q_norm), similarity, k, buf, C.byref(n))
    assert rc == 0
    return _hits_to_list(buf, n.value)


def rrf(lex, vec):
    a = (OrcHit * max(len(lex), 1))(*[OrcHit(d, s, 0) for d, s in lex])
    b = (OrcHit * max(len(vec), 1))(*[OrcHit(d, s, 0) for d, s in vec])
    out = (OrcHit * max(len(lex) + len(vec), 1))()
    n = C.c_uint32(0)
    lib().orc_rrf(a, len(lex), b, len(vec), out, C.byref(n))
    return _hits_to_list(out, n.value)
