"""Host-side mirror of the reference's search interface for the hot path, on top of the C-ABI.

Mirrors (names, argument meaning, error behaviour) of /root/reference/seekstorm/src/:
  * `Search::search`                        search.rs:1134-1150 (impl 1153-2131)  -> Index.search
  * `QueryType`, `ResultType`, `SearchMode`, `AnnMode`   search.rs:120-185, vector_similarity.rs:43-67
  * `ResultObject` / `Result`               search.rs:186-213, min_heap.rs:17-40
  * per-shard seams search_lexical_shard (search.rs:2427-2458) / search_vector_shard (vector.rs:1105-1115)
    -> Index.search_lexical_batch / Index.search_vector_batch (batched: the GPU path amortises launches)

The reference's `search` is infallible by type: failures yield an empty `ResultObject` (search.rs:1630-1631).
The same holds here for "term not in dictionary"/empty queries; programming errors (k > 32, no GPU,
library missing) raise `SsbError` — there is no CPU fallback.

The tokenizer (tokenizer.rs) is out of scope: query strings are split on whitespace, a leading '+' marks a
mandatory term (tokenizer.rs:546-563); terms are mapped to 64-bit keys by `term_key_fn`.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import SsbConfig, SsbHit, SsbLevelDesc, SsbLexBatch, SsbStats, check, lib


class QueryType(enum.IntEnum):
    """search.rs `QueryType`.  Phrase: the query's terms in phrase order, repeats included; needs levels loaded with positions."""
    Union = 0
    Intersection = 1
    Phrase = 2


class ResultType(enum.IntEnum):
    """search.rs:150-175."""
    Count = 0
    Topk = 1
    TopkCount = 2


class VectorSimilarity(enum.IntEnum):
    """vector_similarity.rs:20-30."""
    Dot = 0
    Cosine = 1
    Euclidean = 2


@dataclass(frozen=True)
class AnnMode:
    """vector_similarity.rs:43-67: which IVF clusters of each level are searched (vector.rs:1300-1392)."""
    kind: int = 0                 # SSB_ANN_*: 0 All, 1 Nprobe, 2 Similaritythreshold, 3 NprobeSimilaritythreshold
    n_probe: int = 0
    threshold: float = 0.0

    @staticmethod
    def Nprobe(n):
        return AnnMode(1, int(n))

    @staticmethod
    def Similaritythreshold(t):
        return AnnMode(2, 0, float(t))

    @staticmethod
    def NprobeSimilaritythreshold(n, t):
        return AnnMode(3, int(n), float(t))


AnnMode.All = AnnMode()


@dataclass
class SearchMode:
    """search.rs `SearchMode::{Lexical, Vector{..}, Hybrid{..}}`."""
    kind: str = "Lexical"
    similarity_threshold: Optional[float] = None
    ann_mode: AnnMode = AnnMode.All

    @staticmethod
    def Lexical():
        return SearchMode("Lexical")

    @staticmethod
    def Vector(similarity_threshold=None, ann_mode=AnnMode.All):
        return SearchMode("Vector", similarity_threshold, ann_mode)

    @staticmethod
    def Hybrid(similarity_threshold=None, ann_mode=AnnMode.All):
        return SearchMode("Hybrid", similarity_threshold, ann_mode)


@dataclass(frozen=True)
class FacetFilter:
    """`FacetFilter` (search.rs:735-860): a range filter `start <= value < end` (Rust `Range<T>`) on a numeric / timestamp facet field, or a
    value-id set on a String16 / String32 facet (values = the ids the reference resolves the filter strings to, FilterSparse::String16/32).
    field: the facet's name (Index.set_facets) or its index."""
    field: object
    start: object = None
    end: object = None
    values: Optional[Sequence[int]] = None


@dataclass
class Result:
    """min_heap.rs:17-40."""
    doc_id: int
    score: float


@dataclass
class ResultObject:
    """search.rs:186-213 (facets / suggestions are outside the hot path)."""
    original_query: str = ""
    query: str = ""
    query_terms: list = field(default_factory=list)
    result_count: int = 0
    result_count_total: int = 0
    results: list = field(default_factory=list)
    observed_vector_count: int = 0
    observed_cluster_count: int = 0


def fnv1a64(term: str) -> int:
    h = 0xCBF29CE484222325
    for b in term.encode("utf-8"):
        h = ((h ^ b) * 0x100000001B3) & ((1 << 64) - 1)
    return h & ~7


def synthetic_term_key(term: str) -> int:
    """Key of a synthetic term 't<id>' (synth.py) or FNV-1a of any other string; low 3 bits clear
    (reserved for the n-gram type in the reference, index.rs:4165-4225)."""
    from .synth import splitmix64
    if len(term) > 1 and term[0] == "t" and term[1:].isdigit():
        return splitmix64(int(term[1:])) & ~7
    return fnv1a64(term)


def _addr(x):
    """Address of a numpy array / torch tensor (host or device) or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch.Tensor


def _hits_array(n):
    return np.zeros(n, dtype=np.dtype([("doc_id", "<u8"), ("score", "<f4"), ("pad", "<u4")]))


class Index:
    """One shard's GPU-resident mirror: committed lexical levels + vector levels on one B200."""

    def __init__(self, device: int = 0, vector_dims: int = 0,
                 vector_similarity: VectorSimilarity = VectorSimilarity.Cosine, max_batch: int = 4096,
                 term_key_fn: Callable[[str], int] = synthetic_term_key, vector_kernel: int = 0,
                 vector_quantization: int = 0):
        """vector_kernel: SSB_VEC_KERNEL_* (0 = auto, 1 = FP32 FFMA scan, 2/3 = tcgen05 3xTF32, 4/5/6 = tcgen05 3xBF16 with 128/64/256
        queries per pass, 7/8 = bf16 filter scan + exact f32 refine with 128/256 queries per pass)."""
        self._h = C.c_void_p()
        cfg = SsbConfig(device, max_batch, vector_dims, int(vector_similarity), vector_kernel, int(vector_quantization),
                        (C.c_uint32 * 2)(0, 0))
        check(lib().ssb_create(C.byref(cfg), C.byref(self._h)))
        self.vector_dims = vector_dims
        self.vector_similarity = VectorSimilarity(vector_similarity)
        self.term_key_fn = term_key_fn
        self.indexed_doc_count = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().ssb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ load
    def set_field_boosts(self, boosts):
        """Several indexed fields (BM25F, add_result.rs:1171-1426): per-field boosts, before the first level.  Levels then carry
        tfs [n_postings, n_fields] (0 = term not in that field) and doc_len_bytes [n_fields, n_docs]."""
        b = np.ascontiguousarray(np.asarray(boosts, dtype=np.float32))
        check(lib().ssb_lexical_set_field_boosts(self._h, len(b), b.ctypes.data))
        self._n_fields = len(b)
        if len(getattr(self, "field_names", [])) != len(b):
            self.field_names = [f"field{f}" for f in range(len(b))]      # names of the indexed fields in schema order (Index.search field_filter)

    def add_lexical_level(self, level_id: int, n_docs: int, term_keys, posting_offsets, doc_ids, tfs, doc_len_bytes, positions=None):
        """One committed 64K-doc level in the neutral layout (arrays: numpy on host or torch on the device).  positions: u16 [sum of tfs], the
        term positions of every posting in posting order (phrase queries), or None."""
        n_terms = int(term_keys.shape[0])
        d = SsbLevelDesc(level_id, n_docs, n_terms, getattr(self, "_n_fields", 1), _addr(term_keys), _addr(posting_offsets), _addr(doc_ids),
                         _addr(tfs), _addr(doc_len_bytes), _addr(positions))
        check(lib().ssb_lexical_add_level(self._h, C.byref(d)))

    def add_synth_level(self, lv):
        """Convenience: a seekstorm_b200.synth.Level (tensors on CPU or on this index's device)."""
        if lv.term_keys.is_cuda:
            self.add_lexical_level(lv.level_id, lv.n_docs, lv.term_keys, lv.posting_offsets, lv.doc_ids, lv.tfs,
                                   lv.doc_len_bytes, getattr(lv, "positions", None))
        else:
            n = lv.to_numpy()
            self.add_lexical_level(n["level_id"], n["n_docs"], n["term_keys"], n["posting_offsets"], n["doc_ids"],
                                   n["tfs"], n["doc_len_bytes"], n.get("positions"))

    def load_index_bin(self, data, indexed_field_count: int = 1, key_head_size: int = 20, segment_number_bits: int = 11, decode_positions: bool = False) -> int:
        """Load one shard's index.bin (bytes / mmap / numpy uint8 array, the reference's own format, index.rs:3253-3516) and commit.
        Returns indexed_doc_count."""
        from ._lib import SsbIndexBinParams
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        prm = SsbIndexBinParams(indexed_field_count, key_head_size, segment_number_bits, 1 if decode_positions else 0)   # positions: phrase queries
        n = C.c_uint64(0)
        check(lib().ssb_load_index_bin(self._h, buf.ctypes.data, buf.size, C.byref(prm), C.byref(n)))
        self.indexed_doc_count = n.value
        return n.value

    def load_vector_bin(self, data) -> int:
        """Load one shard's vector.bin (vector.rs:1066-1094, f32 records).  Returns the number of vectors."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        n = C.c_uint64(0)
        check(lib().ssb_load_vector_bin(self._h, buf.ctypes.data, buf.size, C.byref(n)))
        return n.value

    def commit(self, n_docs: int, len_sum_normalized: int):
        """Global statistics of the whole shard (commit.rs:318-319) + directory / block-max build."""
        check(lib().ssb_lexical_commit(self._h, n_docs, len_sum_normalized))
        self.indexed_doc_count = n_docs

    def dict_export(self):
        n = C.c_uint64(0)
        check(lib().ssb_lexical_dict_size(self._h, C.byref(n)))
        keys = np.zeros(n.value, dtype=np.uint64)
        dfs = np.zeros(n.value, dtype=np.uint32)
        check(lib().ssb_lexical_dict_export(self._h, keys.ctypes.data, dfs.ctypes.data, n.value))
        return keys, dfs

    def set_global_df(self, keys: np.ndarray, dfs: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        dfs = np.ascontiguousarray(dfs, dtype=np.uint32)
        check(lib().ssb_lexical_set_global_df(self._h, keys.ctypes.data, dfs.ctypes.data, len(keys)))

    # ------------------------------------------------------------------ sharded index (one process per GPU)
    def comm_unique_id(self) -> np.ndarray:
        """ncclGetUniqueId through the library (rank 0); ship the 128 bytes to the other ranks and call comm_init everywhere."""
        ident = np.zeros(128, dtype=np.uint8)
        check(lib().ssb_comm_unique_id(ident.ctypes.data))
        return ident

    def comm_init(self, ident: np.ndarray, rank: int, world: int):
        """Collective.  From here on every search_* call returns the GLOBAL result on every rank: the library all-gathers the
        per-shard top-k keys over NCCL and merges them (all-reduces the counts) on the search stream."""
        ident = np.ascontiguousarray(ident, dtype=np.uint8)
        assert ident.size == 128
        check(lib().ssb_comm_init(self._h, ident.ctypes.data, rank, world))

    def comm_destroy(self):
        check(lib().ssb_comm_destroy(self._h))

    def sync_df(self):
        """Collective: install the index-wide document frequencies (idf of the whole index on every shard)."""
        check(lib().ssb_lexical_sync_df(self._h))

    def set_deleted(self, doc_ids):
        """shard.delete_hashset: these docs are neither scored nor counted (lexical, vector, hybrid); [] clears the set."""
        a = np.ascontiguousarray(np.asarray(list(doc_ids), dtype=np.uint64))
        check(lib().ssb_set_deleted(self._h, a.ctypes.data if a.size else None, a.size))

    def set_facets(self, columns: dict, first_doc_id: int = 0, string_facets: Sequence[str] = (), timestamp_facets: Sequence[str] = ()):
        """The shard's facet file (`facets_file_mmap`, add_result.rs:343-347): one typed value per doc and facet field.  columns: name ->
        numpy array [n_docs] (dtype = the facet's FieldType; names in string_facets are String16 / String32 value ids, names in
        timestamp_facets Timestamp); rows are packed field after field like the reference's facet file and handed to ssb_set_facets."""
        from ._lib import SsbFacetField
        kinds = {"uint8": _lib.FACET_U8, "uint16": _lib.FACET_U16, "uint32": _lib.FACET_U32, "uint64": _lib.FACET_U64, "int8": _lib.FACET_I8,
                 "int16": _lib.FACET_I16, "int32": _lib.FACET_I32, "int64": _lib.FACET_I64, "float32": _lib.FACET_F32, "float64": _lib.FACET_F64}
        names = list(columns)
        n = len(next(iter(columns.values()))) if names else 0
        fields, off, self._facet_schema = (SsbFacetField * max(len(names), 1))(), 0, {}
        for i, name in enumerate(names):
            a = np.asarray(columns[name])
            t = kinds[a.dtype.name]
            if name in string_facets:
                t = {"uint16": _lib.FACET_STRING16, "uint32": _lib.FACET_STRING32}[a.dtype.name]
            if name in timestamp_facets:
                t = {"int64": _lib.FACET_TIMESTAMP}[a.dtype.name]
            fields[i] = SsbFacetField(t, off)
            self._facet_schema[name] = (i, t)
            off += a.dtype.itemsize
        rows = np.zeros((max(n, 1), max(off, 1)), dtype=np.uint8)
        for i, name in enumerate(names):
            a = np.ascontiguousarray(columns[name])
            rows[:n, fields[i].offset:fields[i].offset + a.dtype.itemsize] = a.view(np.uint8).reshape(n, a.dtype.itemsize)
        self._facet_rows = (rows, fields, int(first_doc_id), n, off)      # also what the tests hand to the oracle
        check(lib().ssb_set_facets(self._h, rows.ctypes.data, int(first_doc_id), n, off, fields, len(names)))

    def _encode_filters(self, filters):
        """filters: per query a list of FacetFilter -> (filter_offsets, ssb_facet_filter array, set values)"""
        from ._lib import SsbFacetFilter
        offs = np.zeros(len(filters) + 1, dtype=np.uint32)
        flat, sets = [], []
        for i, fl in enumerate(filters):
            for f in fl or ():
                idx, t = self._facet_schema[f.field] if isinstance(f.field, str) else (int(f.field), None)
                if t is None:
                    t = next((v[1] for v in getattr(self, "_facet_schema", {}).values() if v[0] == idx), _lib.FACET_U64)
                if f.values is not None:
                    flat.append(SsbFacetFilter(idx, _lib.FILTER_SET, 0, 0, len(sets), len(f.values)))
                    sets.extend(int(v) for v in f.values)
                else:
                    if t in (_lib.FACET_F32, _lib.FACET_F64):
                        enc = lambda x: int(np.float64(x).view(np.uint64))
                    elif t in (_lib.FACET_I8, _lib.FACET_I16, _lib.FACET_I32, _lib.FACET_I64, _lib.FACET_TIMESTAMP):
                        enc = lambda x: int(np.int64(x).view(np.uint64))
                    else:
                        enc = lambda x: int(np.uint64(x))
                    flat.append(SsbFacetFilter(idx, _lib.FILTER_RANGE, enc(f.start), enc(f.end), 0, 0))
            offs[i + 1] = len(flat)
        arr = (SsbFacetFilter * max(len(flat), 1))(*flat)
        sv = np.asarray(sets if sets else [0], dtype=np.uint64)
        return offs, arr, sv

    def add_vector_level(self, level_id: int, rows, local_ids=None, cluster_counts=None):
        """rows: [n, dims] f32 (numpy or torch, host or device), n <= 65536.  cluster_counts: the level's IVF cluster table (rows in
        cluster order, medoid = first row of each cluster; vector.rs:1066-1094) or None = one cluster."""
        n, dims = int(rows.shape[0]), int(rows.shape[1])
        stride = rows.strides[0] // 4 if isinstance(rows, np.ndarray) else rows.stride(0)
        if cluster_counts is None:
            check(lib().ssb_vector_add_level(self._h, level_id, _addr(rows), stride, _addr(local_ids), n, dims))
        else:
            cc = np.ascontiguousarray(cluster_counts, dtype=np.uint32)
            check(lib().ssb_vector_add_level_clustered(self._h, level_id, _addr(rows), stride, _addr(local_ids), n, dims, cc.ctypes.data, len(cc)))

    def set_turboquant_mask(self, seed_mask):
        """TurboQuantI8 indexes (vector_quantization = 2): the index's +-1 sign mask (TurboQuant.seed_mask, next_power_of_two(dims) values)"""
        m = np.ascontiguousarray(seed_mask, dtype=np.float32)
        check(lib().ssb_vector_set_turboquant_mask(self._h, m.ctypes.data, m.size))

    def reserve_vectors(self, n_rows: int):
        """Capacity hint (ssb_vector_reserve): one allocation for n_rows rows instead of geometric growth while loading."""
        check(lib().ssb_vector_reserve(self._h, int(n_rows)))

    def add_vectors(self, rows, first_level: int = 0):
        """Split a big [N, dims] matrix into 64K-row levels (doc_id = row index when first_level = 0)."""
        n = int(rows.shape[0])
        self.reserve_vectors(self.vector_count + n)
        for s in range(0, n, 65536):
            self.add_vector_level(first_level + s // 65536, rows[s:min(n, s + 65536)])

    def set_vector_kernel(self, kernel: int):
        """0 = auto, 1 = FP32 FFMA2 scan, 2/3 = tcgen05 3xTF32 (128/64 queries per pass), 4/5/6 = tcgen05 3xBF16 (128/64/256)."""
        check(lib().ssb_set_vector_kernel(self._h, kernel))

    @property
    def vector_count(self) -> int:
        n = C.c_uint64(0)
        check(lib().ssb_vector_count(self._h, C.byref(n)))
        return n.value

    # ------------------------------------------------------------------ batched shard-level search
    def _lex_batch(self, queries_keys: Sequence[Sequence[int]], query_type: QueryType, not_keys: Optional[Sequence[Sequence[int]]] = None,
                   filters: Optional[Sequence[Sequence["FacetFilter"]]] = None, field_masks: Optional[Sequence[int]] = None):
        """not_keys: per query the keys of its '-' terms (not_query_list, add_result.rs:3440-3496) or None.
        filters: per query its FacetFilter list (facet_filter of search_lexical_shard) or None."""
        nots = not_keys if not_keys is not None else [[] for _ in queries_keys]
        offs = np.zeros(len(queries_keys) + 1, dtype=np.uint32)
        for i, q in enumerate(queries_keys):
            if len(q) > _lib.MAX_QUERY_TERMS:
                raise _lib.SsbError(f"query {i} has {len(q)} terms (> {_lib.MAX_QUERY_TERMS})")
            offs[i + 1] = offs[i] + len(q) + len(nots[i])
        keys = np.zeros(max(int(offs[-1]), 1), dtype=np.uint64)
        flags = np.zeros(max(int(offs[-1]), 1), dtype=np.uint8)
        p = 0
        for q, nq_ in zip(queries_keys, nots):
            for t in q:
                keys[p] = t
                p += 1
            for t in nq_:
                keys[p] = t
                flags[p] = 1
                p += 1
        b = SsbLexBatch(len(queries_keys), int(query_type), offs.ctypes.data, keys.ctypes.data,
                        flags.ctypes.data if not_keys is not None else None, None, None, None, None)
        keep = [offs, keys, flags]
        if filters is not None:
            foffs, farr, fsets = self._encode_filters(filters)
            b.filter_offsets, b.filters, b.filter_set_values = foffs.ctypes.data, C.addressof(farr), fsets.ctypes.data
            keep += [foffs, farr, fsets]
        if field_masks is not None:       # field_filter: per query a bitmask of indexed fields (0 = none)
            fm = np.ascontiguousarray(np.asarray(list(field_masks), dtype=np.uint32))
            b.field_masks = fm.ctypes.data
            keep.append(fm)
        return b, tuple(keep)

    def search_lexical_batch(self, queries_keys, query_type: QueryType, k: int,
                             result_type: ResultType = ResultType.TopkCount, not_keys=None, filters=None, field_masks=None):
        """Batched search_lexical_shard.  Returns (list of [(doc_id, score)...], counts ndarray).  not_keys: '-' terms per query;
        filters: FacetFilter list per query (needs set_facets)."""
        nq = len(queries_keys)
        b, keep = self._lex_batch(queries_keys, query_type, not_keys, filters, field_masks)
        hits = _hits_array(max(nq * max(k, 1), 1))
        n_hits = np.zeros(max(nq, 1), dtype=np.uint32)
        counts = np.zeros(max(nq, 1), dtype=np.uint64)
        check(lib().ssb_search_lexical(self._h, C.byref(b), k, int(result_type), hits.ctypes.data, n_hits.ctypes.data,
                                       counts.ctypes.data))
        out = []
        for i in range(nq):
            h = hits[i * k: i * k + int(n_hits[i])]
            out.append([(int(d), float(s)) for d, s in zip(h["doc_id"], h["score"])])
        return out, counts[:nq]

    def search_vector_batch(self, queries, k: int):
        """Batched search_vector_shard (AnnMode::All).  queries: [nq, dims] f32 numpy/torch."""
        nq = int(queries.shape[0])
        if isinstance(queries, np.ndarray):
            queries = np.ascontiguousarray(queries, dtype=np.float32)
        hits = _hits_array(max(nq * k, 1))
        n_hits = np.zeros(max(nq, 1), dtype=np.uint32)
        check(lib().ssb_search_vector(self._h, _addr(queries), nq, k, hits.ctypes.data, n_hits.ctypes.data))
        out = []
        for i in range(nq):
            h = hits[i * k: i * k + int(n_hits[i])]
            out.append([(int(d), float(s)) for d, s in zip(h["doc_id"], h["score"])])
        return out

    # ---- raw C-ABI calls with caller-owned numpy buffers (no per-hit Python objects; what a Rust shim would do) ----
    def make_lex_batch(self, queries_keys, query_type: QueryType):
        """Build the host-side ssb_lex_batch once; returns (struct, keepalive)."""
        return self._lex_batch(queries_keys, query_type)

    def search_lexical_raw(self, batch_struct, k: int, result_type: ResultType, hits, n_hits, counts):
        """ssb_search_lexical with preallocated outputs: hits = structured array [nq*k] (doc_id u8, score f4, pad u4)."""
        check(lib().ssb_search_lexical(self._h, C.byref(batch_struct), k, int(result_type), hits.ctypes.data,
                                       n_hits.ctypes.data, counts.ctypes.data if counts is not None else None))

    def search_vector_raw(self, queries, k: int, hits, n_hits):
        """ssb_search_vector with preallocated outputs; queries: host or device [nq, dims] f32."""
        check(lib().ssb_search_vector(self._h, _addr(queries), int(queries.shape[0]), k, hits.ctypes.data, n_hits.ctypes.data))

    @staticmethod
    def hits_buffer(n):
        return _hits_array(n)

    def search_vector_ex(self, queries, k: int, similarity_threshold=None, int8_queries: bool = False, ann_mode: int = 0, n_probe: int = 0,
                         cluster_threshold: float = 0.0):
        """ssb_search_vector_ex: threshold (vector.rs:388-399), int8 query codes, vb result fields, observed_vector_count.
        Returns (hits per query, ext structured array [nq, k], observed [nq])."""
        from ._lib import SsbHitExt, SsbVecQuery
        nq = int(queries.shape[0])
        if isinstance(queries, np.ndarray):
            queries = np.ascontiguousarray(queries, dtype=np.int8 if int8_queries else np.float32)
        hits = _hits_array(max(nq * k, 1))
        n_hits = np.zeros(max(nq, 1), dtype=np.uint32)
        ext = (SsbHitExt * max(nq * k, 1))()
        observed = np.zeros(max(nq, 1), dtype=np.uint64)
        vq = SsbVecQuery(_addr(queries), nq, k, 1 if int8_queries else 0, 0 if similarity_threshold is None else 1,
                         0.0 if similarity_threshold is None else float(similarity_threshold), int(ann_mode), int(n_probe), float(cluster_threshold))
        check(lib().ssb_search_vector_ex(self._h, C.byref(vq), hits.ctypes.data, n_hits.ctypes.data, C.addressof(ext), observed.ctypes.data))
        out = []
        for i in range(nq):
            h = hits[i * k: i * k + int(n_hits[i])]
            out.append([(int(d), float(s)) for d, s in zip(h["doc_id"], h["score"])])
        return out, ext, observed[:nq]

    def search_hybrid_batch(self, queries_keys, query_type: QueryType, queries, k: int):
        nq = len(queries_keys)
        b, keep = self._lex_batch(queries_keys, query_type)
        if isinstance(queries, np.ndarray):
            queries = np.ascontiguousarray(queries, dtype=np.float32)
        hits = _hits_array(max(nq * k, 1))
        n_hits = np.zeros(max(nq, 1), dtype=np.uint32)
        check(lib().ssb_search_hybrid(self._h, C.byref(b), _addr(queries), k, hits.ctypes.data, n_hits.ctypes.data))
        out = []
        for i in range(nq):
            h = hits[i * k: i * k + int(n_hits[i])]
            out.append([(int(d), float(s)) for d, s in zip(h["doc_id"], h["score"])])
        return out

    def last_stats(self) -> dict:
        s = SsbStats()
        check(lib().ssb_last_stats(self._h, C.byref(s)))
        return dict(kernel_launches=s.kernel_launches, algorithmic_bytes=s.algorithmic_bytes, h2d_bytes=s.h2d_bytes,
                    d2h_bytes=s.d2h_bytes, postings_visited=s.postings_visited, probes=s.probes,
                    items_processed=s.items_processed, items_skipped=s.items_skipped,
                    dominant_kernel_ns=s.dominant_kernel_ns, scan_bytes_read=s.scan_bytes_read,
                    filter_fallbacks=s.filter_fallbacks)

    # ------------------------------------------------------------------ device-resident API (bench / multi-GPU)
    def search_vector_keys(self, queries_dev, k: int, keys_out_dev):
        check(lib().ssb_search_vector_keys(self._h, _addr(queries_dev), int(queries_dev.shape[0]), k, _addr(keys_out_dev)))

    def search_lexical_keys(self, batch_struct, k: int, result_type: ResultType, keys_out_dev, counts_dev=None):
        check(lib().ssb_search_lexical_keys(self._h, C.byref(batch_struct), k, int(result_type), _addr(keys_out_dev),
                                            _addr(counts_dev)))

    def merge_keys(self, keys_dev, n_lists: int, nq: int, k: int):
        hits = _hits_array(max(nq * k, 1))
        n_hits = np.zeros(max(nq, 1), dtype=np.uint32)
        check(lib().ssb_merge_keys(self._h, _addr(keys_dev), n_lists, nq, k, hits.ctypes.data, n_hits.ctypes.data))
        return [[(int(d), float(s)) for d, s in zip(hits[i * k: i * k + int(n_hits[i])]["doc_id"],
                                                     hits[i * k: i * k + int(n_hits[i])]["score"])] for i in range(nq)]

    def merge_keys_raw(self, keys_dev, n_lists: int, nq: int, k: int, hits, n_hits):
        """ssb_merge_keys into caller-owned numpy buffers (no per-hit Python objects)."""
        check(lib().ssb_merge_keys(self._h, _addr(keys_dev), n_lists, nq, k, hits.ctypes.data, n_hits.ctypes.data))

    def sync(self):
        check(lib().ssb_sync(self._h))

    @property
    def stream(self) -> int:
        return lib().ssb_stream(self._h) or 0

    def set_stream(self, cuda_stream):
        """Run on a caller-owned CUDA stream handle (e.g. torch.cuda.current_stream().cuda_stream; 0 = the CUDA legacy
        default stream); None restores the index's own stream."""
        check(lib().ssb_set_stream(self._h, C.c_void_p(-1 if cuda_stream is None else cuda_stream)))

    # ------------------------------------------------------------------ the reference's public call
    def search(self, query_string: str, query_vector=None, query_type_default: QueryType = QueryType.Union,
               search_mode: SearchMode = None, enable_empty_query: bool = False, offset: int = 0, length: int = 10,
               result_type: ResultType = ResultType.TopkCount, include_uncommitted: bool = False,
               field_filter: Sequence[str] = (), query_facets: Sequence = (), facet_filter: Sequence = (),
               result_sort: Sequence = (), query_rewriting=None) -> ResultObject:
        """`Search::search` (search.rs:1134-1150) for committed data, 1-shard semantics.

        facet_filter: FacetFilter objects (range / value-set filters on the facet fields given to set_facets) — applied to the lexical
        search like the reference does (the vector search takes no facet filter, vector.rs:1105-1115).
        Unsupported reference features (facet counting, field filters, sort, uncommitted, rewriting, phrase) raise
        NotImplementedError rather than being silently ignored."""
        if query_facets or result_sort or include_uncommitted:
            raise NotImplementedError("facet counts / sort / uncommitted search are outside the GPU hot path")
        # field_filter: names of indexed fields (self.field_names, in schema order) or their indices -> one bitmask
        fmask = 0
        for f in field_filter:
            names = getattr(self, "field_names", [])
            if isinstance(f, str) and f not in names:
                raise ValueError(f"field_filter: unknown indexed field {f!r} (indexed fields: {names})")
            fmask |= 1 << (names.index(f) if isinstance(f, str) else int(f))
        search_mode = search_mode or SearchMode.Lexical()
        ro = ResultObject(original_query=query_string, query=query_string)
        heap = offset + length                       # search.rs:1708 per-shard length = offset+length
        # tokenizer stand-in: whitespace, '+' = mandatory (tokenizer.rs:546-563); unique terms (search.rs:3023-3039)
        toks = query_string.split()
        phrase = query_type_default == QueryType.Phrase
        if len(query_string) >= 2 and query_string.startswith('"') and query_string.endswith('"'):     # "..." = a phrase (tokenizer.rs:546-563)
            phrase, toks = True, query_string[1:-1].split()
        elif any(t.startswith('"') for t in toks):
            raise NotImplementedError("a phrase mixed with other terms is outside the GPU hot path")
        not_terms = [t[1:] for t in toks if t.startswith("-") and len(t) > 1]          # '-' operator: not_query_list (tokenizer.rs:546-563)
        toks = [t for t in toks if not t.startswith("-")]
        qt = query_type_default
        if toks and all(t.startswith("+") for t in toks):
            qt = QueryType.Intersection
        terms = []
        for t in toks:
            t = t.lstrip("+")
            if t and (phrase or t not in terms):       # a phrase keeps its repeated terms: their order is the query
                terms.append(t)
        if phrase and len(terms) >= 2:
            qt = QueryType.Phrase
            if not_terms:
                raise NotImplementedError("NOT terms next to a phrase are outside the GPU hot path")
        elif phrase:
            qt = QueryType.Intersection
        ro.query_terms = list(dict.fromkeys(terms))
        keys = [self.term_key_fn(t) for t in terms]
        nkeys = [self.term_key_fn(t) for t in dict.fromkeys(not_terms)]
        lex, vec, total = [], [], 0
        want_lex = search_mode.kind in ("Lexical", "Hybrid") and len(keys) > 0
        want_vec = search_mode.kind in ("Vector", "Hybrid") and query_vector is not None
        rt = ResultType(result_type)
        if length == 0 and rt == ResultType.TopkCount:   # search.rs:2472-2478
            rt = ResultType.Count
        if want_lex:
            res, counts = self.search_lexical_batch([keys], qt, heap if rt != ResultType.Count else 0, rt, [nkeys] if nkeys else None,
                                                    [list(facet_filter)] if facet_filter else None, [fmask] if fmask else None)
            lex, total = res[0], int(counts[0])
        if want_vec:
            qv = np.asarray(query_vector, dtype=np.float32).reshape(1, -1)
            # similarity_threshold (TopK::new, vector.rs:388-399) and observed_vector_count are handled behind the C-ABI
            am = search_mode.ann_mode or AnnMode.All
            res, _, observed = self.search_vector_ex(qv, max(heap, 1), search_mode.similarity_threshold, ann_mode=am.kind, n_probe=am.n_probe,
                                                     cluster_threshold=am.threshold)
            vec = res[0][:heap]
            ro.observed_vector_count = int(observed[0])
        if search_mode.kind == "Lexical":
            fused = lex
            ro.result_count_total = total
        elif search_mode.kind == "Vector":
            fused = vec
            ro.result_count_total = len(vec)       # vector.rs:1509 (scan-order dependent in the reference)
        else:
            a = _hits_array(max(len(lex), 1)); b = _hits_array(max(len(vec), 1)); o = _hits_array(len(lex) + len(vec) + 1)
            for i, (d, s) in enumerate(lex):
                a[i] = (d, s, 0)
            for i, (d, s) in enumerate(vec):
                b[i] = (d, s, 0)
            n = C.c_uint32(0)
            check(lib().ssb_rrf_fuse(a.ctypes.data, len(lex), b.ctypes.data, len(vec), o.ctypes.data, C.byref(n)))
            fused = [(int(o[i]["doc_id"]), float(o[i]["score"])) for i in range(n.value)]
            ro.result_count_total = total
        # search.rs:2108-2121: drop offset, truncate length
        fused = fused[offset:offset + length] if offset < len(fused) else []
        ro.results = [Result(d, s) for d, s in fused]
        ro.result_count = len(ro.results)
        return ro
