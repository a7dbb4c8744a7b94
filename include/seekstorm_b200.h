/*
 * seekstorm_b200.h — C ABI of libseekstorm_b200.so, the B200 (sm_100a) drop-in for the two query-time hot
 * paths behind SeekStorm's Index::search():
 *   (a) BM25 top-k over block-partitioned posting lists (AND / OR with block-max pruning), and
 *   (b) the brute-force f32 dot / cosine / Euclidean vector scan with fused top-k,
 * plus their RRF hybrid.  Plain pointers and sizes only; every pointer argument may be a HOST pointer or a
 * DEVICE pointer of the index's device (detected with cudaPointerGetAttributes) unless stated otherwise.
 *
 * The reference has no FFI for this path; the seams this ABI replaces are (all paths relative to
 * /root/reference/seekstorm/src/):
 *   ssb_search_lexical  <- SearchLexicalShard::search_lexical_shard  search.rs:2427-2458 (body 2445-3767),
 *                          i.e. the kernel-level calls single_blockid / union_docid_2 / union_docid_3 /
 *                          union_blockid / intersection_blockid dispatched at search.rs:3370-3563
 *   ssb_search_vector   <- SearchVectorShard::search_vector_shard    vector.rs:1105-1115 (body 1202-1514)
 *   ssb_search_hybrid   <- Search::search, SearchMode::Hybrid        search.rs:1134-1150, RRF 1962-2035
 *   ssb_lexical_add_level / ssb_lexical_commit <- the committed level as written by commit.rs:203-467 and
 *                          read back by index.rs:3253-3830 (postings, tf = positions_count, byte4 doc lengths)
 *   ssb_vector_add_level <- vector.bin level records written by vector.rs:969-1100
 * INTEGRATION.md shows the Rust `extern "C"` block + shim a maintainer would add.
 *
 * Conventions: every call returns int32_t status (0 = OK, <0 = SSB_E_*), never unwinds, never aborts.
 * Outputs are caller-allocated.  search_* calls on one handle may run CONCURRENTLY from many threads: each takes a
 * search context (CUDA stream + workspaces) from an internal pool, the committed index data is immutable and shared.
 * Index mutation (add_level / commit / set_global_df / set_stream) takes the handle exclusively, like the reference's
 * RwLock around a shard (commit.rs:142, index.rs:5508).
 * Doc ids on the ABI are the reference's shard-local ids: (level << 16) | local (vector.rs:1448,
 * add_result.rs docid = block_id<<16 | local), widened to u64.
 */
#ifndef SEEKSTORM_B200_H
#define SEEKSTORM_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_ABI_VERSION 3
#define SSB_K_MAX 32u            /* top-k capacity of one kernel pass (lane-distributed lists); *_keys calls */
#define SSB_K_LIMIT 1024u        /* ssb_search_lexical / ssb_search_vector page beyond 32 internally         */
#define SSB_MAX_QUERY_TERMS 32u  /* unique terms per lexical query (<= 4: record path; 5..32: one term per lane)  */

enum { SSB_OK = 0, SSB_E_INVALID = -1, SSB_E_CUDA = -2, SSB_E_NOMEM = -3, SSB_E_STATE = -4,
       SSB_E_UNSUPPORTED = -5, SSB_E_NO_DEVICE = -6 };

/* QueryType (search.rs, enum QueryType): Union / Intersection / Phrase.  A PHRASE batch lists every query's terms in phrase order,
 * repeated terms included ("to be or not to be" = 6 keys); a doc matches when it contains all of them and token i occurs at position
 * p + i for some p (add_result.rs:3586-3684); scores and counts as for an intersection of the unique terms.  Needs levels loaded with
 * ssb_level_desc.positions (one indexed field); NOT terms are not accepted in a phrase batch. */
enum { SSB_QUERY_UNION = 0, SSB_QUERY_INTERSECTION = 1, SSB_QUERY_PHRASE = 2 };
/* ResultType (search.rs:150-175): Count / Topk / TopkCount (default) */
enum { SSB_RESULT_COUNT = 0, SSB_RESULT_TOPK = 1, SSB_RESULT_TOPKCOUNT = 2 };
/* VectorSimilarity (vector_similarity.rs:20-30) */
enum { SSB_SIM_DOT = 0, SSB_SIM_COSINE = 1, SSB_SIM_EUCLIDEAN = 2 };
/* Quantization (vector_similarity.rs `Quantization`): SCALAR_I8 = ScalarQuantizationI8.
 *   Cosine   : rows and queries are normalised, then quantised round(v*127) clamped to [-127,127] (vector_similarity.rs:1226-1232,
 *              vector.rs:585-640); score = the exact int32 dot product as f32 (vector_similarity.rs:193-206, 1011-1016).
 *   Dot      : per-vector scale = max|x|/127, codes round(x/scale) (QuantizedVector::new_scale, vector_similarity.rs:1340-1353);
 *              score = dot_i32 as f32 * query_scale * row_scale (dot_i8_quantized, :1754-1758).
 *   Euclidean: new_scale_norm (:1356-1371), the NON-AFFINE variant the reference uses for non-integer data (vector.rs:651-660);
 *              score = -max(0, query_norm + row_norm - 2*dot) (euclidean_i8_quantized, :1721-1734).  When the FIRST vector of the index is
 *              integer-valued in 0..255 (SIFT-like data) the reference switches the shard to its AFFINE quantiser for good
 *              (new_scale_norm_affine, :1414-1463: codes = round(x / scale) + zero_point with scale / zero point from the running min / max
 *              of everything indexed so far; euclidean_i8_quantized_affine, :1770-1795) — so does this library: rows must then be added
 *              in the reference's ingestion order, queries are quantised with the state the index has reached.
 * All three are bit-exact with the scalar CPU arithmetic (integer accumulation on tcgen05 kind::i8, reference operation order). */
enum { SSB_QUANT_NONE = 0, SSB_QUANT_SCALAR_I8 = 1,
       /* TurboQuantI8 (vector_similarity.rs:1825-2093): every vector (after normalize_f32 for Cosine) is zero-padded to the next power of
        * two, sign-flipped by the index's seed mask, rotated by the normalised fast Walsh-Hadamard transform and quantised with
        * scale = max(||x|| / sqrt(dim) / 32, 1e-8); rows are stored at next_power_of_two(dims) bytes.  Scores as in the reference:
        * Dot / Cosine = -(dot_i32 as f32 * query_scale * row_scale) (it negates the estimate, :161-176, 220-235), Euclidean =
        * -max(0, query_norm + row_norm - 2 * dot_i32 as f32 * query_scale * row_scale) (:2058-2069).  Bit-exact with the scalar CPU
        * arithmetic.  Needs ssb_vector_set_turboquant_mask before the first level. */
       SSB_QUANT_TURBO_I8 = 2 };
/* which vector scan kernel to use */
/* FFMA: packed-FP32 scan, 16 queries per corpus pass (HBM-bound).  TCGEN05[_N64]: tensor-core scan with the 3xTF32
 * split, 128 (or 64) queries per corpus pass.  TCGEN05_BF16[_N64]: tensor-core scan with the 3xBF16 split (half the
 * operand bytes; score error ~1e-5 relative, inside the 1e-4 tolerance).  AUTO: FFMA up to 16 queries and
 * for Euclidean; above that TCGEN05_BF16, or TCGEN05_BF16_N256 when its passes take less time for the batch size. */
enum { SSB_VEC_KERNEL_AUTO = 0, SSB_VEC_KERNEL_FFMA = 1, SSB_VEC_KERNEL_TCGEN05 = 2, SSB_VEC_KERNEL_TCGEN05_N64 = 3,
       SSB_VEC_KERNEL_TCGEN05_BF16 = 4, SSB_VEC_KERNEL_TCGEN05_BF16_N64 = 5,
       SSB_VEC_KERNEL_TCGEN05_BF16_N256 = 6 /* 256 queries per corpus pass: half the HBM bytes per query, tensor / shared-memory bound */,
       /* FILTER: one bf16 product over the 2-byte hi plane of the corpus selects, with a proven error margin, the <= 32 rows that can be
        * in the top-k (k <= 16); those are re-scored with the plain f32 dot product and queries whose candidate set did not fit are re-run
        * by an exact f32 scan on the device.  Results are the exact f32 top-k.  128 / 256 queries per corpus pass. */
       SSB_VEC_KERNEL_TCGEN05_FILTER = 7, SSB_VEC_KERNEL_TCGEN05_FILTER_N256 = 8,
       /* the 256-query filter scan on CTA pairs (tcgen05 cta_group::2, clusters of 2): the two SMs of a pair share one copy of the query block */
       SSB_VEC_KERNEL_TCGEN05_FILTER_N256_PAIR = 9 };

typedef struct ssb_index ssb_index;

/* min_heap.rs:17-40 `Result` {doc_id, score}; 16 bytes */
typedef struct { uint64_t doc_id; float score; uint32_t pad; } ssb_hit;
/* the `vb`-feature fields of `Result` (min_heap.rs:21-39), parallel to an ssb_hit array; 48 bytes.
 * source: ResultSource (Lexical / Vector / Hybrid). */
enum { SSB_SOURCE_LEXICAL = 0, SSB_SOURCE_VECTOR = 1, SSB_SOURCE_HYBRID = 2 };
typedef struct {
    uint32_t field_id, chunk_id, level_id, shard_id, cluster_id;
    float cluster_score, vector_score, lexical_score;
    uint32_t source; uint32_t pad[3];
} ssb_hit_ext;

typedef struct {
    int32_t  device;             /* CUDA device ordinal                                                  */
    uint32_t max_batch;          /* max queries per search_* call (workspace sizing); 0 -> 4096          */
    uint32_t vector_dims;        /* 0 = no vector index                                                   */
    uint32_t vector_similarity;  /* SSB_SIM_*  (meta.inference similarity)                               */
    uint32_t vector_kernel;      /* SSB_VEC_KERNEL_*                                                      */
    uint32_t vector_quantization;/* SSB_QUANT_* (0 = f32 corpus)                                           */
    uint32_t reserved[2];
} ssb_config;

/* One committed level = one 64K-doc block of the shard, in a neutral (decoded) layout.
 * Replaces the per-level bytes the reference keeps in index.bin (ARCHITECTURE.md:77-82):
 *   doc_ids / tfs : per term, ascending local doc ids and positions_count (tf), what
 *                   intersection.rs:199-247 + add_result.rs:2036-2197 decode per candidate
 *   doc_len_bytes : the level's byte4 field-length array (index.rs:770-776)                    */
typedef struct {
    uint32_t level_id;                /* block id; doc_id = level_id<<16 | local                         */
    uint32_t n_docs;                  /* <= 65536                                                         */
    uint32_t n_terms;
    uint32_t n_fields;                /* indexed fields: 0 / 1 = one field; F > 1 needs ssb_lexical_set_field_boosts(ix, F, ..) first  */
    const uint64_t* term_keys;        /* [n_terms] 64-bit term hash (reference key_hash), any order       */
    const uint32_t* posting_offsets;  /* [n_terms+1]                                                      */
    const uint16_t* doc_ids;          /* [n_postings] ascending within a term                            */
    const uint16_t* tfs;              /* [n_postings]; F fields: [n_postings][F], 0 = the term does not occur in that field           */
    const uint8_t*  doc_len_bytes;    /* [n_docs];     F fields: [F][n_docs] (document_length_compressed_array[field], index.rs:770-776) */
    const uint16_t* positions;        /* [sum of tfs] or NULL: the term positions of every posting, in posting order, ascending inside a   */
                                      /* posting (what get_next_position_singlefield decodes, add_result.rs:2036-2197); needed by          */
                                      /* SSB_QUERY_PHRASE only; either every level carries them or none; one indexed field                  */
} ssb_level_desc;

/* ---- facets and facet filters ---------------------------------------------------- */
/* FieldType of a facet field (index.rs `FieldType`; FilterSparse search.rs:863-881).  Point (geo) filters are not built. */
enum { SSB_FACET_U8 = 0, SSB_FACET_U16 = 1, SSB_FACET_U32 = 2, SSB_FACET_U64 = 3, SSB_FACET_I8 = 4, SSB_FACET_I16 = 5,
       SSB_FACET_I32 = 6, SSB_FACET_I64 = 7, SSB_FACET_TIMESTAMP = 8, SSB_FACET_F32 = 9, SSB_FACET_F64 = 10,
       SSB_FACET_STRING16 = 11, SSB_FACET_STRING32 = 12 };
/* one facet field of the shard's facet file: its type and its byte offset inside a doc's row (`facet.offset`, add_result.rs:345) */
typedef struct { uint32_t type; uint32_t offset; } ssb_facet_field;
/* One filter of one query.  RANGE = Rust `Range<T>::contains`: start <= value < end, compared in the facet's own type (floats by
 * PartialOrd: NaN is never inside).  start / end carry the bound widened to 8 bytes: unsigned types as u64, signed types and
 * Timestamp as i64 (two's complement), F32 / F64 as the bits of the f64 value.  SET (String16 / String32) = `values.contains(id)`
 * over filter_set_values[set_first .. set_first + set_count).  A facet without a filter entry is FilterSparse::None. */
enum { SSB_FILTER_RANGE = 0, SSB_FILTER_SET = 1 };
typedef struct ssb_facet_filter {
    uint32_t facet;                   /* index into the fields given to ssb_set_facets                    */
    uint32_t kind;                    /* SSB_FILTER_*                                                     */
    uint64_t start, end;              /* RANGE bounds (see above)                                         */
    uint32_t set_first, set_count;    /* SET: slice of ssb_lex_batch.filter_set_values                    */
} ssb_facet_filter;
#define SSB_MAX_FACETS 16u
#define SSB_MAX_FILTERS_PER_QUERY 16u

/* A batch of lexical queries, already tokenised by the host (tokenizer.rs is out of scope): unique terms
 * per query as 64-bit keys, CSR layout; at most SSB_MAX_QUERY_TERMS per query (checked for host arrays; with device
 * arrays extra terms are ignored).  Repeated keys inside a query count once, as in the reference's unique_terms. */
#define SSB_TERM_NOT 1u            /* term_flags bit 0: the '-' operator — docs containing the term are excluded (not_query_list,  */
                                  /* add_result.rs:3440-3496); NOT terms neither score nor count towards the 32-term limit       */
#define SSB_MAX_NOT_TERMS 4u
typedef struct {
    uint32_t n_queries;
    uint32_t query_type;              /* SSB_QUERY_* (applies to the whole batch)                         */
    const uint32_t* term_offsets;     /* [n_queries+1]                                                    */
    const uint64_t* term_keys;        /* [term_offsets[n_queries]]                                        */
    const uint8_t*  term_flags;       /* [term_offsets[n_queries]] SSB_TERM_* per term, or NULL (all positive); at most        */
                                      /* SSB_MAX_NOT_TERMS NOT terms per query                                                 */
    /* facet filters (ABI v3; `facet_filter: Vec<FacetFilter>` of search_lexical_shard, search.rs:2427-2458, resolved to one      */
    /* FilterSparse per facet, search.rs:863-881): HOST arrays or NULL.  A doc enters neither the top-k nor the counts unless      */
    /* every filter of its query accepts its facet value (is_facet_filter, add_result.rs:340-478).  Needs ssb_set_facets.          */
    const uint32_t* filter_offsets;           /* [n_queries+1] or NULL (no query is filtered)                                      */
    const struct ssb_facet_filter* filters;   /* [filter_offsets[n_queries]]                                                       */
    const uint64_t* filter_set_values;        /* value ids of the SSB_FILTER_SET filters (String16 / String32), or NULL            */
    /* field filter (`field_filter: Vec<String>` -> field_filter_set, add_result.rs:3124-3137, 3558-3571): HOST array [n_queries] or   */
    /* NULL; bit f = indexed field f is in the query's filter, 0 = no field filter.  A doc is dropped when a query term it contains    */
    /* occurs in none of the filter's fields (tested like the reference only when term fields + filter fields <= indexed fields);      */
    /* scores still sum every field.  Indexes with one indexed field ignore it (the reference's test can never fire there).            */
    const uint32_t* field_masks;
} ssb_lex_batch;

uint32_t    ssb_abi_version(void);
const char* ssb_last_error(void);                              /* thread-local, never NULL              */

int32_t ssb_create(const ssb_config* cfg, ssb_index** out);
int32_t ssb_destroy(ssb_index* ix);

/* ---- lexical index: add immutable levels, then commit global statistics --------------------------- */
int32_t ssb_lexical_add_level(ssb_index* ix, const ssb_level_desc* level);
/* n_docs = indexed_doc_count, len_sum_normalized = positions_sum_normalized (commit.rs:318-319) of the
 * WHOLE shard (all GPUs' levels).  Builds the dictionary, bm25_component_cache (commit.rs:321-325),
 * per-(term,block) block-max (index.rs:2938-3049) and the bitmap containers for dense lists. */
/* Several indexed fields (BM25F, get_bm25f_multiterm_multifield add_result.rs:1171-1426): n_fields <= 4 and the per-field boosts
 * (indexed_schema_vec[f].boost; NULL = 1.0) — before the first ssb_lexical_add_level.  Score of a doc = sum over query terms (query
 * order) and over the fields the term occurs in (ascending) of boost_f * idf_t * tf_tf*(K+1)/(tf_tf + cache[len_byte_f(doc)]),
 * accumulated left to right like the reference.  Pruning bounds use an upper bound of the boost-weighted per-posting sum; such an
 * index is searched by the general (one term per lane) kernel. */
int32_t ssb_lexical_set_field_boosts(ssb_index* ix, uint32_t n_fields, const float* boosts);
int32_t ssb_lexical_commit(ssb_index* ix, uint64_t n_docs, uint64_t len_sum_normalized);
/* dictionary export / global document-frequency override (multi-GPU block-range sharding: idf uses the
 * global df, search.rs:3225).  keys/dfs are host pointers. */
int32_t ssb_lexical_dict_size(const ssb_index* ix, uint64_t* n_terms);
int32_t ssb_lexical_dict_export(const ssb_index* ix, uint64_t* keys, uint32_t* dfs, uint64_t cap);
int32_t ssb_lexical_set_global_df(ssb_index* ix, const uint64_t* keys, const uint32_t* dfs, uint64_t n);

/* ---- loading the reference's own shard files ------------------------------------ */
/* index.bin of ONE shard as written by commit.rs:203-467 / read by open_shard (index.rs:3253-3516): header, then per 64K-doc level
 * the byte4 length array, the cumulative statistics, the segment table, key heads (compress_postinglist.rs:339-409) and key bodies
 * (Array / Bitmap / RLE doc-id containers, compress_postinglist.rs:694-977; tf from the rank-position pointers,
 * add_result.rs:2036-2197).  Adds every level and commits with the file's own indexed_doc_count / positions_sum_normalized.
 * bytes: HOST memory (e.g. the mmap of the file).  params come from the shard's index.json / schema.json. */
typedef struct {
    uint32_t indexed_field_count;   /* schema: indexed fields; only 1 is supported by the loader (multi-field BM25F: neutral layout, ssb_lexical_set_field_boosts)        */
    uint32_t key_head_size;         /* 20 without n-gram indexing, 22 / 23 with bigram / trigram df bytes                      */
    uint32_t segment_number_bits;   /* 11 (create_shard(.., 11, ..), index.rs:3295): 2048 dictionary segments per level        */
    uint32_t decode_positions;      /* != 0: also decode every posting's term positions (embedded layouts index_posting.rs:590-640, VINT delta  */
                                    /* blobs compress_postinglist.rs:946-977) and load them for SSB_QUERY_PHRASE; positions above 65535 fail   */
} ssb_index_bin_params;
int32_t ssb_load_index_bin(ssb_index* ix, const void* bytes, uint64_t len, const ssb_index_bin_params* params, uint64_t* n_docs_out);
/* host-only walk of an index.bin (no GPU needed): out = {levels, single-term keys, postings, sum of tf, indexed_doc_count,
 * positions_sum_normalized, FNV checksum over every (key, level, doc id, tf) in file order, FNV checksum over every decoded
 * position (decode_positions) or 0} */
int32_t ssb_index_bin_inspect(const void* bytes, uint64_t len, const ssb_index_bin_params* params, uint64_t out[8]);
/* vector.bin of one shard (vector.rs:1066-1094; Precision::F32 records of 24 + 4*dims bytes); dims = the index's vector_dims */
int32_t ssb_load_vector_bin(ssb_index* ix, const void* bytes, uint64_t len, uint64_t* n_vectors_out);

/* ---- delete set ---------------------------------------------------------------------------------------- */
/* shard.delete_hashset (index.rs:1594; delete_document index.rs:5110): deleted docs are neither scored nor counted, in the
 * lexical path (add_result.rs:3435, union_count union.rs:975-1000) and in the vector scan (vector.rs:1450-1451).  doc_ids: host
 * array of shard-local ids (level << 16 | local); replaces the current set; n = 0 clears it.  Exclusive like a commit. */
int32_t ssb_set_deleted(ssb_index* ix, const uint64_t* doc_ids, uint64_t n);

/* ---- facets ---------------------------------------------------------------------------------------------- */
/* The shard's facet file (`facets_file_mmap`: one row of `facets_size_sum` bytes per doc, doc id = level << 16 | local;
 * is_facet_filter reads `row_bytes * docid + field.offset`, add_result.rs:343-347).  rows: HOST array [n_docs * row_bytes] holding
 * the rows of doc ids first_doc_id .. first_doc_id + n_docs (a shard of a sharded index passes the rows of its own level range).
 * Every value is converted once into an order-preserving 64-bit key and kept as one column per facet in HBM (8 bytes per doc and
 * facet); a doc outside the covered range fails every filter.  Replaces the current facets; n_docs = 0 clears them.  Exclusive. */
int32_t ssb_set_facets(ssb_index* ix, const void* rows, uint64_t first_doc_id, uint64_t n_docs, uint32_t row_bytes,
                       const ssb_facet_field* fields, uint32_t n_fields);

/* ---- vector index -------------------------------------------------------------------------------- */
/* rows: [n, dims] row-major f32 (row_stride_floats >= dims, 0 = dims); local_ids: [n] u16 or NULL (= 0..n-1).
 * Cosine: rows are L2-normalised on load (vector.rs:585-596 does this at index time). */
int32_t ssb_vector_add_level(ssb_index* ix, uint32_t level_id, const float* rows, uint64_t row_stride_floats,
                             const uint16_t* local_ids, uint32_t n, uint32_t dims);
/* The same level with its IVF cluster table (vector.bin: `u32 clusters; u32 child_count x clusters; records`, vector.rs:1066-1094): rows
 * are in cluster order, cluster c holds the next cluster_counts[c] rows, its medoid is its FIRST row (vector.rs:1316-1320).  Levels added
 * without a table are one cluster (what the reference writes below 100 vectors or with Clustering::None, vector.rs:1048-1062).  The
 * clusters only matter to ssb_search_vector_ex calls with ann_mode != SSB_ANN_ALL.  f32 indexes only. */
int32_t ssb_vector_add_level_clustered(ssb_index* ix, uint32_t level_id, const float* rows, uint64_t row_stride_floats,
                                       const uint16_t* local_ids, uint32_t n, uint32_t dims,
                                       const uint32_t* cluster_counts, uint32_t n_clusters);
/* TurboQuantI8 indexes: the index's sign mask `TurboQuant.seed_mask` (dim = next_power_of_two(vector_dims) values of +1 / -1).  The
 * reference draws it once per index from ChaCha8Rng::seed_from_u64(1234) (vector_similarity.rs:1845-1859, index.rs:2215-2216) — a
 * third-party generator (rand_chacha) that this library does not restate: the host hands over the mask it holds.  Before the first level. */
int32_t ssb_vector_set_turboquant_mask(ssb_index* ix, const float* seed_mask, uint32_t dim);
int32_t ssb_vector_count(const ssb_index* ix, uint64_t* n_rows);
/* capacity hint: size the vector arenas for n_rows rows up front (loading level by level otherwise grows them geometrically) */
int32_t ssb_vector_reserve(ssb_index* ix, uint64_t n_rows);
/* switch the scan kernel (SSB_VEC_KERNEL_*) of an existing index */
int32_t ssb_set_vector_kernel(ssb_index* ix, uint32_t vector_kernel);

/* ---- search ---------------------------------------------------------------------------------------- */
/* hits: [n_queries * k] best-first (score desc, doc id asc); n_hits: [n_queries]; count_total: [n_queries] or
 * NULL (result_count_total: exact for Count/TopkCount, unspecified for Topk — search.rs:196-198). */
int32_t ssb_search_lexical(ssb_index* ix, const ssb_lex_batch* q, uint32_t k, uint32_t result_type,
                           ssb_hit* hits, uint32_t* n_hits, uint64_t* count_total);
/* queries: [n_queries, dims] f32; Cosine: normalised by the callee (search.rs:1464-1475).  score = dot
 * (Dot/Cosine) or -Σ(q-x)² (Euclidean) exactly as Result.score in vector.rs:1489. */
int32_t ssb_search_vector(ssb_index* ix, const float* queries, uint32_t n_queries, uint32_t k,
                          ssb_hit* hits, uint32_t* n_hits);
/* search_vector_shard with all of its arguments (vector.rs:1105-1115): `similarity_threshold: Option<f32>` (pre-mapped as in
 * TopK::new, vector.rs:388-399: (2t-1)*16129 for Dot/Cosine, -t for Euclidean; hits scoring below it are dropped), queries as
 * f32 or — ScalarQuantizationI8 indexes only — as the int8 codes the reference's server holds after quantize_f32_to_i8
 * (search.rs:1477-1490).  ext: [n_queries * k] or NULL (vb fields: level_id, vector_score post-map vector.rs:1495-1499, source);
 * observed: [n_queries] or NULL (observed_vector_count: every record under AnnMode::All).  Rows that share a doc id (one vector
 * per chunk) are collapsed to the best-scoring one, as TopK::push does (vector.rs:436-470). */
enum { SSB_QFMT_F32 = 0, SSB_QFMT_I8 = 1 };
/* AnnMode (vector_similarity.rs:43-66) — which IVF clusters of each level are searched (vector.rs:1300-1392): per (query, level) the
 * query is scored against every cluster's medoid, the n_probe best clusters (score desc, cluster id asc) whose medoid score is not below
 * the pre-mapped cluster threshold are scanned, the others are skipped.  observed = the vectors in the selected clusters. */
enum { SSB_ANN_ALL = 0, SSB_ANN_NPROBE = 1, SSB_ANN_SIMILARITY_THRESHOLD = 2, SSB_ANN_NPROBE_SIMILARITY_THRESHOLD = 3 };
typedef struct {
    const void* queries;          /* [n_queries, dims] f32 or i8, host or device                              */
    uint32_t n_queries, k;
    uint32_t query_format;        /* SSB_QFMT_*                                                               */
    uint32_t has_threshold;       /* 0 = None                                                                 */
    float    similarity_threshold;
    uint32_t ann_mode;            /* SSB_ANN_* (0 = All: exhaustive)                                          */
    uint32_t n_probe;             /* Nprobe / NprobeSimilaritythreshold                                       */
    float    cluster_threshold;   /* Similaritythreshold / NprobeSimilaritythreshold (pre-mapped like similarity_threshold) */
} ssb_vec_query;
int32_t ssb_search_vector_ex(ssb_index* ix, const ssb_vec_query* q, ssb_hit* hits, uint32_t* n_hits, ssb_hit_ext* ext,
                             uint64_t* observed);
/* SearchMode::Hybrid: both searches with length k, RRF (k=0.6, rank from 0), sort, truncate to k.
 * hits: [n_queries * k]. */
int32_t ssb_search_hybrid(ssb_index* ix, const ssb_lex_batch* q, const float* queries, uint32_t k,
                          ssb_hit* hits, uint32_t* n_hits);

/* RRF on two host lists (search.rs:1962-2035): out capacity n_lex+n_vec; sorted score desc, doc id asc. */
int32_t ssb_rrf_fuse(const ssb_hit* lex, uint32_t n_lex, const ssb_hit* vec, uint32_t n_vec,
                     ssb_hit* out, uint32_t* n_out);

/* ---- device-resident variants (multi-GPU merge, benchmarking with inputs already in HBM) ----------- */
/* Packed top-k keys: u64 = (ordered(score) << 32) | (0xFFFFFFFF - doc_id); larger = better.  keys_out is a
 * DEVICE buffer [n_queries * 32]; entry j of query i is its j-th best or 0 if none.  Asynchronous on the
 * index stream; ssb_sync() waits. */
int32_t ssb_search_vector_keys(ssb_index* ix, const float* queries, uint32_t n_queries, uint32_t k,
                               uint64_t* keys_out_dev);
int32_t ssb_search_lexical_keys(ssb_index* ix, const ssb_lex_batch* q, uint32_t k, uint32_t result_type,
                                uint64_t* keys_out_dev, uint64_t* count_total_dev);
/* merge `n_lists` key lists per query ([n_lists][n_queries][32], device) into hits (host) */
int32_t ssb_merge_keys(ssb_index* ix, const uint64_t* keys_dev, uint32_t n_lists, uint32_t n_queries,
                       uint32_t k, ssb_hit* hits, uint32_t* n_hits);
/* ---- sharded index over several GPUs: one process per GPU, each handle holds a contiguous range of levels ---- */
/* The reference fans a query out over its shards and concatenates + sorts their results (search.rs:1637-1743, 1875-1928, 2097-2106).
 * Here every rank calls the same ssb_search_* with the same queries; once a communicator is set the library itself enqueues the
 * exchange on the search stream — ncclAllGather of the packed top-k keys + the G*k -> k merge, ncclAllReduce of the match
 * counts; hybrid: both lists are merged over the shards first and fused (RRF) afterwards — and every rank returns the GLOBAL
 * result.  Searches on a handle with a communicator are serialised (collectives must be issued in the same order everywhere).
 * NCCL is resolved at run time (the copy already loaded in the process, libnccl.so.2, or $SSB_NCCL_LIB). */
#define SSB_COMM_ID_BYTES 128
int32_t ssb_comm_unique_id(uint8_t* id128);                       /* ncclGetUniqueId; rank 0 calls it and ships the bytes to the others */
int32_t ssb_comm_init(ssb_index* ix, const uint8_t* id128, uint32_t rank, uint32_t world);   /* collective: ncclCommInitRank      */
int32_t ssb_comm_attach(ssb_index* ix, void* nccl_comm, uint32_t rank, uint32_t world);      /* borrow the caller's ncclComm_t      */
int32_t ssb_comm_destroy(ssb_index* ix);
/* collective: install the index-wide document frequencies (idf uses the df of the whole index, search.rs:3225-3230; the commit
 * already received the global N and length sum).  After it, scores equal those of an unsharded index bit for bit. */
int32_t ssb_lexical_sync_df(ssb_index* ix);

int32_t ssb_sync(ssb_index* ix);
/* the CUDA stream the index launches on (cudaStream_t as void*), for event timing */
void*   ssb_stream(ssb_index* ix);
/* run all further work of this index on a caller-owned stream (e.g. the stream NCCL collectives are enqueued
 * on, so the per-GPU top-k -> all-gather -> merge chain needs no host synchronisation).  The value is used
 * as a cudaStream_t as is (NULL = the CUDA legacy default stream); SSB_OWN_STREAM restores the index's own
 * stream. */
#define SSB_OWN_STREAM ((void*)(intptr_t)-1)
int32_t ssb_set_stream(ssb_index* ix, void* cuda_stream);

/* ---- statistics of the last search_* call (for roofline accounting) -------------------------------- */
typedef struct {
    uint64_t kernel_launches;     /* kernels launched by the last call                                   */
    uint64_t algorithmic_bytes;   /* SURVEY.md §8(d) bytes the call's kernels had to move                 */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t postings_visited;    /* lexical: driver postings enumerated after pruning                    */
    uint64_t probes;              /* lexical: membership probes into other lists                          */
    uint64_t items_processed;     /* lexical: (query, block) work items executed                          */
    uint64_t items_skipped;       /* lexical: work items pruned by block-max                              */
    uint64_t dominant_kernel_ns;  /* CUDA-event duration of the call's dominant kernel (scan / scoring)   */
    uint64_t scan_bytes_read;     /* vector: bytes the scan (+ refine) kernels stream from HBM by construction */
    uint64_t filter_fallbacks;    /* vector, host-facing calls: queries of the filter scan that took the exact fallback scan */
    uint64_t reserved[1];
} ssb_stats;
int32_t ssb_last_stats(const ssb_index* ix, ssb_stats* out);

#ifdef __cplusplus
}
#endif
#endif
