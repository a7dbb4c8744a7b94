/*
 * ssb_oracle.h — CPU ORACLE for the seekstorm_b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's (SeekStorm 3.3.4) query-time arithmetic for
 * BM25 AND / OR / phrase top-k (single- and multi-field BM25F) with the per-candidate chain (delete set, NOT lists, facet filters, field
 * filter), brute-force f32 vector top-k, the int8 quantisers (ScalarQuantizationI8 Cosine / Dot / Euclidean incl. the affine variant,
 * TurboQuantI8) with their scores, the IVF probe (python side) and RRF fusion.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may link or call it.
 * The product path (seekstorm_b200/csrc, libseekstorm_b200.so) never does.
 *
 * PARITY PIN STATUS: the reference cannot be built here (Rust, no toolchain), and its own tests
 * assert only result COUNTS (tests/test.rs:150-208, 693-745) plus aarch64-only kernel-vs-scalar
 * checks (seekstorm/src/vector_similarity.rs:3008-3146).  Those fixtures are reproduced in
 * tests/golden/.  BM25 scores / rank order / cosine scores / RRF scores are "parity unpinned" by
 * the reference; they are pinned here by hand-computed known-answer vectors (tests/golden/) and, for the rows added in round 2, by
 * independent restatements in tests/ (typed numpy columns for the facet filters, substring search over token sequences for phrases,
 * a Hadamard matrix for TurboQuant, exact squared distances for the affine quantiser on integer data).
 *
 * All file:line citations are relative to /root/reference/seekstorm/src/.
 */
#ifndef SSB_ORACLE_H
#define SSB_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same values as include/seekstorm_b200.h */
enum { ORC_QUERY_UNION = 0, ORC_QUERY_INTERSECTION = 1 };
enum { ORC_RESULT_COUNT = 0, ORC_RESULT_TOPK = 1, ORC_RESULT_TOPKCOUNT = 2 };
enum { ORC_SIM_DOT = 0, ORC_SIM_COSINE = 1, ORC_SIM_EUCLIDEAN = 2 };

typedef struct { uint64_t doc_id; float score; uint32_t pad; } orc_hit;

/* One committed level (= one 64K-doc block of one shard) in the neutral layout that
 * include/seekstorm_b200.h's ssb_level_desc uses.  doc_ids ascending per term. */
typedef struct {
    uint32_t level_id;
    uint32_t n_docs;
    uint32_t n_terms;
    uint32_t reserved;
    const uint64_t* term_keys;        /* [n_terms]                         */
    const uint32_t* posting_offsets;  /* [n_terms+1]                       */
    const uint16_t* doc_ids;          /* [n_postings] local ids            */
    const uint16_t* tfs;              /* [n_postings] positions_count      */
    const uint8_t*  doc_len_bytes;    /* [n_docs] byte4 field length codes */
} orc_level;

typedef struct orc_index orc_index;

/* ---- doc-length codec: index.rs:4237-4279 (Lucene SmallFloat byte4) ---- */
uint8_t  orc_int_to_byte4(uint32_t i);
uint32_t orc_byte4_to_int(uint8_t b);

/* ---- BM25 statistics ---- */
/* commit.rs:318-325: cache[b] = K*(1-B+B*DLC[b]/avgdl), avgdl = len_sum as f32 / n_docs as f32 */
void  orc_bm25_cache(uint64_t n_docs, uint64_t len_sum_normalized, float cache[256]);
/* search.rs:3225-3230: ln((N - df + 0.5)/(df + 0.5) + 1) in f32 */
float orc_idf(uint64_t n_docs, uint32_t df);
/* add_result.rs:1450-1452 single term contribution: idf*((tf*(K+1)/(tf+comp))+SIGMA) */
float orc_bm25_term(float idf, uint32_t tf, float comp);

/* ---- index ---- */
#define ORC_MAX_FIELDS 4
orc_index* orc_index_new(void);
/* several indexed fields with boosts (before the first level): tfs become [n_postings][n_fields], doc_len_bytes [n_fields][n_docs];
 * scoring = get_bm25f_multiterm_multifield (add_result.rs:1226-1262) */
int        orc_index_set_fields(orc_index*, uint32_t n_fields, const float* boosts);
void       orc_index_free(orc_index*);
int        orc_index_add_level(orc_index*, const orc_level*);   /* copies everything */
/* global statistics (1-shard semantics); finalises the dictionary */
int        orc_index_commit(orc_index*, uint64_t n_docs, uint64_t len_sum_normalized);
uint32_t   orc_index_df(const orc_index*, uint64_t term_key);
/* shard.delete_hashset (add_result.rs:3435): honoured by the EXHAUSTIVE search only (the pruned variant is the timed baseline) */
int        orc_index_set_deleted(orc_index*, const uint64_t* doc_ids, uint64_t n);

/* Exhaustive BM25 search, canonical tie rule (score desc, doc id asc).
 * Candidate semantics: AND intersection.rs:2023-2301, OR union.rs:1168-1479; score
 * add_result.rs:1429-1482 summed in QUERY ORDER from 0.0; heap min_heap.rs:1193-1259. */
int orc_search_lexical(const orc_index*, const uint64_t* term_keys, uint32_t n_terms,
                       uint32_t query_type, uint32_t k, uint32_t result_type,
                       orc_hit* hits, uint32_t* n_hits, uint64_t* count_total);

/* the same with NOT terms ('-' operator, not_query_list add_result.rs:3440-3496) */
int orc_search_lexical_not(const orc_index*, const uint64_t* term_keys, uint32_t n_terms, const uint64_t* not_keys, uint32_t n_not,
                           uint32_t query_type, uint32_t k, uint32_t result_type,
                           orc_hit* hits, uint32_t* n_hits, uint64_t* count_total);

/* ---- facets and facet filters (same values as include/seekstorm_b200.h): the shard's facet file + FilterSparse per facet
 * (search.rs:863-881), applied by is_facet_filter (add_result.rs:340-478).  Honoured by the EXHAUSTIVE search only. */
enum { ORC_FACET_U8 = 0, ORC_FACET_U16, ORC_FACET_U32, ORC_FACET_U64, ORC_FACET_I8, ORC_FACET_I16, ORC_FACET_I32, ORC_FACET_I64,
       ORC_FACET_TIMESTAMP, ORC_FACET_F32, ORC_FACET_F64, ORC_FACET_STRING16, ORC_FACET_STRING32 };
enum { ORC_FILTER_RANGE = 0, ORC_FILTER_SET = 1 };
#define ORC_MAX_FACETS 16
typedef struct { uint32_t type; uint32_t offset; } orc_facet_field;
/* start / end: the Range<T> bounds widened to 8 bytes (u64 / i64 / f64 bits); SET: set_values[set_first .. +set_count) */
typedef struct { uint32_t facet, kind; uint64_t start, end; uint32_t set_first, set_count; } orc_facet_filter;
int orc_index_set_facets(orc_index*, const void* rows, uint64_t first_doc_id, uint64_t n_docs, uint32_t row_bytes,
                         const orc_facet_field* fields, uint32_t n_fields);
int orc_search_lexical_filtered(const orc_index*, const uint64_t* term_keys, uint32_t n_terms, const uint64_t* not_keys, uint32_t n_not,
                                const orc_facet_filter* filters, uint32_t n_filters, const uint64_t* set_values,
                                uint32_t query_type, uint32_t k, uint32_t result_type,
                                orc_hit* hits, uint32_t* n_hits, uint64_t* count_total);

/* + field filter: bit f of field_mask = indexed field f is in field_filter_set (add_result.rs:3124-3137); 0 = none */
int orc_search_lexical_ex(const orc_index*, const uint64_t* term_keys, uint32_t n_terms, const uint64_t* not_keys, uint32_t n_not,
                          const orc_facet_filter* filters, uint32_t n_filters, const uint64_t* set_values, uint32_t field_mask,
                          uint32_t query_type, uint32_t k, uint32_t result_type,
                          orc_hit* hits, uint32_t* n_hits, uint64_t* count_total);

/* ---- phrase queries (QueryType::Phrase, add_result.rs:3586-3684): positions of the level added last (single field; [sum of tfs], posting
 * order, ascending inside a posting), then seq = the phrase's terms in order, repeats included */
int orc_index_set_last_level_positions(orc_index*, const uint16_t* positions, uint64_t n_positions);
int orc_search_lexical_phrase(const orc_index*, const uint64_t* seq, uint32_t n_seq, uint32_t k, uint32_t result_type,
                              orc_hit* hits, uint32_t* n_hits, uint64_t* count_total);

/* Reference-shaped search: block-max ordered, heap-pruned, same control flow as
 * single.rs:292-417, intersection.rs:2023-2301, union.rs:1168-1479.  Used as the timed CPU baseline
 * ("port") and cross-checked against the exhaustive search in tests. */
int orc_search_lexical_pruned(const orc_index*, const uint64_t* term_keys, uint32_t n_terms,
                              uint32_t query_type, uint32_t k, uint32_t result_type,
                              orc_hit* hits, uint32_t* n_hits, uint64_t* count_total);

/* ---- vectors ---- */
void  orc_normalize_f32(float* v, uint32_t n);                        /* vector_similarity.rs:70-74   */
float orc_dot_f32(const float* a, const float* b, uint32_t n);        /* :1006-1008 scalar, in order  */
float orc_dot_f32_lanes8(const float* q, const float* e, uint32_t n); /* :1120-1142 8-lane FMA order  */
float orc_euclidean_f32(const float* a, const float* b, uint32_t n);  /* :912-918 Σ(x−y)²             */
/* vector.rs:1489-1499: ((score * (1/16129)) + 1) * 0.5, or -score for Euclidean */
float orc_vector_score_postmap(float score, uint32_t similarity);

/* Exhaustive brute-force scan (vector.rs:1397-1467, AnnMode::All) + top-k (vector.rs:410-497),
 * canonical tie rule.  rows are used as given (normalise beforehand for cosine, as the reference does at
 * index time vector.rs:585-596); the query is used as given.  n_threads>1 splits rows (timing only). */
int orc_search_vector(const float* rows, const uint32_t* doc_ids, uint64_t n_rows, uint32_t dims,
                      uint32_t row_pitch_floats, const float* query, uint32_t similarity, uint32_t k,
                      uint32_t use_lanes8, uint32_t n_threads, orc_hit* hits, uint32_t* n_hits);

/* ---- int8 scalar quantisation (Cosine + Quantization::ScalarQuantizationI8) ----
 * index time (vector.rs:585-640): normalize_f32 then quantize_f32_to_i8 = round(v*127) clamped to [-127,127]
 * (vector_similarity.rs:1226-1232; Rust f32::round = half away from zero = roundf); scale = 1.
 * query time: same normalise + quantise; similarity = dot_i8 as f32 (vector_similarity.rs:1011-1016, 193-206). */
void  orc_quantize_f32_to_i8(const float* v, uint32_t n, int8_t* out);
/* normalize_f32 + quantize_f32_to_i8 of every row (what index time does for Cosine + SQ-I8, vector.rs:585-640) */
void  orc_quantize_rows_i8(const float* rows, uint64_t n_rows, uint32_t dims, uint64_t row_pitch_floats, int8_t* out, uint64_t out_pitch);
int32_t orc_dot_i8(const int8_t* a, const int8_t* b, uint32_t n);
int orc_search_vector_i8(const int8_t* rows, const uint32_t* doc_ids, uint64_t n_rows, uint32_t dims, uint32_t row_pitch,
                         const int8_t* query, uint32_t k, orc_hit* hits, uint32_t* n_hits);

/* Dot / Euclidean + ScalarQuantizationI8: QuantizedVector::new_scale / new_scale_norm (vector_similarity.rs:1340-1371),
 * dot_i8_quantized (:1754-1758), -euclidean_i8_quantized (:1721-1734; the non-affine variant, vector.rs:651-660) */
void  orc_quantize_scale_i8(const float* v, uint32_t n, int want_norm, int8_t* out, float* scale_out, float* norm_out);
float orc_score_i8_scaled(const int8_t* q, float q_scale, float q_norm, const int8_t* e, float e_scale, float e_norm, uint32_t n, uint32_t similarity);
int   orc_search_vector_i8_scaled(const int8_t* rows, const float* row_scale, const float* row_norm, const uint32_t* doc_ids, uint64_t n_rows, uint32_t dims,
                                  uint32_t row_pitch, const int8_t* query, float q_scale, float q_norm, uint32_t similarity, uint32_t k,
                                  orc_hit* hits, uint32_t* n_hits);

/* affine Euclidean SQ for integer-valued 0..255 data (vector_similarity.rs:1414-1472, 1770-1795); min_state / max_state are the shard's
 * running min_vector_value / max_vector_value (start: FLT_MAX / -FLT_MAX), updated in place */
void  orc_quantize_affine_i8(const float* v, uint32_t n, float* min_state, float* max_state, int8_t* out, float* scale_out, float* norm_out,
                             int32_t* zero_point_out, int32_t* sum_q_out);
float orc_score_i8_affine(const int8_t* q, float q_scale, float q_norm, int32_t q_zp, int32_t q_sum, const int8_t* e, float e_scale, float e_norm,
                          int32_t e_zp, int32_t e_sum, uint32_t n);
int   orc_search_vector_i8_affine(const int8_t* rows, const float* row_scale, const float* row_norm, const int32_t* row_zp, const int32_t* row_sum,
                                  const uint32_t* doc_ids, uint64_t n_rows, uint32_t dims, uint32_t row_pitch, const int8_t* query, float q_scale, float q_norm,
                                  int32_t q_zp, int32_t q_sum, uint32_t k, orc_hit* hits, uint32_t* n_hits);

/* TurboQuantI8 (vector_similarity.rs:1825-2093): out [dim] codes, dim = next power of two >= n, seed_mask [dim] of +-1 (an input: the
 * reference draws it from ChaCha8Rng(1234), a third-party generator).  Scores: Dot / Cosine = -(dot * s1 * s2) (the reference negates it,
 * :161-176), Euclidean = -max(0, n1 + n2 - 2 * dot * s1 * s2). */
void  orc_turboquant_i8(const float* v, uint32_t n, uint32_t dim, const float* seed_mask, int8_t* out, float* scale_out, float* norm_out);
float orc_score_i8_turbo(const int8_t* q, float q_scale, float q_norm, const int8_t* e, float e_scale, float e_norm, uint32_t dim, uint32_t similarity);
int   orc_search_vector_i8_turbo(const int8_t* rows, const float* row_scale, const float* row_norm, const uint32_t* doc_ids, uint64_t n_rows, uint32_t dim,
                                 uint32_t row_pitch, const int8_t* query, float q_scale, float q_norm, uint32_t similarity, uint32_t k,
                                 orc_hit* hits, uint32_t* n_hits);

/* ---- hybrid: search.rs:1962-2035 RRF k=0.6, rank from 0; then sort score desc (:2097-2121).
 * Tie order in the reference is hash-map iteration order; canonical here: doc id asc. */
int orc_rrf(const orc_hit* lex, uint32_t n_lex, const orc_hit* vec, uint32_t n_vec,
            orc_hit* out /* [n_lex+n_vec] */, uint32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
