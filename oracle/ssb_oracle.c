/*
 * ssb_oracle.c — CPU ORACLE (test infrastructure only; see ssb_oracle.h header comment).
 *
 * Restates, function by function, the reference's query-time arithmetic.  Citations are
 * relative to /root/reference/seekstorm/src/.  Compile with -ffp-contract=off so that every f32
 * operation is individually rounded like rustc's output (no FMA contraction), except where the
 * reference itself uses an explicit FMA (dot_f32_avx2).
 *
 * PARITY: "unpinned" by the reference's own tests beyond result counts; pinned by the hand-computed
 * known-answer vectors in tests/golden/ (see tests/test_oracle_golden.py).
 */
#include "ssb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* add_result.rs:20-22 */
static const float ORC_K = 1.2f;
static const float ORC_B = 0.75f;
static const float ORC_SIGMA = 0.0f;

/* ------------------------------------------------------------------ codec: index.rs:4232-4279 */
#define NUM_FREE_VALUES 24u

uint8_t orc_int_to_byte4(uint32_t i) {
    if (i < NUM_FREE_VALUES) return (uint8_t)i;
    uint32_t ii = i - NUM_FREE_VALUES;
    uint32_t num_bits = ii ? 32u - (uint32_t)__builtin_clz(ii) : 0u;
    if (num_bits < 4) return (uint8_t)(NUM_FREE_VALUES + ii);
    uint32_t shift = num_bits - 4;
    return (uint8_t)(NUM_FREE_VALUES + (((ii >> shift) & 0x07u) | ((shift + 1) << 3)));
}

uint32_t orc_byte4_to_int(uint8_t b) {
    if ((uint32_t)b < NUM_FREE_VALUES) return b;
    uint32_t i = (uint32_t)b - NUM_FREE_VALUES;
    uint32_t bits = i & 0x07u, shift = i >> 3;
    if (shift == 0) return NUM_FREE_VALUES + bits;
    return NUM_FREE_VALUES + ((bits | 0x08u) << (shift - 1));
}

/* ------------------------------------------------------------------ statistics */
void orc_bm25_cache(uint64_t n_docs, uint64_t len_sum, float cache[256]) {
    /* commit.rs:318-325 */
    float avgdl = (float)len_sum / (float)n_docs;
    for (int i = 0; i < 256; i++) {
        float q = (float)orc_byte4_to_int((uint8_t)i) / avgdl;
        cache[i] = ORC_K * (1.0f - ORC_B + ORC_B * q);
    }
}

float orc_idf(uint64_t n_docs, uint32_t df) {
    /* search.rs:3225-3230; Rust f32::ln == logf */
    return logf((((float)n_docs - (float)df + 0.5f) / ((float)df + 0.5f)) + 1.0f);
}

float orc_bm25_term(float idf, uint32_t tf_u, float comp) {
    /* add_result.rs:1450-1452 */
    float tf = (float)tf_u;
    return idf * ((tf * (ORC_K + 1.0f) / (tf + comp)) + ORC_SIGMA);
}

/* ------------------------------------------------------------------ index */
typedef struct {
    uint32_t level_id, n_docs, n_terms;
    uint64_t* term_keys;
    uint32_t* posting_offsets;
    uint16_t* doc_ids;
    uint16_t* tfs;
    uint8_t* doc_len_bytes;
    float* max_comp; /* [n_terms] max over postings of tf*(K+1)/(tf+cache[len]) (block-max basis) */
    uint16_t* positions; uint32_t* pos_off;   /* term positions of every posting, in posting order (phrase queries); pos_off = prefix sums of tfs */
} lvl_t;

typedef struct { uint64_t key; uint32_t level, idx; } dict_ent;

struct orc_index {
    lvl_t* levels;
    uint32_t n_levels, cap_levels;
    dict_ent* dict; /* sorted by (key, level) */
    uint64_t n_dict;
    uint64_t n_docs, len_sum;
    float cache[256];
    int committed;
    uint32_t n_fields;            /* indexed fields (0 / 1 = single field) */
    float boost[ORC_MAX_FIELDS];  /* indexed_schema_vec[f].boost (add_result.rs:1258) */
    uint64_t* deleted; uint64_t n_deleted;   /* shard.delete_hashset, sorted (add_result.rs:3435) */
    /* facets_file_mmap: one row of row_bytes per doc id (facet_first_doc + row), typed fields at their offsets (add_result.rs:343-347) */
    uint8_t* facet_rows; uint64_t facet_n_docs, facet_first_doc; uint32_t facet_row_bytes, n_facets;
    orc_facet_field facet_fields[ORC_MAX_FACETS];
};

orc_index* orc_index_new(void) { return (orc_index*)calloc(1, sizeof(orc_index)); }

/* several indexed fields (before the first level): every level then carries tfs[n_postings][n_fields] (0 = the term does not occur in
 * that field) and doc_len_bytes[n_fields][n_docs]; scores follow get_bm25f_multiterm_multifield (add_result.rs:1226-1262) */
int orc_index_set_fields(orc_index* ix, uint32_t n_fields, const float* boosts) {
    if (!ix || ix->n_levels || n_fields == 0 || n_fields > ORC_MAX_FIELDS) return -1;
    ix->n_fields = n_fields;
    for (uint32_t f = 0; f < n_fields; f++) ix->boost[f] = boosts ? boosts[f] : 1.0f;
    return 0;
}
static inline uint32_t nf_of(const orc_index* ix) { return ix->n_fields > 1 ? ix->n_fields : 1; }

void orc_index_free(orc_index* ix) {
    if (!ix) return;
    for (uint32_t i = 0; i < ix->n_levels; i++) {
        lvl_t* l = &ix->levels[i];
        free(l->term_keys); free(l->posting_offsets); free(l->doc_ids); free(l->tfs);
        free(l->doc_len_bytes); free(l->max_comp); free(l->positions); free(l->pos_off);
    }
    free(ix->levels); free(ix->dict); free(ix->deleted); free(ix->facet_rows); free(ix);
}

static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : (x > y); }
/* delete set: docs in it are neither scored nor counted by orc_search_lexical (add_result.rs:3435; union_count union.rs:975-1000) */
int orc_index_set_deleted(orc_index* ix, const uint64_t* doc_ids, uint64_t n) {
    if (!ix) return -1;
    free(ix->deleted); ix->deleted = NULL; ix->n_deleted = 0;
    if (!n) return 0;
    ix->deleted = (uint64_t*)malloc(n * 8); memcpy(ix->deleted, doc_ids, n * 8); ix->n_deleted = n;
    qsort(ix->deleted, n, 8, cmp_u64);
    return 0;
}
static inline int is_deleted(const orc_index* ix, uint64_t doc) {
    uint64_t lo = 0, hi = ix->n_deleted;
    while (lo < hi) { uint64_t m = (lo + hi) / 2; if (ix->deleted[m] < doc) lo = m + 1; else hi = m; }
    return lo < ix->n_deleted && ix->deleted[lo] == doc;
}

/* the shard's facet file (copied) */
int orc_index_set_facets(orc_index* ix, const void* rows, uint64_t first_doc_id, uint64_t n_docs, uint32_t row_bytes,
                         const orc_facet_field* fields, uint32_t n_fields) {
    if (!ix || n_fields > ORC_MAX_FACETS) return -1;
    free(ix->facet_rows); ix->facet_rows = NULL; ix->facet_n_docs = 0; ix->n_facets = 0;
    if (!n_docs || !n_fields) return 0;
    ix->facet_rows = (uint8_t*)malloc((size_t)n_docs * row_bytes); memcpy(ix->facet_rows, rows, (size_t)n_docs * row_bytes);
    ix->facet_n_docs = n_docs; ix->facet_first_doc = first_doc_id; ix->facet_row_bytes = row_bytes; ix->n_facets = n_fields;
    memcpy(ix->facet_fields, fields, n_fields * sizeof(orc_facet_field));
    return 0;
}

/* is_facet_filter (add_result.rs:340-478): 1 = the doc is filtered OUT.  One typed read + `!range.contains(&value)` (Rust Range<T>:
 * start <= x && x < end; floats by PartialOrd, so NaN is never contained) or `!values.contains(&id)` per filtered facet, each in the
 * facet's own C type — deliberately NOT the order-preserving-key trick of the product. */
#define ORC_RANGE_CASE(T, ST) { T x; ST a, b; memcpy(&x, p, sizeof(T)); memcpy(&a, &f->start, 8); memcpy(&b, &f->end, 8); if (!((ST)x >= a && (ST)x < b)) return 1; break; }
static int is_facet_filter(const orc_index* ix, const orc_facet_filter* fl, uint32_t n_fl, const uint64_t* set_values, uint64_t doc) {
    if (doc < ix->facet_first_doc || doc - ix->facet_first_doc >= ix->facet_n_docs) return 1;
    const uint8_t* row = ix->facet_rows + (size_t)(doc - ix->facet_first_doc) * ix->facet_row_bytes;
    for (uint32_t i = 0; i < n_fl; i++) {
        const orc_facet_filter* f = &fl[i];
        const uint8_t* p = row + ix->facet_fields[f->facet].offset;
        switch (ix->facet_fields[f->facet].type) {
            case ORC_FACET_U8:  ORC_RANGE_CASE(uint8_t, uint64_t)
            case ORC_FACET_U16: ORC_RANGE_CASE(uint16_t, uint64_t)
            case ORC_FACET_U32: ORC_RANGE_CASE(uint32_t, uint64_t)
            case ORC_FACET_U64: ORC_RANGE_CASE(uint64_t, uint64_t)
            case ORC_FACET_I8:  ORC_RANGE_CASE(int8_t, int64_t)
            case ORC_FACET_I16: ORC_RANGE_CASE(int16_t, int64_t)
            case ORC_FACET_I32: ORC_RANGE_CASE(int32_t, int64_t)
            case ORC_FACET_I64: case ORC_FACET_TIMESTAMP: ORC_RANGE_CASE(int64_t, int64_t)
            case ORC_FACET_F32: ORC_RANGE_CASE(float, double)     /* f32 -> f64 is exact and order-preserving; the bounds arrive as f64 */
            case ORC_FACET_F64: ORC_RANGE_CASE(double, double)
            case ORC_FACET_STRING16: case ORC_FACET_STRING32: {
                uint64_t id;
                if (ix->facet_fields[f->facet].type == ORC_FACET_STRING16) { uint16_t x; memcpy(&x, p, 2); id = x; } else { uint32_t x; memcpy(&x, p, 4); id = x; }
                int in = 0;
                for (uint32_t s = 0; s < f->set_count; s++) if (set_values[f->set_first + s] == id) in = 1;
                if (!in) return 1;
                break;
            }
            default: return 1;
        }
    }
    return 0;
}

static void* dup_mem(const void* p, size_t n) {
    void* q = malloc(n ? n : 1);
    if (n) memcpy(q, p, n);
    return q;
}

int orc_index_add_level(orc_index* ix, const orc_level* d) {
    if (!ix || !d || d->n_docs > 65536) return -1;
    if (ix->n_levels == ix->cap_levels) {
        ix->cap_levels = ix->cap_levels ? ix->cap_levels * 2 : 16;
        ix->levels = (lvl_t*)realloc(ix->levels, ix->cap_levels * sizeof(lvl_t));
    }
    lvl_t* l = &ix->levels[ix->n_levels++];
    memset(l, 0, sizeof(*l));
    l->level_id = d->level_id; l->n_docs = d->n_docs; l->n_terms = d->n_terms;
    uint32_t np = d->n_terms ? d->posting_offsets[d->n_terms] : 0;
    l->term_keys = (uint64_t*)dup_mem(d->term_keys, (size_t)d->n_terms * 8);
    l->posting_offsets = (uint32_t*)dup_mem(d->posting_offsets, ((size_t)d->n_terms + 1) * 4);
    l->doc_ids = (uint16_t*)dup_mem(d->doc_ids, (size_t)np * 2);
    l->tfs = (uint16_t*)dup_mem(d->tfs, (size_t)np * 2 * nf_of(ix));
    l->doc_len_bytes = (uint8_t*)dup_mem(d->doc_len_bytes, (size_t)d->n_docs * nf_of(ix));
    ix->committed = 0;
    return 0;
}

/* term positions of the level added last (single field): positions [sum of tfs], posting order, ascending inside a posting */
int orc_index_set_last_level_positions(orc_index* ix, const uint16_t* positions, uint64_t n_positions) {
    if (!ix || !ix->n_levels || nf_of(ix) != 1) return -1;
    lvl_t* l = &ix->levels[ix->n_levels - 1];
    uint32_t np = l->n_terms ? l->posting_offsets[l->n_terms] : 0;
    free(l->positions); free(l->pos_off);
    l->pos_off = (uint32_t*)malloc(((size_t)np + 1) * 4);
    uint64_t acc = 0;
    for (uint32_t j = 0; j < np; j++) { l->pos_off[j] = (uint32_t)acc; acc += l->tfs[j]; }
    l->pos_off[np] = (uint32_t)acc;
    if (acc != n_positions) { free(l->pos_off); l->pos_off = NULL; l->positions = NULL; return -2; }
    l->positions = (uint16_t*)dup_mem(positions, (size_t)n_positions * 2);
    return 0;
}

static int cmp_dict(const void* a, const void* b) {
    const dict_ent* x = (const dict_ent*)a; const dict_ent* y = (const dict_ent*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->level != y->level) return x->level < y->level ? -1 : 1;
    return 0;
}

/* query-independent part of the score of one posting: tf*(K+1)/(tf+cache[len]) */
static inline float comp_of(const orc_index* ix, uint32_t tf_u, uint8_t len_byte) {
    float tf = (float)tf_u;
    return tf * (ORC_K + 1.0f) / (tf + ix->cache[len_byte]);
}

int orc_index_commit(orc_index* ix, uint64_t n_docs, uint64_t len_sum) {
    if (!ix) return -1;
    ix->n_docs = n_docs; ix->len_sum = len_sum;
    orc_bm25_cache(n_docs, len_sum, ix->cache);
    uint64_t total = 0;
    for (uint32_t i = 0; i < ix->n_levels; i++) total += ix->levels[i].n_terms;
    free(ix->dict);
    ix->dict = (dict_ent*)malloc((total ? total : 1) * sizeof(dict_ent));
    uint64_t p = 0;
    for (uint32_t i = 0; i < ix->n_levels; i++) {
        lvl_t* l = &ix->levels[i];
        free(l->max_comp);
        l->max_comp = (float*)malloc(((size_t)l->n_terms ? l->n_terms : 1) * sizeof(float));
        for (uint32_t t = 0; t < l->n_terms; t++) {
            ix->dict[p].key = l->term_keys[t]; ix->dict[p].level = i; ix->dict[p].idx = t; p++;
            float m = 0.0f;
            if (nf_of(ix) == 1) for (uint32_t j = l->posting_offsets[t]; j < l->posting_offsets[t + 1]; j++) {
                float c = comp_of(ix, l->tfs[j], l->doc_len_bytes[l->doc_ids[j]]);
                if (c > m) m = c;
            }
            l->max_comp[t] = m;
        }
    }
    ix->n_dict = total;
    qsort(ix->dict, total, sizeof(dict_ent), cmp_dict);
    ix->committed = 1;
    return 0;
}

/* first dict entry with key >= key */
static uint64_t dict_lower(const orc_index* ix, uint64_t key) {
    uint64_t lo = 0, hi = ix->n_dict;
    while (lo < hi) { uint64_t m = (lo + hi) / 2; if (ix->dict[m].key < key) lo = m + 1; else hi = m; }
    return lo;
}

uint32_t orc_index_df(const orc_index* ix, uint64_t key) {
    uint64_t i = dict_lower(ix, key); uint64_t df = 0;
    for (; i < ix->n_dict && ix->dict[i].key == key; i++) {
        const lvl_t* l = &ix->levels[ix->dict[i].level];
        df += l->posting_offsets[ix->dict[i].idx + 1] - l->posting_offsets[ix->dict[i].idx];
    }
    return (uint32_t)df;
}

/* ------------------------------------------------------------------ canonical top-k */
/* canonical order: score desc, doc id asc.  better(a,b) = a ranks before b. */
static inline int better(float sa, uint64_t da, float sb, uint64_t db) {
    return (sa > sb) || (sa == sb && da < db);
}

typedef struct { orc_hit* h; uint32_t n, k; } topk_t;

/* simple insertion top-k kept sorted best-first */
static void topk_push(topk_t* t, uint64_t doc, float score) {
    if (t->k == 0) return;
    if (t->n == t->k && !better(score, doc, t->h[t->n - 1].score, t->h[t->n - 1].doc_id)) return;
    uint32_t i = t->n < t->k ? t->n++ : t->k - 1;
    while (i > 0 && better(score, doc, t->h[i - 1].score, t->h[i - 1].doc_id)) { t->h[i] = t->h[i - 1]; i--; }
    t->h[i].doc_id = doc; t->h[i].score = score; t->h[i].pad = 0;
}

/* ------------------------------------------------------------------ exhaustive lexical search */
typedef struct { uint64_t first, last; float idf; uint32_t df; } qterm_t;

#define ORC_MAX_TERMS 64

static int resolve_terms(const orc_index* ix, const uint64_t* keys, uint32_t n, qterm_t* qt) {
    for (uint32_t t = 0; t < n; t++) {
        uint64_t i = dict_lower(ix, keys[t]); uint64_t j = i; uint64_t df = 0;
        for (; j < ix->n_dict && ix->dict[j].key == keys[t]; j++) {
            const lvl_t* l = &ix->levels[ix->dict[j].level];
            df += l->posting_offsets[ix->dict[j].idx + 1] - l->posting_offsets[ix->dict[j].idx];
        }
        qt[t].first = i; qt[t].last = j; qt[t].df = (uint32_t)df;
        qt[t].idf = df ? orc_idf(ix->n_docs, (uint32_t)df) : 0.0f;
    }
    return 0;
}

/* dict entry of term t in level li, or -1 */
static int64_t term_in_level(const orc_index* ix, const qterm_t* q, uint32_t li) {
    uint64_t lo = q->first, hi = q->last;
    while (lo < hi) { uint64_t m = (lo + hi) / 2; if (ix->dict[m].level < li) lo = m + 1; else hi = m; }
    if (lo < q->last && ix->dict[lo].level == li) return (int64_t)lo;
    return -1;
}

int orc_search_lexical(const orc_index* ix, const uint64_t* keys, uint32_t n_terms, uint32_t query_type,
                       uint32_t k, uint32_t result_type, orc_hit* hits, uint32_t* n_hits,
                       uint64_t* count_total) {
    return orc_search_lexical_filtered(ix, keys, n_terms, NULL, 0, NULL, 0, NULL, query_type, k, result_type, hits, n_hits, count_total);
}

/* same with NOT terms ('-' operator, not_query_list add_result.rs:3440-3496): a doc that contains any of them is neither scored nor counted */
int orc_search_lexical_not(const orc_index* ix, const uint64_t* keys, uint32_t n_terms, const uint64_t* not_keys, uint32_t n_not,
                           uint32_t query_type, uint32_t k, uint32_t result_type, orc_hit* hits, uint32_t* n_hits,
                           uint64_t* count_total) {
    return orc_search_lexical_filtered(ix, keys, n_terms, not_keys, n_not, NULL, 0, NULL, query_type, k, result_type, hits, n_hits, count_total);
}

/* ... and with facet filters (facet_filter, add_result.rs:3498-3500: after the delete set and the NOT lists, before counting and scoring) */
int orc_search_lexical_filtered(const orc_index* ix, const uint64_t* keys, uint32_t n_terms, const uint64_t* not_keys, uint32_t n_not,
                                const orc_facet_filter* filters, uint32_t n_filters, const uint64_t* set_values,
                                uint32_t query_type, uint32_t k, uint32_t result_type, orc_hit* hits, uint32_t* n_hits,
                                uint64_t* count_total) {
    return orc_search_lexical_ex(ix, keys, n_terms, not_keys, n_not, filters, n_filters, set_values, 0, query_type, k, result_type, hits, n_hits, count_total);
}

/* ... and with a field filter (field_filter_set, add_result.rs:3124-3137 multi-field / 3558-3571 single-field): bit f of field_mask = indexed
 * field f is in the filter.  For every query term the doc contains: if (fields the term occurs in) + (fields of the filter) <= indexed fields
 * and none of the term's fields is in the filter, the doc is dropped — before it is counted or scored; the score still sums every field. */
int orc_search_lexical_ex(const orc_index* ix, const uint64_t* keys, uint32_t n_terms, const uint64_t* not_keys, uint32_t n_not,
                          const orc_facet_filter* filters, uint32_t n_filters, const uint64_t* set_values, uint32_t field_mask,
                          uint32_t query_type, uint32_t k, uint32_t result_type, orc_hit* hits, uint32_t* n_hits,
                          uint64_t* count_total) {
    if (!ix || !ix->committed || n_terms > ORC_MAX_TERMS || n_not > ORC_MAX_TERMS) return -1;
    if (n_filters && !ix->n_facets) return -1;
    for (uint32_t i = 0; i < n_filters; i++) if (filters[i].facet >= ix->n_facets) return -1;
    if (n_hits) *n_hits = 0;
    if (count_total) *count_total = 0;
    if (n_terms == 0) return 0;
    qterm_t qt[ORC_MAX_TERMS];
    resolve_terms(ix, keys, n_terms, qt);
    uint32_t n_live = 0; qterm_t live[ORC_MAX_TERMS];
    for (uint32_t t = 0; t < n_terms; t++) {
        if (qt[t].df == 0) {
            /* search.rs:3290-3296: AND with a missing term -> empty; OR drops the term */
            if (query_type == ORC_QUERY_INTERSECTION) return 0;
            continue;
        }
        live[n_live++] = qt[t];
    }
    if (n_live == 0) return 0;
    /* search.rs:2527-2531 heap capacity min(k, indexed_doc_count) */
    uint32_t kk = k; if ((uint64_t)kk > ix->n_docs) kk = (uint32_t)ix->n_docs;
    if (result_type == ORC_RESULT_COUNT) kk = 0;
    topk_t tk = { hits, 0, kk };
    float* acc = (float*)malloc(65536 * sizeof(float));
    uint8_t* cnt = (uint8_t*)malloc(65536);
    uint8_t* excl = (uint8_t*)malloc(65536);
    uint32_t n_filter_fields = 0;
    if (nf_of(ix) > 1) field_mask &= (1u << nf_of(ix)) - 1u; else field_mask = 0;   /* one indexed field: 1 + len <= 1 never holds */
    for (uint32_t f = 0; f < ORC_MAX_FIELDS; f++) n_filter_fields += (field_mask >> f) & 1u;
    qterm_t nq[ORC_MAX_TERMS];
    if (n_not) resolve_terms(ix, not_keys, n_not, nq);
    uint64_t total = 0;
    for (uint32_t li = 0; li < ix->n_levels; li++) {
        const lvl_t* l = &ix->levels[li];
        int any = 0;
        for (uint32_t t = 0; t < n_live; t++) if (term_in_level(ix, &live[t], li) >= 0) { any = 1; break; }
        if (!any) continue;
        memset(acc, 0, l->n_docs * sizeof(float));
        memset(cnt, 0, l->n_docs);
        memset(excl, 0, l->n_docs);
        for (uint32_t t = 0; t < n_not; t++) {
            int64_t e = nq[t].df ? term_in_level(ix, &nq[t], li) : -1;
            if (e < 0) continue;
            uint32_t ti = ix->dict[e].idx;
            for (uint32_t j = l->posting_offsets[ti]; j < l->posting_offsets[ti + 1]; j++) excl[l->doc_ids[j]] = 1;
        }
        for (uint32_t t = 0; t < n_live; t++) {   /* QUERY ORDER: bm25f += ... from 0.0 */
            int64_t e = term_in_level(ix, &live[t], li);
            if (e < 0) continue;
            uint32_t ti = ix->dict[e].idx;
            const uint32_t nf = nf_of(ix);
            for (uint32_t j = l->posting_offsets[ti]; j < l->posting_offsets[ti + 1]; j++) {
                uint16_t d = l->doc_ids[j];
                if (nf == 1) {
                    float comp = ix->cache[l->doc_len_bytes[d]];
                    acc[d] += orc_bm25_term(live[t].idf, l->tfs[j], comp);
                } else {
                    /* get_bm25f_multiterm_multifield, NgramType::SingleTerm (add_result.rs:1232-1262): for every field the term occurs in,
                     * bm25f += weight * idf * ((tf * (K + 1) / (tf + cache[len_byte of THAT field])) + SIGMA), fields in ascending order */
                    if (field_mask) {
                        uint32_t present = 0, np = 0;
                        for (uint32_t f = 0; f < nf; f++) if (l->tfs[(size_t)j * nf + f]) { present |= 1u << f; np++; }
                        if (np + n_filter_fields <= nf && !(present & field_mask)) excl[d] = 1;
                    }
                    for (uint32_t f = 0; f < nf; f++) {
                        uint32_t tfu = l->tfs[(size_t)j * nf + f];
                        if (!tfu) continue;
                        float tf = (float)tfu, comp = ix->cache[l->doc_len_bytes[(size_t)f * l->n_docs + d]];
                        acc[d] += ix->boost[f] * live[t].idf * ((tf * (ORC_K + 1.0f) / (tf + comp)) + ORC_SIGMA);
                    }
                }
                cnt[d]++;
            }
        }
        for (uint32_t d = 0; d < l->n_docs; d++) {
            int match = query_type == ORC_QUERY_INTERSECTION ? (cnt[d] == n_live) : (cnt[d] > 0);
            if (!match || excl[d]) continue;
            if (ix->n_deleted && is_deleted(ix, ((uint64_t)l->level_id << 16) | d)) continue;
            if (n_filters && is_facet_filter(ix, filters, n_filters, set_values, ((uint64_t)l->level_id << 16) | d)) continue;
            total++;
            if (kk) topk_push(&tk, ((uint64_t)l->level_id << 16) | d, acc[d]);
        }
    }
    free(acc); free(cnt); free(excl);
    if (n_hits) *n_hits = tk.n;
    if (count_total) *count_total = total;
    return 0;
}

/* ------------------------------------------------------------------ phrase search (QueryType::Phrase)
 * seq = the phrase's terms in order, repeats included (non_unique_query_list; term_index_nonunique = the index in seq).  A doc matches
 * iff it contains every term and some start position p has seq[i] at p + i for all i (add_result.rs:3586-3684: the k-way merge over the
 * position lists aligned by term_index_nonunique stops at the first common value, phrasematch_count >= 1).  Checked here the NAIVE way:
 * every position of seq[0] is tried as the start and each following token is looked up by binary search.  Score = BM25 of the unique
 * terms in first-occurrence order (query_list), count = matching docs; delete set honoured. */
static int has_position(const lvl_t* l, uint32_t j, uint32_t want) {
    uint32_t lo = l->pos_off[j], hi = l->pos_off[j + 1];
    while (lo < hi) { uint32_t m = (lo + hi) / 2; if (l->positions[m] < want) lo = m + 1; else hi = m; }
    return lo < l->pos_off[j + 1] && l->positions[lo] == want;
}
int orc_search_lexical_phrase(const orc_index* ix, const uint64_t* seq, uint32_t n_seq, uint32_t k, uint32_t result_type,
                              orc_hit* hits, uint32_t* n_hits, uint64_t* count_total) {
    if (!ix || !ix->committed || n_seq > ORC_MAX_TERMS || nf_of(ix) != 1) return -1;
    if (n_hits) *n_hits = 0;
    if (count_total) *count_total = 0;
    if (n_seq < 2) return orc_search_lexical(ix, seq, n_seq, ORC_QUERY_INTERSECTION, k, result_type, hits, n_hits, count_total);
    uint64_t ukeys[ORC_MAX_TERMS]; uint32_t uidx[ORC_MAX_TERMS], nu = 0;
    for (uint32_t i = 0; i < n_seq; i++) {
        uint32_t u = nu;
        for (uint32_t t = 0; t < nu; t++) if (ukeys[t] == seq[i]) u = t;
        if (u == nu) ukeys[nu++] = seq[i];
        uidx[i] = u;
    }
    qterm_t qt[ORC_MAX_TERMS];
    resolve_terms(ix, ukeys, nu, qt);
    for (uint32_t t = 0; t < nu; t++) if (qt[t].df == 0) return 0;
    uint32_t kk = k; if ((uint64_t)kk > ix->n_docs) kk = (uint32_t)ix->n_docs;
    if (result_type == ORC_RESULT_COUNT) kk = 0;
    topk_t tk = { hits, 0, kk };
    float* acc = (float*)malloc(65536 * sizeof(float));
    uint8_t* cnt = (uint8_t*)malloc(65536);
    uint32_t* pj = (uint32_t*)malloc((size_t)nu * 65536 * 4);
    uint64_t total = 0;
    for (uint32_t li = 0; li < ix->n_levels; li++) {
        const lvl_t* l = &ix->levels[li];
        int all = 1;
        for (uint32_t t = 0; t < nu; t++) if (term_in_level(ix, &qt[t], li) < 0) { all = 0; break; }
        if (!all) continue;
        if (!l->positions) { free(acc); free(cnt); free(pj); return -3; }
        memset(acc, 0, l->n_docs * sizeof(float)); memset(cnt, 0, l->n_docs);
        for (uint32_t t = 0; t < nu; t++) {
            uint32_t ti = ix->dict[term_in_level(ix, &qt[t], li)].idx;
            for (uint32_t j = l->posting_offsets[ti]; j < l->posting_offsets[ti + 1]; j++) {
                uint16_t d = l->doc_ids[j];
                acc[d] += orc_bm25_term(qt[t].idf, l->tfs[j], ix->cache[l->doc_len_bytes[d]]);
                cnt[d]++; pj[(size_t)t * 65536 + d] = j;
            }
        }
        for (uint32_t d = 0; d < l->n_docs; d++) {
            if (cnt[d] != nu) continue;
            if (ix->n_deleted && is_deleted(ix, ((uint64_t)l->level_id << 16) | d)) continue;
            uint32_t j0 = pj[(size_t)uidx[0] * 65536 + d];
            int match = 0;
            for (uint32_t a = l->pos_off[j0]; a < l->pos_off[j0 + 1] && !match; a++) {
                uint32_t p = l->positions[a]; int ok = 1;
                for (uint32_t i = 1; i < n_seq && ok; i++) ok = has_position(l, pj[(size_t)uidx[i] * 65536 + d], p + i);
                match = ok;
            }
            if (!match) continue;
            total++;
            if (kk) topk_push(&tk, ((uint64_t)l->level_id << 16) | d, acc[d]);
        }
    }
    free(acc); free(cnt); free(pj);
    if (n_hits) *n_hits = tk.n;
    if (count_total) *count_total = total;
    return 0;
}

/* ------------------------------------------------------------------ reference-shaped pruned search
 * A faithful CPU restatement of the reference control flow, used (a) as the timed CPU baseline and
 * (b) as a cross-check that pruning == exhaustive.  Heap: min_heap.rs (binary min-heap on score,
 * strict > replacement, doc-id dedup map for OR).  Here the heap keeps canonical (score, docid) order so
 * that its result equals the exhaustive canonical top-k even inside tie groups. */
typedef struct { float s; uint64_t d; } hent;
typedef struct { hent* e; uint32_t n, k; } heap_t;

static inline int worse(const hent* a, const hent* b) { /* a ranks after b */
    return better(b->s, b->d, a->s, a->d);
}
static void heap_sift_down(heap_t* h, uint32_t i) {
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < h->n && worse(&h->e[l], &h->e[m])) m = l;
        if (r < h->n && worse(&h->e[r], &h->e[m])) m = r;
        if (m == i) return;
        hent t = h->e[i]; h->e[i] = h->e[m]; h->e[m] = t; i = m;
    }
}
static void heap_sift_up(heap_t* h, uint32_t i) {
    while (i > 0) {
        uint32_t p = (i - 1) / 2;
        if (!worse(&h->e[i], &h->e[p])) return;
        hent t = h->e[i]; h->e[i] = h->e[p]; h->e[p] = t; i = p;
    }
}
/* min_heap.rs:1193-1259 add_topk, with doc-id dedup (linear over k entries; k is small) */
static void heap_add(heap_t* h, uint64_t d, float s, int dedup) {
    if (h->k == 0) return;
    hent c = { s, d };
    if (dedup) {
        for (uint32_t i = 0; i < h->n; i++) if (h->e[i].d == d) {
            if (s > h->e[i].s) { h->e[i].s = s; heap_sift_down(h, i); heap_sift_up(h, i); }
            return;
        }
    }
    if (h->n < h->k) { h->e[h->n++] = c; heap_sift_up(h, h->n - 1); return; }
    if (worse(&h->e[0], &c)) { h->e[0] = c; heap_sift_down(h, 0); }
}
static inline float heap_min(const heap_t* h) { return h->e[0].s; }
static inline int heap_full(const heap_t* h) { return h->n == h->k; }

typedef struct { uint32_t level; uint32_t tidx[ORC_MAX_TERMS]; float bound; } blk_t;

static int cmp_blk_desc(const void* a, const void* b) {
    float x = ((const blk_t*)a)->bound, y = ((const blk_t*)b)->bound;
    if (x != y) return x > y ? -1 : 1;
    uint32_t la = ((const blk_t*)a)->level, lb = ((const blk_t*)b)->level;
    return la < lb ? -1 : (la > lb);
}

/* galloping lower bound in a sorted u16 array starting from *pos (intersection.rs:352-362) */
static inline uint32_t gallop(const uint16_t* a, uint32_t pos, uint32_t n, uint16_t x) {
    uint32_t step = 1, lo = pos;
    while (lo + step < n && a[lo + step] < x) { lo += step; step <<= 1; }
    uint32_t hi = lo + step < n ? lo + step : n;
    while (lo < hi) { uint32_t m = (lo + hi) / 2; if (a[m] < x) lo = m + 1; else hi = m; }
    return lo;
}

/* AND over the given term subset (intersection_blockid + intersection_docid); returns match count */
static uint64_t and_pass(const orc_index* ix, const qterm_t* terms, const uint32_t* order /*query order idx*/,
                         uint32_t nt, heap_t* h, int prune_blocks, int dedup) {
    /* block list = levels where every term occurs; bound = Σ idf*max_comp (intersection.rs:2090-2109) */
    blk_t* blocks = (blk_t*)malloc((ix->n_levels ? ix->n_levels : 1) * sizeof(blk_t));
    uint32_t nb = 0;
    for (uint32_t li = 0; li < ix->n_levels; li++) {
        blk_t b; b.level = li; b.bound = 0.0f; int ok = 1;
        for (uint32_t t = 0; t < nt; t++) {
            int64_t e = term_in_level(ix, &terms[t], li);
            if (e < 0) { ok = 0; break; }
            b.tidx[t] = ix->dict[e].idx;
            b.bound += terms[t].idf * ix->levels[li].max_comp[b.tidx[t]];
        }
        if (ok) blocks[nb++] = b;
    }
    qsort(blocks, nb, sizeof(blk_t), cmp_blk_desc); /* intersection.rs:2225 */
    uint64_t count = 0;
    for (uint32_t bi = 0; bi < nb; bi++) {
        const blk_t* b = &blocks[bi];
        int score_block = 1;
        /* canonical tie rule: an equal-score doc with a smaller id could still displace the heap root, so
         * only strict < prunes; the reference uses <= (intersection.rs:2227-2233, add_result.rs:3512-3536) */
        if (h->k > 0 && heap_full(h) && b->bound < heap_min(h)) { if (prune_blocks) break; score_block = 0; }
        if (h->k == 0) score_block = 0;
        const lvl_t* l = &ix->levels[b->level];
        /* drive with the shortest list (intersection.rs:258-273) */
        uint32_t drv = 0, dn = 0xffffffffu;
        for (uint32_t t = 0; t < nt; t++) {
            uint32_t c = l->posting_offsets[b->tidx[t] + 1] - l->posting_offsets[b->tidx[t]];
            if (c < dn) { dn = c; drv = t; }
        }
        uint32_t pos[ORC_MAX_TERMS]; memset(pos, 0, sizeof(pos));
        const uint16_t* da = l->doc_ids + l->posting_offsets[b->tidx[drv]];
        for (uint32_t i = 0; i < dn; i++) {
            uint16_t d = da[i]; int all = 1; uint32_t tfv[ORC_MAX_TERMS];
            tfv[drv] = l->tfs[l->posting_offsets[b->tidx[drv]] + i];
            for (uint32_t t = 0; t < nt && all; t++) {
                if (t == drv) continue;
                uint32_t off = l->posting_offsets[b->tidx[t]];
                uint32_t n = l->posting_offsets[b->tidx[t] + 1] - off;
                uint32_t p = gallop(l->doc_ids + off, pos[t], n, d);
                pos[t] = p;
                if (p >= n || l->doc_ids[off + p] != d) all = 0; else tfv[t] = l->tfs[off + p];
            }
            if (!all) continue;
            count++;
            if (!score_block) continue;
            float comp = ix->cache[l->doc_len_bytes[d]];
            /* sum in QUERY ORDER (order[] maps sorted position -> term), from 0.0 */
            float s = 0.0f;
            for (uint32_t t = 0; t < nt; t++) s += orc_bm25_term(terms[order[t]].idf, tfv[order[t]], comp);
            heap_add(h, ((uint64_t)l->level_id << 16) | d, s, dedup);
        }
    }
    free(blocks);
    return count;
}

/* single_blockid + single_docid (single.rs:292-417): blocks by bound desc, stop at first block that
 * cannot beat heap.min; at most k blocks are needed when unfiltered (:375). */
static void single_pass(const orc_index* ix, const qterm_t* term, heap_t* h, int dedup) {
    uint64_t nb = term->last - term->first;
    blk_t* blocks = (blk_t*)malloc((nb ? nb : 1) * sizeof(blk_t));
    for (uint64_t i = 0; i < nb; i++) {
        const dict_ent* e = &ix->dict[term->first + i];
        blocks[i].level = e->level; blocks[i].tidx[0] = e->idx;
        blocks[i].bound = term->idf * ix->levels[e->level].max_comp[e->idx];
    }
    qsort(blocks, nb, sizeof(blk_t), cmp_blk_desc);
    for (uint64_t bi = 0; bi < nb; bi++) {
        if (heap_full(h) && blocks[bi].bound < heap_min(h)) break;
        const lvl_t* l = &ix->levels[blocks[bi].level];
        uint32_t off = l->posting_offsets[blocks[bi].tidx[0]], end = l->posting_offsets[blocks[bi].tidx[0] + 1];
        for (uint32_t j = off; j < end; j++) {
            uint16_t d = l->doc_ids[j];
            float s = 0.0f; s += orc_bm25_term(term->idf, l->tfs[j], ix->cache[l->doc_len_bytes[d]]);
            heap_add(h, ((uint64_t)l->level_id << 16) | d, s, dedup);
        }
    }
    free(blocks);
}

/* exact |union| per level via 64K-bit bitmap OR + popcount (union.rs:807-1164 union_count) */
static uint64_t union_count_all(const orc_index* ix, const qterm_t* terms, uint32_t nt) {
    uint64_t total = 0; uint64_t* bm = (uint64_t*)malloc(1024 * 8);
    for (uint32_t li = 0; li < ix->n_levels; li++) {
        const lvl_t* l = &ix->levels[li]; int any = 0;
        for (uint32_t t = 0; t < nt; t++) {
            int64_t e = term_in_level(ix, &terms[t], li);
            if (e < 0) continue;
            if (!any) { memset(bm, 0, 8192); any = 1; }
            uint32_t ti = ix->dict[e].idx;
            for (uint32_t j = l->posting_offsets[ti]; j < l->posting_offsets[ti + 1]; j++)
                bm[l->doc_ids[j] >> 6] |= 1ull << (l->doc_ids[j] & 63);
        }
        if (any) for (int w = 0; w < 1024; w++) total += (uint64_t)__builtin_popcountll(bm[w]);
    }
    free(bm);
    return total;
}

/* OR over 3+ terms: block-max ordered levels + MAXSCORE inside a level — the work union_docid_3's sub-query queue does
 * (union.rs:1308-1479: sub-queries ordered by their summed max scores, each skipped once it cannot beat heap.min), restated
 * as the textbook document-at-a-time loop so that the timed CPU baseline is not a strawman: levels by Σ block-max bound
 * descending, stop at the first level whose bound < heap.min (only strictly smaller bounds prune, intersection.rs:2227-2233);
 * inside a level the lists are driven in block-max order while the in-query-order sum of the not-yet-driven bounds can
 * reach heap.min; a posting is scored fully (galloping probes into the other lists, monotone cursors) unless its own
 * contribution + the later bounds cannot reach heap.min; docs found in an earlier-driven list are duplicates. */
typedef struct { const orc_index* ix; const qterm_t* all; uint32_t n_all; heap_t* h; float max_list[ORC_MAX_TERMS]; } orctx;
typedef struct { uint32_t level; float bound; } lvb_t;
static int cmp_lvb_desc(const void* a, const void* b) {
    float x = ((const lvb_t*)a)->bound, y = ((const lvb_t*)b)->bound;
    if (x != y) return x > y ? -1 : 1;
    uint32_t la = ((const lvb_t*)a)->level, lb = ((const lvb_t*)b)->level;
    return la < lb ? -1 : (la > lb);
}

static void or_maxscore(orctx* c) {
    const orc_index* ix = c->ix; const uint32_t n = c->n_all; heap_t* h = c->h;
    float* bound = (float*)calloc(ix->n_levels ? ix->n_levels : 1, sizeof(float));
    uint8_t* seen = (uint8_t*)calloc(ix->n_levels ? ix->n_levels : 1, 1);
    for (uint32_t t = 0; t < n; t++)                     /* query order: the bound is summed like a score */
        for (uint64_t i = c->all[t].first; i < c->all[t].last; i++) {
            uint32_t lv = ix->dict[i].level;
            bound[lv] += c->all[t].idf * ix->levels[lv].max_comp[ix->dict[i].idx]; seen[lv] = 1;
        }
    lvb_t* order = (lvb_t*)malloc((ix->n_levels ? ix->n_levels : 1) * sizeof(lvb_t)); uint32_t nl = 0;
    for (uint32_t lv = 0; lv < ix->n_levels; lv++) if (seen[lv]) { order[nl].level = lv; order[nl].bound = bound[lv]; nl++; }
    qsort(order, nl, sizeof(lvb_t), cmp_lvb_desc);
    for (uint32_t oi = 0; oi < nl; oi++) {
        if (heap_full(h) && order[oi].bound < heap_min(h)) break;
        const uint32_t li = order[oi].level; const lvl_t* l = &ix->levels[li];
        uint32_t off[ORC_MAX_TERMS], cnt[ORC_MAX_TERMS], rank[ORC_MAX_TERMS], pos[ORC_MAX_TERMS]; float ub[ORC_MAX_TERMS];
        for (uint32_t t = 0; t < n; t++) {
            int64_t e = term_in_level(ix, &c->all[t], li);
            if (e < 0) { cnt[t] = 0; off[t] = 0; ub[t] = 0.0f; continue; }
            uint32_t ti = ix->dict[e].idx;
            off[t] = l->posting_offsets[ti]; cnt[t] = l->posting_offsets[ti + 1] - off[t];
            ub[t] = c->all[t].idf * l->max_comp[ti];
        }
        for (uint32_t t = 0; t < n; t++) {                /* MAXSCORE order: present lists by bound desc */
            uint32_t r = 0;
            for (uint32_t u = 0; u < n; u++) if (u != t && cnt[u] && (ub[u] > ub[t] || (ub[u] == ub[t] && u < t))) r++;
            rank[t] = cnt[t] ? r : 0xFFFFu;
        }
        uint32_t np = 0; for (uint32_t t = 0; t < n; t++) np += cnt[t] ? 1u : 0u;
        for (uint32_t p = 0; p < np; p++) {
            float S = 0.0f, R = 0.0f; uint32_t drv = 0;
            for (uint32_t t = 0; t < n; t++) { if (!cnt[t]) continue; if (rank[t] >= p) S += ub[t]; if (rank[t] > p) R += ub[t]; if (rank[t] == p) drv = t; }
            if (heap_full(h) && S < heap_min(h)) break;
            memset(pos, 0, sizeof(uint32_t) * n);
            const float didf = c->all[drv].idf;
            for (uint32_t j = 0; j < cnt[drv]; j++) {
                const uint16_t d = l->doc_ids[off[drv] + j];
                const float comp = ix->cache[l->doc_len_bytes[d]];
                const float cd = orc_bm25_term(didf, l->tfs[off[drv] + j], comp);
                if (heap_full(h) && (cd + R) * 1.000002f < heap_min(h)) continue;
                float sc = 0.0f; int dup = 0;
                for (uint32_t t = 0; t < n && !dup; t++) {   /* full score, query order */
                    if (t == drv) { sc += cd; continue; }
                    if (!cnt[t]) continue;
                    uint32_t q = gallop(l->doc_ids + off[t], pos[t], cnt[t], d); pos[t] = q;
                    if (q < cnt[t] && l->doc_ids[off[t] + q] == d) {
                        if (rank[t] < p) dup = 1;            /* already emitted when that list was the driver */
                        else sc += orc_bm25_term(c->all[t].idf, l->tfs[off[t] + q], comp);
                    }
                }
                if (!dup) heap_add(h, ((uint64_t)l->level_id << 16) | d, sc, 0);
            }
        }
    }
    free(order); free(seen); free(bound);
}

int orc_search_lexical_pruned(const orc_index* ix, const uint64_t* keys, uint32_t n_terms, uint32_t query_type,
                              uint32_t k, uint32_t result_type, orc_hit* hits, uint32_t* n_hits,
                              uint64_t* count_total) {
    if (ix && ix->n_fields > 1) return -2;   /* the reference-shaped pruned walk is restated for one indexed field only */
    if (!ix || !ix->committed || n_terms > ORC_MAX_TERMS) return -1;
    if (n_hits) *n_hits = 0;
    if (count_total) *count_total = 0;
    if (n_terms == 0) return 0;
    qterm_t qt[ORC_MAX_TERMS]; resolve_terms(ix, keys, n_terms, qt);
    qterm_t live[ORC_MAX_TERMS]; uint32_t nl = 0;
    for (uint32_t t = 0; t < n_terms; t++) {
        if (qt[t].df == 0) { if (query_type == ORC_QUERY_INTERSECTION) return 0; continue; }
        live[nl++] = qt[t];
    }
    if (nl == 0) return 0;
    uint32_t kk = k; if ((uint64_t)kk > ix->n_docs) kk = (uint32_t)ix->n_docs;
    if (result_type == ORC_RESULT_COUNT) kk = 0;
    heap_t h; h.e = (hent*)malloc(((size_t)kk ? kk : 1) * sizeof(hent)); h.n = 0; h.k = kk;
    uint32_t order[ORC_MAX_TERMS]; for (uint32_t t = 0; t < nl; t++) order[t] = t;
    uint64_t total = 0;
    if (query_type == ORC_QUERY_INTERSECTION || nl == 1) {
        if (nl == 1 && query_type != ORC_QUERY_INTERSECTION) {
            total = live[0].df;                       /* single.rs:314-322 */
            if (kk) single_pass(ix, &live[0], &h, 0);
        } else {
            total = and_pass(ix, live, order, nl, &h, result_type == ORC_RESULT_TOPK, 0);
        }
    } else if (nl == 2) {
        /* union_docid_2 (union.rs:1168-1304) */
        uint64_t both = and_pass(ix, live, order, 2, &h, result_type == ORC_RESULT_TOPK, 0);
        total = (uint64_t)live[0].df + live[1].df;
        total -= both; /* inaccurate under Topk, as in the reference (search.rs:196-198) */
        if (kk) for (uint32_t t = 0; t < 2; t++) {
            float mx = 0.0f;
            for (uint64_t i = live[t].first; i < live[t].last; i++) {
                float b = live[t].idf * ix->levels[ix->dict[i].level].max_comp[ix->dict[i].idx];
                if (b > mx) mx = b;
            }
            if (!heap_full(&h) || mx >= heap_min(&h)) single_pass(ix, &live[t], &h, 1);
        }
    } else {
        orctx c; c.ix = ix; c.all = live; c.n_all = nl; c.h = &h;
        for (uint32_t t = 0; t < nl; t++) { float mx = 0.0f;
            for (uint64_t i = live[t].first; i < live[t].last; i++) {
                float b = live[t].idf * ix->levels[ix->dict[i].level].max_comp[ix->dict[i].idx];
                if (b > mx) mx = b; }
            c.max_list[t] = mx; }
        if (kk) or_maxscore(&c);
        if (result_type != ORC_RESULT_TOPK) total = union_count_all(ix, live, nl);
    }
    /* search.rs:3565-3596: take heap, sort by score desc (canonical: then doc id asc) */
    topk_t tk = { hits, 0, kk };
    for (uint32_t i = 0; i < h.n; i++) topk_push(&tk, h.e[i].d, h.e[i].s);
    free(h.e);
    if (n_hits) *n_hits = tk.n;
    if (count_total) *count_total = total;
    return 0;
}

/* ------------------------------------------------------------------ vectors */
void orc_normalize_f32(float* v, uint32_t n) {
    /* vector_similarity.rs:70-74 */
    float s = 0.0f; for (uint32_t i = 0; i < n; i++) s += v[i] * v[i];
    float f = 1.0f / sqrtf(s);
    for (uint32_t i = 0; i < n; i++) v[i] *= f;
}

float orc_dot_f32(const float* a, const float* b, uint32_t n) {
    /* vector_similarity.rs:1006-1008: left-to-right sum of products */
    float s = 0.0f; for (uint32_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

float orc_dot_f32_lanes8(const float* q, const float* e, uint32_t n) {
    /* vector_similarity.rs:1120-1142: 8 lanes of fused multiply-add, then in-order sum of the lanes */
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t i = 0;
    for (; i + 8 <= n; i += 8)
        for (int j = 0; j < 8; j++) lane[j] = fmaf(q[i + j], e[i + j], lane[j]);
    float s = 0.0f; for (int j = 0; j < 8; j++) s += lane[j];
    for (; i < n; i++) s += q[i] * e[i];
    return s;
}

float orc_euclidean_f32(const float* a, const float* b, uint32_t n) {
    /* vector_similarity.rs:912-918 */
    float s = 0.0f; for (uint32_t i = 0; i < n; i++) { float d = a[i] - b[i]; s += d * d; }
    return s;
}

float orc_vector_score_postmap(float score, uint32_t sim) {
    /* vector.rs:1489-1499, SIMILARITY_NORMALIZATION_64_I8 = 1/16129 (vector.rs:29) */
    if (sim == ORC_SIM_EUCLIDEAN) return -score;
    return ((score * (1.0f / 16129.0f)) + 1.0f) * 0.5f;
}

typedef struct {
    const float* rows; const uint32_t* ids; uint64_t lo, hi; uint32_t dims, pitch, sim, lanes8;
    const float* q; topk_t tk;
} vjob;

static void* vscan(void* p) {
    vjob* j = (vjob*)p;
    for (uint64_t r = j->lo; r < j->hi; r++) {
        const float* e = j->rows + r * j->pitch; float s;
        if (j->sim == ORC_SIM_EUCLIDEAN) s = -orc_euclidean_f32(j->q, e, j->dims);
        else s = j->lanes8 ? orc_dot_f32_lanes8(j->q, e, j->dims) : orc_dot_f32(j->q, e, j->dims);
        topk_push(&j->tk, j->ids ? j->ids[r] : r, s);
    }
    return NULL;
}

int orc_search_vector(const float* rows, const uint32_t* ids, uint64_t n_rows, uint32_t dims, uint32_t pitch,
                      const float* query, uint32_t sim, uint32_t k, uint32_t lanes8, uint32_t n_threads,
                      orc_hit* hits, uint32_t* n_hits) {
    if (!rows || !query || !hits) return -1;
    if (pitch == 0) pitch = dims;
    if (n_threads == 0) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    vjob* jobs = (vjob*)calloc(n_threads, sizeof(vjob));
    pthread_t* th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
    for (uint32_t t = 0; t < n_threads; t++) {
        vjob* j = &jobs[t];
        j->rows = rows; j->ids = ids; j->dims = dims; j->pitch = pitch; j->sim = sim; j->lanes8 = lanes8; j->q = query;
        j->lo = n_rows * t / n_threads; j->hi = n_rows * (t + 1) / n_threads;
        j->tk.h = (orc_hit*)malloc(((size_t)k ? k : 1) * sizeof(orc_hit)); j->tk.n = 0; j->tk.k = k;
        if (n_threads > 1) pthread_create(&th[t], NULL, vscan, j); else vscan(j);
    }
    topk_t out = { hits, 0, k };
    for (uint32_t t = 0; t < n_threads; t++) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        for (uint32_t i = 0; i < jobs[t].tk.n; i++) topk_push(&out, jobs[t].tk.h[i].doc_id, jobs[t].tk.h[i].score);
        free(jobs[t].tk.h);
    }
    free(jobs); free(th);
    if (n_hits) *n_hits = out.n;
    return 0;
}

/* ------------------------------------------------------------------ int8 scalar quantisation */
void orc_quantize_f32_to_i8(const float* v, uint32_t n, int8_t* out) {
    /* vector_similarity.rs:1226-1232 */
    for (uint32_t i = 0; i < n; i++) {
        float r = roundf(v[i] * 127.0f);
        if (r < -127.0f) r = -127.0f;
        if (r > 127.0f) r = 127.0f;
        out[i] = r == r ? (int8_t)r : (int8_t)0;   /* NaN as i8 = 0 (Rust saturating cast) */
    }
}

void orc_quantize_rows_i8(const float* rows, uint64_t n_rows, uint32_t dims, uint64_t pitch, int8_t* out, uint64_t out_pitch) {
    float* tmp = (float*)malloc((size_t)dims * sizeof(float));
    for (uint64_t r = 0; r < n_rows; r++) {
        memcpy(tmp, rows + r * pitch, (size_t)dims * sizeof(float));
        orc_normalize_f32(tmp, dims);
        orc_quantize_f32_to_i8(tmp, dims, out + r * out_pitch);
    }
    free(tmp);
}

int32_t orc_dot_i8(const int8_t* a, const int8_t* b, uint32_t n) {
    /* vector_similarity.rs:1011-1016 */
    int32_t s = 0;
    for (uint32_t i = 0; i < n; i++) s += (int32_t)a[i] * (int32_t)b[i];
    return s;
}

int orc_search_vector_i8(const int8_t* rows, const uint32_t* ids, uint64_t n_rows, uint32_t dims, uint32_t pitch,
                         const int8_t* query, uint32_t k, orc_hit* hits, uint32_t* n_hits) {
    if (!rows || !query || !hits) return -1;
    if (pitch == 0) pitch = dims;
    topk_t tk = { hits, 0, k };
    for (uint64_t r = 0; r < n_rows; r++)
        topk_push(&tk, ids ? ids[r] : r, (float)orc_dot_i8(query, rows + r * pitch, dims));   /* dot_i8(a, b) as f32 */
    if (n_hits) *n_hits = tk.n;
    return 0;
}

/* ------------------------------------------------------------------ RRF: search.rs:1962-2035 */
int orc_rrf(const orc_hit* lex, uint32_t n_lex, const orc_hit* vec, uint32_t n_vec, orc_hit* out, uint32_t* n_out) {
    const float kf = 0.6f;
    uint32_t n = 0;
    /* inputs are already sorted score desc (canonical) — the reference re-sorts them (:1970, :1989) */
    for (uint32_t i = 0; i < n_lex; i++) {
        uint32_t j = 0; for (; j < n; j++) if (out[j].doc_id == lex[i].doc_id) break;
        float s = 1.0f / (kf + (float)i);
        if (j == n) { out[n].doc_id = lex[i].doc_id; out[n].score = s; out[n].pad = 0; n++; }
        else out[j].score = s; /* HashMap::insert overwrites (:1975) */
    }
    for (uint32_t i = 0; i < n_vec; i++) {
        uint32_t j = 0; for (; j < n; j++) if (out[j].doc_id == vec[i].doc_id) break;
        float s = 1.0f / (kf + (float)i);
        if (j == n) { out[n].doc_id = vec[i].doc_id; out[n].score = s; out[n].pad = 0; n++; }
        else out[j].score += s;
    }
    /* :2097-2106 sort score desc; canonical tie: doc id asc */
    for (uint32_t i = 1; i < n; i++) { orc_hit x = out[i]; uint32_t j = i;
        while (j > 0 && better(x.score, x.doc_id, out[j-1].score, out[j-1].doc_id)) { out[j] = out[j-1]; j--; }
        out[j] = x; }
    if (n_out) *n_out = n;
    return 0;
}


/* ---- Dot / Euclidean + ScalarQuantizationI8 (vector.rs:597-660, search.rs:1499-1530) ----
 * QuantizedVector::new_scale / new_scale_norm (vector_similarity.rs:1340-1371): scale = max|x| / 127, codes = (x / scale).round() as i8
 * (Rust `as i8` saturates, NaN -> 0), norm = sum(code^2) as f32 * scale * scale. */
void orc_quantize_scale_i8(const float* v, uint32_t n, int want_norm, int8_t* out, float* scale_out, float* norm_out) {
    float mx = 0.0f;
    for (uint32_t i = 0; i < n; i++) { float a = fabsf(v[i]); if (a > mx) mx = a; }   /* fold(0.0, f32::max) */
    volatile float scale = mx / 127.0f;
    int32_t sum = 0;
    for (uint32_t i = 0; i < n; i++) {
        volatile float q = v[i] / scale;
        float r = roundf(q);
        int8_t c = 0;
        if (r == r) { if (r > 127.0f) r = 127.0f; if (r < -128.0f) r = -128.0f; c = (int8_t)r; }
        out[i] = c; sum += (int32_t)c * (int32_t)c;
    }
    *scale_out = scale;
    if (norm_out) { volatile float a = (float)sum * scale; volatile float b = a * scale; *norm_out = want_norm ? b : 0.0f; }
}
/* dot_i8_quantized (vector_similarity.rs:1754-1758) / -euclidean_i8_quantized (:1721-1734), query = v1 */
float orc_score_i8_scaled(const int8_t* q, float q_scale, float q_norm, const int8_t* e, float e_scale, float e_norm, uint32_t n, uint32_t similarity) {
    int32_t dot = orc_dot_i8(q, e, n);
    volatile float a = (float)dot * q_scale; volatile float d = a * e_scale;
    if (similarity != ORC_SIM_EUCLIDEAN) return d;
    volatile float s = q_norm + e_norm; volatile float t = 2.0f * d; volatile float r = s - t;
    return -(r > 0.0f ? r : 0.0f);
}
/* ---- affine Euclidean SQ (integer-valued 0..255 data): QuantizedVector::new_scale_norm_affine (vector_similarity.rs:1414-1463), raster_range
 * (:1465-1472), euclidean_i8_quantized_affine (:1770-1795).  min_state / max_state = shard.min_vector_value / max_vector_value (initially
 * f32::MAX / f32::MIN), updated in place exactly like the reference does (note its quirk: when a vector raises the maximum, the state
 * receives the RASTERED RANGE, not the maximum). */
static float orc_raster_range(float range) {
    if (!(range > 1.0f)) return range;
    uint64_t v = (uint64_t)(int64_t)range + 1u, p = 1;
    while (p < v) p <<= 1;
    return (float)(p - 1);
}
void orc_quantize_affine_i8(const float* v, uint32_t n, float* min_state, float* max_state, int8_t* out, float* scale_out, float* norm_out,
                            int32_t* zero_point_out, int32_t* sum_q_out) {
    float mn = INFINITY, mx = -INFINITY;
    for (uint32_t i = 0; i < n; i++) { mn = fminf(mn, v[i]); mx = fmaxf(mx, v[i]); }
    if (mn < *min_state) *min_state = mn; else mn = *min_state;
    if (mx > *max_state) { volatile float d = mx - mn; *max_state = orc_raster_range(d); } else mx = *max_state;
    volatile float d = mx - mn;
    volatile float range = orc_raster_range(d);
    volatile float scale = range / 255.0f;
    volatile float q = mn / scale;
    volatile float zf = -128.0f - q;
    float z = roundf(zf);
    if (z < -128.0f) z = -128.0f;
    if (z > 127.0f) z = 127.0f;
    int32_t zp = z == z ? (int32_t)z : 0;
    int32_t sq = 0, sum = 0;
    for (uint32_t i = 0; i < n; i++) {
        volatile float t = v[i] / scale;
        float r = roundf(t);
        int64_t xi = r != r ? 0 : (r >= 2147483648.0f ? 2147483647LL : (r <= -2147483648.0f ? -2147483648LL : (int64_t)r));
        int64_t s = xi + zp;
        int8_t c = (int8_t)(s < -128 ? -128 : (s > 127 ? 127 : s));
        out[i] = c; sq += (int32_t)c * c; sum += c;
    }
    int32_t norm_i = sq - 2 * zp * sum + (int32_t)n * zp * zp;
    *scale_out = scale; *zero_point_out = zp; *sum_q_out = sum;
    { volatile float a = (float)norm_i * scale; volatile float b = a * scale; *norm_out = b; }
}
/* -euclidean_i8_quantized_affine, query = v1 */
float orc_score_i8_affine(const int8_t* q, float q_scale, float q_norm, int32_t q_zp, int32_t q_sum, const int8_t* e, float e_scale, float e_norm,
                          int32_t e_zp, int32_t e_sum, uint32_t n) {
    int32_t dot = orc_dot_i8(q, e, n);
    dot = dot - e_zp * q_sum - q_zp * e_sum + (int32_t)n * q_zp * e_zp;
    volatile float a = (float)dot * q_scale; volatile float d = a * e_scale;
    volatile float s = q_norm + e_norm; volatile float t = 2.0f * d; volatile float r = s - t;
    return -(r > 0.0f ? r : 0.0f);
}
int orc_search_vector_i8_affine(const int8_t* rows, const float* row_scale, const float* row_norm, const int32_t* row_zp, const int32_t* row_sum,
                                const uint32_t* doc_ids, uint64_t n_rows, uint32_t dims, uint32_t row_pitch, const int8_t* query, float q_scale, float q_norm,
                                int32_t q_zp, int32_t q_sum, uint32_t k, orc_hit* hits, uint32_t* n_hits) {
    topk_t tk = { hits, 0, k };
    for (uint64_t r = 0; r < n_rows; r++) {
        float s = orc_score_i8_affine(query, q_scale, q_norm, q_zp, q_sum, rows + r * row_pitch, row_scale[r], row_norm[r], row_zp[r], row_sum[r], dims);
        topk_push(&tk, doc_ids ? doc_ids[r] : r, s);
    }
    if (n_hits) *n_hits = tk.n;
    return 0;
}

/* ---- TurboQuantI8 (vector_similarity.rs:1825-2093): the vector is zero-padded to the next power of two, sign-flipped by the index's seed
 * mask (+-1, drawn once from ChaCha8Rng(seed 1234) — a third-party generator, so the mask is an INPUT here), rotated by the normalised
 * fast Walsh-Hadamard transform and quantised with scale = max(sigma / 32, 1e-8), sigma = ||x|| / sqrt(dim).  Scalar variant
 * (quantize_f32_i8 :1929-1958, fwht :1861-1880, calculate_scale :2035-2039): every sum left to right. */
void orc_turboquant_i8(const float* v, uint32_t n, uint32_t dim, const float* seed_mask, int8_t* out, float* scale_out, float* norm_out) {
    float* a = (float*)calloc(dim, sizeof(float));
    for (uint32_t i = 0; i < dim && i < n; i++) a[i] = v[i];
    for (uint32_t i = 0; i < dim; i++) { volatile float p = a[i] * seed_mask[i]; a[i] = p; }
    for (uint32_t h = 1; h < dim; h *= 2)
        for (uint32_t i = 0; i < dim; i += 2 * h)
            for (uint32_t j = i; j < i + h; j++) { volatile float x = a[j], y = a[j + h]; volatile float s = x + y, d = x - y; a[j] = s; a[j + h] = d; }
    volatile float nrm = sqrtf((float)dim);
    for (uint32_t i = 0; i < dim; i++) { volatile float q = a[i] / nrm; a[i] = q; }
    volatile float ss = 0.0f;
    for (uint32_t i = 0; i < dim; i++) { volatile float p = a[i] * a[i]; ss = ss + p; }
    volatile float l2 = sqrtf(ss);
    volatile float sigma = l2 / nrm;                       /* (self.dim as f32).sqrt() */
    volatile float sc = sigma / 32.0f;
    float scale = sc > 1e-8f ? sc : 1e-8f;                 /* f32::max */
    int32_t sq = 0;
    for (uint32_t i = 0; i < dim; i++) {
        volatile float q = a[i] / scale;
        float r = roundf(q);
        int8_t c = 0;
        if (r == r) { if (r > 127.0f) r = 127.0f; if (r < -127.0f) r = -127.0f; c = (int8_t)r; }
        out[i] = c; sq += (int32_t)c * (int32_t)c;
    }
    *scale_out = scale;
    { volatile float x = (float)sq * scale; volatile float y = x * scale; *norm_out = y; }
    free(a);
}
/* -dot_i8_turboquant (Dot / Cosine: vector_similarity.rs:161-176, 220-235 — the reference negates it) / -euclidean_i8_turboquant (:300-320, 2058-2069) */
float orc_score_i8_turbo(const int8_t* q, float q_scale, float q_norm, const int8_t* e, float e_scale, float e_norm, uint32_t dim, uint32_t similarity) {
    int32_t dot = orc_dot_i8(q, e, dim);
    volatile float a = (float)dot * q_scale; volatile float d = a * e_scale;
    if (similarity != ORC_SIM_EUCLIDEAN) return -d;
    volatile float s = q_norm + e_norm; volatile float t = 2.0f * d; volatile float r = s - t;
    return -(r > 0.0f ? r : 0.0f);
}
int orc_search_vector_i8_turbo(const int8_t* rows, const float* row_scale, const float* row_norm, const uint32_t* doc_ids, uint64_t n_rows, uint32_t dim,
                               uint32_t row_pitch, const int8_t* query, float q_scale, float q_norm, uint32_t similarity, uint32_t k,
                               orc_hit* hits, uint32_t* n_hits) {
    topk_t tk = { hits, 0, k };
    for (uint64_t r = 0; r < n_rows; r++) {
        float s = orc_score_i8_turbo(query, q_scale, q_norm, rows + r * row_pitch, row_scale[r], row_norm[r], dim, similarity);
        topk_push(&tk, doc_ids ? doc_ids[r] : r, s);
    }
    if (n_hits) *n_hits = tk.n;
    return 0;
}

int orc_search_vector_i8_scaled(const int8_t* rows, const float* row_scale, const float* row_norm, const uint32_t* doc_ids, uint64_t n_rows, uint32_t dims,
                                uint32_t row_pitch, const int8_t* query, float q_scale, float q_norm, uint32_t similarity, uint32_t k,
                                orc_hit* hits, uint32_t* n_hits) {
    topk_t tk = { hits, 0, k };
    for (uint64_t r = 0; r < n_rows; r++) {
        float s = orc_score_i8_scaled(query, q_scale, q_norm, rows + r * row_pitch, row_scale[r], row_norm ? row_norm[r] : 0.0f, dims, similarity);
        topk_push(&tk, doc_ids ? doc_ids[r] : r, s);
    }
    if (n_hits) *n_hits = tk.n;
    return 0;
}
