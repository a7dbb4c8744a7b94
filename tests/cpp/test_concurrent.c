/* Plain C: 8 threads call ssb_search_vector / ssb_search_lexical / ssb_search_hybrid concurrently on ONE handle (the reference
 * serves many tokio tasks under a read lock, search.rs:1134-1153) and every result must equal the serial run.
 * exit codes: 0 = OK, 3 = no CUDA device (expected on the CPU-only build box), 1 = failure. */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "seekstorm_b200.h"

#define N_DOCS 40000u
#define DIMS 64u
#define N_TERMS 300u
#define NQ 24u
#define K 10u
#define THREADS 8
#define ROUNDS 6

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static ssb_index* ix;
static float* queries;            /* [THREADS][NQ][DIMS] */
static uint64_t* qkeys;           /* [THREADS][NQ][2] */
static uint32_t qoffs[NQ + 1];
static ssb_hit *want_v, *want_l, *want_h;   /* [THREADS][NQ][K] */
static int failures;

static void* worker(void* arg) {
    const long t = (long)arg;
    ssb_hit* hv = malloc(sizeof(ssb_hit) * NQ * K); ssb_hit* hl = malloc(sizeof(ssb_hit) * NQ * K); ssb_hit* hh = malloc(sizeof(ssb_hit) * NQ * K);
    uint32_t nh[NQ]; uint64_t cnt[NQ];
    ssb_lex_batch b = {NQ, SSB_QUERY_UNION, qoffs, qkeys + t * NQ * 2};
    for (int r = 0; r < ROUNDS; r++) {
        int bad = 0;
        if (ssb_search_vector(ix, queries + t * NQ * DIMS, NQ, K, hv, nh) != SSB_OK) bad = 1;
        if (ssb_search_lexical(ix, &b, K, SSB_RESULT_TOPKCOUNT, hl, nh, cnt) != SSB_OK) bad = 1;
        if (ssb_search_hybrid(ix, &b, queries + t * NQ * DIMS, K, hh, nh) != SSB_OK) bad = 1;
        if (bad) { printf("thread %ld: %s\n", t, ssb_last_error()); __sync_fetch_and_add(&failures, 1); break; }
        if (memcmp(hv, want_v + t * NQ * K, sizeof(ssb_hit) * NQ * K) || memcmp(hl, want_l + t * NQ * K, sizeof(ssb_hit) * NQ * K) ||
            memcmp(hh, want_h + t * NQ * K, sizeof(ssb_hit) * NQ * K)) {
            printf("thread %ld round %d: result differs from the serial run\n", t, r); __sync_fetch_and_add(&failures, 1);
        }
    }
    free(hv); free(hl); free(hh);
    return NULL;
}

int main(void) {
    ssb_config cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0; cfg.max_batch = 64; cfg.vector_dims = DIMS; cfg.vector_similarity = SSB_SIM_COSINE;
    int32_t rc = ssb_create(&cfg, &ix);
    if (rc == SSB_E_NO_DEVICE) { printf("no CUDA device: %s\n", ssb_last_error()); return 3; }
    if (rc != SSB_OK) { printf("create: %s\n", ssb_last_error()); return 1; }
    /* lexical level: N_TERMS terms, term t holds every doc d with (d * (t + 3)) % (t + 7) == 0 — ascending ids, tf 1..3 */
    uint64_t* keys = malloc(8 * N_TERMS); uint32_t* offs = malloc(4 * (N_TERMS + 1));
    uint16_t* ids = malloc(2 * (size_t)N_TERMS * N_DOCS); uint16_t* tfs = malloc(2 * (size_t)N_TERMS * N_DOCS);
    uint8_t* len = malloc(N_DOCS); uint64_t len_sum = 0; uint32_t np = 0;
    for (uint32_t d = 0; d < N_DOCS; d++) { len[d] = (uint8_t)(8 + d % 16); len_sum += len[d]; }   /* byte4 codes < 24 are the lengths */
    for (uint32_t t = 0; t < N_TERMS; t++) {
        keys[t] = ((uint64_t)(t + 1) * 0x9E3779B97F4A7C15ull) & ~7ull; offs[t] = np;
        for (uint32_t d = 0; d < N_DOCS; d++) if (((uint64_t)d * (t + 3)) % (t + 7) == 0) { ids[np] = (uint16_t)d; tfs[np] = (uint16_t)(1 + (d + t) % 3); np++; }
    }
    offs[N_TERMS] = np;
    ssb_level_desc lv = {0, N_DOCS, N_TERMS, 0, keys, offs, ids, tfs, len};
    if (ssb_lexical_add_level(ix, &lv) != SSB_OK || ssb_lexical_commit(ix, N_DOCS, len_sum) != SSB_OK) { printf("load: %s\n", ssb_last_error()); return 1; }
    float* rows = malloc(sizeof(float) * (size_t)N_DOCS * DIMS);
    for (size_t i = 0; i < (size_t)N_DOCS * DIMS; i++) rows[i] = (float)((double)(rnd() % 20001) / 10000.0 - 1.0);
    if (ssb_vector_add_level(ix, 0, rows, DIMS, NULL, N_DOCS, DIMS) != SSB_OK) { printf("vectors: %s\n", ssb_last_error()); return 1; }
    queries = malloc(sizeof(float) * THREADS * NQ * DIMS); qkeys = malloc(8 * THREADS * NQ * 2);
    for (size_t i = 0; i < (size_t)THREADS * NQ * DIMS; i++) queries[i] = (float)((double)(rnd() % 20001) / 10000.0 - 1.0);
    for (uint32_t i = 0; i <= NQ; i++) qoffs[i] = 2 * i;
    for (size_t i = 0; i < (size_t)THREADS * NQ * 2; i++) qkeys[i] = keys[rnd() % N_TERMS];
    want_v = malloc(sizeof(ssb_hit) * THREADS * NQ * K); want_l = malloc(sizeof(ssb_hit) * THREADS * NQ * K); want_h = malloc(sizeof(ssb_hit) * THREADS * NQ * K);
    uint32_t nh[NQ]; uint64_t cnt[NQ];
    for (long t = 0; t < THREADS; t++) {            /* serial reference */
        ssb_lex_batch b = {NQ, SSB_QUERY_UNION, qoffs, qkeys + t * NQ * 2};
        if (ssb_search_vector(ix, queries + t * NQ * DIMS, NQ, K, want_v + t * NQ * K, nh) != SSB_OK ||
            ssb_search_lexical(ix, &b, K, SSB_RESULT_TOPKCOUNT, want_l + t * NQ * K, nh, cnt) != SSB_OK ||
            ssb_search_hybrid(ix, &b, queries + t * NQ * DIMS, K, want_h + t * NQ * K, nh) != SSB_OK) { printf("serial: %s\n", ssb_last_error()); return 1; }
    }
    pthread_t th[THREADS];
    for (long t = 0; t < THREADS; t++) pthread_create(&th[t], NULL, worker, (void*)t);
    for (long t = 0; t < THREADS; t++) pthread_join(th[t], NULL);
    ssb_destroy(ix);
    printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
