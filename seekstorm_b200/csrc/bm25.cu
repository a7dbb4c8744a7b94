// bm25.cu — BM25 AND/OR top-k over block-partitioned posting lists (sm_100a).
//
// Replaces, for committed data without facets/filters/phrases (all paths /root/reference/seekstorm/src/):
//   intersection_blockid / intersection_docid   intersection.rs:2023-2301 / 112-2013   (AND)
//   union_docid_2 / union_docid_3 / single_blockid  union.rs:1168-1479, single.rs:292-417 (OR + block-max)
//   add_result_multiterm_singlefield + get_bm25f_multiterm_singlefield  add_result.rs:3418-3706, 1429-1482
//   MinHeap::add_topk  min_heap.rs:1193-1259
//
// HBM layout (built once at load):
//   post[] u32 = id16 | tf8<<16 | doclen_byte<<24 — one word per posting, all levels concatenated (level-major,
//          term-major inside a level).  The doc-length byte is co-located with the posting so scoring needs no
//          random access into the 64 KB per-level length array (the reference does that gather per candidate,
//          add_result.rs:1437-1442), and one load delivers id and payload.
//   directory: sorted dict_keys -> per-term list of (level, offset, count, block-max, bitmap) entries;
//          lists with >= 256 postings additionally get an 8 KB bitmap + 2 KB rank index for O(1) probes (the
//          reference's Bitmap container starts at 4096, compress_postinglist.rs:256-332).
//
// Execution: one batch = plan kernel (per query: dictionary lookup, per-block bound = Σ idf·block-max in
// query order, blocks sorted by bound) + a persistent scoring kernel whose warps pull (query, block) items
// ordered wave-by-wave (all queries' best block first).  A per-query global threshold θ (the k-th best key so
// far) gives block-max pruning across blocks and MAXSCORE-style essential-list pruning inside a block.
// Scores are bit-exact w.r.t. the CPU oracle: every f32 op individually rounded (__fmul_rn/__fdiv_rn/__fadd_rn),
// summed in query order from 0.0 (add_result.rs:1450-1452), idf and the 256-entry cache computed on the host.
#include "bm25.h"

#include <math.h>
#include <string.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/iterator/discard_iterator.h>

namespace ssb {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t DENSE_MIN = 256;    // lists at least this long get a bitmap + rank index (O(1) probes)
constexpr uint32_t MAX_LEVELS = 4096;  // per GPU (268M docs); plan kernel smem bound

// ================================================================= build kernels
// posting i belongs to the term whose offset range contains it (binary search over posting_offsets)
__global__ void validate_level(const uint16_t* __restrict__ ids, const uint16_t* __restrict__ tfs, const uint32_t* __restrict__ offs,
                               uint32_t n_terms, uint32_t n, uint32_t n_docs, uint32_t* bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t lo = 0, hi = n_terms;                     // term t with offs[t] <= i < offs[t+1]
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (offs[m + 1] <= i) lo = m + 1; else hi = m; }
    bool ok = lo < n_terms && ids[i] < n_docs && tfs[i] >= 1;
    if (ok && i > offs[lo] && ids[i] <= ids[i - 1]) ok = false;
    if (!ok) atomicAdd(bad, 1u);
}

__global__ void build_payload(const uint16_t* __restrict__ ids, const uint16_t* __restrict__ tfs,
                              const uint8_t* __restrict__ len_bytes, uint32_t* __restrict__ post, uint32_t n,
                              uint64_t post_base, uint64_t* exc_pos, uint32_t* exc_tf, uint32_t* exc_count,
                              uint32_t exc_cap) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t tf = tfs[i];
    uint32_t len = len_bytes[ids[i]];
    if (tf >= 255) {
        uint32_t s = atomicAdd(exc_count, 1u);
        if (s < exc_cap) { exc_pos[s] = post_base + i; exc_tf[s] = tf; }
        tf = 255;
    }
    post[i] = (uint32_t)ids[i] | ((tf | (len << 8)) << 16);   // id16 | tf8<<16 | len8<<24
}

__global__ void gather_dict(const uint64_t* __restrict__ term_keys, uint32_t n_terms, uint32_t level_idx,
                            uint64_t* __restrict__ keys_out, uint64_t* __restrict__ vals_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_terms) return;
    keys_out[i] = term_keys[i];
    vals_out[i] = ((uint64_t)level_idx << 32) | i;
}

__global__ void build_entries(const uint64_t* __restrict__ vals, uint32_t n, const uint32_t* const* __restrict__ lvl_offsets,
                              const uint64_t* __restrict__ lvl_base, uint32_t* e_level, uint64_t* e_off, uint32_t* e_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t lv = (uint32_t)(vals[i] >> 32), t = (uint32_t)vals[i];
    const uint32_t* po = lvl_offsets[lv];
    uint32_t a = po[t], b = po[t + 1];
    e_level[i] = lv; e_off[i] = lvl_base[lv] + a; e_count[i] = b - a;
}

__device__ __forceinline__ uint32_t exc_lookup(const LexView& v, uint64_t pos) {
    uint32_t lo = 0, hi = v.n_exc;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (v.exc_pos[m] < pos) lo = m + 1; else hi = m; }
    return (lo < v.n_exc && v.exc_pos[lo] == pos) ? v.exc_tf[lo] : 255u;
}

// query-independent posting score component: tf*(K+1)/(tf+cache[len])   (add_result.rs:1450)
__device__ __forceinline__ float comp_of(const LexView& v, uint32_t payload, uint64_t pos) {
    uint32_t tfu = payload & 255u;
    if (tfu == 255u) tfu = exc_lookup(v, pos);
    float tf = (float)tfu;
    return __fdiv_rn(__fmul_rn(tf, v.k1p), __fadd_rn(tf, v.cache[payload >> 8]));
}

// one warp per entry: block-max basis (get_max_score, index.rs:2938-3049 — here exact over the list)
__global__ void entry_maxcomp(LexView v, uint32_t n_entries, float* __restrict__ out) {
    uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n_entries) return;
    int lane = threadIdx.x & 31;
    uint64_t off = v.e_off[e]; uint32_t cnt = v.e_count[e];
    float m = 0.f;
    for (uint32_t i = lane; i < cnt; i += 32) m = fmaxf(m, comp_of(v, v.post[off + i] >> 16, off + i));
    for (int s = 16; s; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
    if (lane == 0) out[e] = m;
}

__global__ void mark_dense(const uint32_t* __restrict__ e_count, uint32_t n, uint32_t* flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = e_count[i] >= DENSE_MIN ? 1u : 0u;
}
__global__ void assign_bitmap(const uint32_t* __restrict__ e_count, const uint32_t* __restrict__ scan, uint32_t n, uint32_t* e_bitmap) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) e_bitmap[i] = e_count[i] >= DENSE_MIN ? scan[i] : NONE;
}
// one CTA (256 threads) per dense entry
__global__ void __launch_bounds__(256) build_bitmaps(const uint32_t* __restrict__ e_bitmap, const uint64_t* __restrict__ e_off,
                                                     const uint32_t* __restrict__ e_count, uint32_t n_entries,
                                                     const uint32_t* __restrict__ post, uint64_t* bm_words, uint16_t* bm_rank,
                                                     const uint32_t* __restrict__ dense_list) {
    __shared__ unsigned long long w[1024];
    __shared__ uint32_t pc[1024];
    uint32_t e = dense_list[blockIdx.x];
    uint32_t b = e_bitmap[e];
    for (int i = threadIdx.x; i < 1024; i += 256) w[i] = 0ull;
    __syncthreads();
    uint64_t off = e_off[e]; uint32_t cnt = e_count[e];
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
        uint32_t d = post[off + i] & 0xFFFFu;
        atomicOr(&w[d >> 6], 1ull << (d & 63));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256) pc[i] = __popcll(w[i]);
    __syncthreads();
    // exclusive prefix over 1024 counts: thread t handles words 4t..4t+3 after a block scan of 4-sums
    __shared__ uint32_t part[256];
    uint32_t s4 = pc[4 * threadIdx.x] + pc[4 * threadIdx.x + 1] + pc[4 * threadIdx.x + 2] + pc[4 * threadIdx.x + 3];
    part[threadIdx.x] = s4;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - s4;
    for (int j = 0; j < 4; j++) {
        int wi = 4 * threadIdx.x + j;
        bm_words[(size_t)b * 1024 + wi] = w[wi];
        bm_rank[(size_t)b * 1024 + wi] = (uint16_t)run;
        run += pc[wi];
    }
}
__global__ void compact_dense(const uint32_t* __restrict__ e_bitmap, uint32_t n, uint32_t* dense_list) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && e_bitmap[i] != NONE) dense_list[e_bitmap[i]] = i;
}

// ================================================================= plan kernel
// One CTA per query.  Dictionary lookup (replaces decode_posting_list_object / segment.get, search.rs:2292-2423,
// 3194-3217), live-term list in query order, per-level bound and presence count, blocks sorted by bound desc
// (intersection.rs:2224-2225, single.rs:372).  For the first FAST_T live terms the entry index of every level is
// recorded with the item so the scoring kernel needs no directory search.
#ifndef SSB_LEX_PRESENCE
#define SSB_LEX_PRESENCE 1   // OR fast path: per-doc bitmap membership filter before the exact re-score
#endif
#ifndef SSB_LEX_AND_SMEM
#define SSB_LEX_AND_SMEM 0   // AND fast path with per-term state in shared memory (unmeasured experiment, see process_item_fast)
#endif
// (measured earlier in the round: fetching 2-4 posting chunks per loop iteration was SLOWER at every occupancy — the loop body
// then still contained the whole exact re-score; worth re-measuring now that it does not)
constexpr uint32_t FAST_T = 4;          // queries with <= 4 live terms take the register-resident fast path
constexpr uint32_t ENT_NONE = 0xFFFFu;

__global__ void __launch_bounds__(128) lex_plan(LexView v, const uint32_t* __restrict__ q_off, const uint64_t* __restrict__ q_keys,
                                                uint32_t query_type, QueryPlan* plans, uint64_t* items, uint2* item_ent, uint32_t* ctr,
                                                uint64_t* theta, int* lock, uint64_t* count, uint64_t* glist, uint32_t n_pow2) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    float* bound = (float*)sm_raw;                                     // [n_levels]
    uint32_t* cnt = (uint32_t*)(bound + v.n_levels);                   // [n_levels]
    uint16_t* ent = (uint16_t*)(cnt + v.n_levels);                     // [FAST_T][n_levels] entry index relative to term.first
    uint64_t* skey = (uint64_t*)(((uintptr_t)(ent + FAST_T * v.n_levels) + 7) & ~(uintptr_t)7);  // [n_pow2]
    __shared__ QTerm st[SSB_MAX_QUERY_TERMS];
    __shared__ QueryPlan pl;
    __shared__ uint32_t n_valid;

    const uint32_t q = blockIdx.x;
    const uint32_t t0 = q_off[q], nt_raw = q_off[q + 1] - t0;
    const uint32_t nt = nt_raw > SSB_MAX_QUERY_TERMS ? SSB_MAX_QUERY_TERMS : nt_raw;
    if (threadIdx.x < 32) glist[(size_t)q * LIST + threadIdx.x] = 0;
    if (threadIdx.x == 0) { theta[q] = 0; lock[q] = 0; count[q] = 0; n_valid = 0; }
    if (threadIdx.x < nt) {
        uint64_t key = q_keys[t0 + threadIdx.x];
        uint32_t lo = 0, hi = v.n_terms;
        while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (v.dict_keys[m] < key) lo = m + 1; else hi = m; }
        QTerm t; t.first = 0; t.n = 0; t.idf = 0.f; t.df = 0;
        if (lo < v.n_terms && v.dict_keys[lo] == key && v.term_df[lo] > 0) {
            t.first = v.term_first[lo]; t.n = v.term_first[lo + 1] - t.first; t.idf = v.term_idf[lo]; t.df = v.term_df[lo];
        }
        st[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t nl = 0; bool missing = false;
        for (uint32_t t = 0; t < nt; t++) {
            if (!st[t].n) { missing = true; continue; }
            bool dup = false;                       // the reference scores unique_terms (search.rs:3023-3039): drop repeated keys
            for (uint32_t u = 0; u < nl; u++) dup = dup || pl.t[u].first == st[t].first;
            if (!dup) pl.t[nl++] = st[t];
        }
        // search.rs:3290-3296: AND with an unknown term -> empty result; OR drops the term
        if (query_type == SSB_QUERY_INTERSECTION && missing) nl = 0;
        pl.n_live = nl; pl.n_items = 0; pl.flags = 0; pl.pad = 0;
    }
    for (uint32_t b = threadIdx.x; b < v.n_levels; b += blockDim.x) {
        bound[b] = 0.f; cnt[b] = 0;
        for (uint32_t t = 0; t < FAST_T; t++) ent[t * v.n_levels + b] = (uint16_t)ENT_NONE;
    }
    __syncthreads();
    const uint32_t nl = pl.n_live;
    for (uint32_t t = 0; t < nl; t++) {   // QUERY ORDER: the bound is summed exactly like a score would be
        const QTerm qt = pl.t[t];
        for (uint32_t e = threadIdx.x; e < qt.n; e += blockDim.x) {
            uint32_t lv = v.e_level[qt.first + e];
            bound[lv] = __fadd_rn(bound[lv], __fmul_rn(qt.idf, v.e_maxcomp[qt.first + e]));
            cnt[lv] += 1;
            if (t < FAST_T) ent[t * v.n_levels + lv] = (uint16_t)e;
        }
        __syncthreads();
    }
    for (uint32_t b = threadIdx.x; b < n_pow2; b += blockDim.x) {
        uint64_t key = 0;
        if (b < v.n_levels) {
            bool ok = query_type == SSB_QUERY_INTERSECTION ? (nl > 0 && cnt[b] == nl) : (cnt[b] > 0);
            if (ok) { key = ((uint64_t)ord_f32(bound[b]) << 32) | (uint64_t)(0xFFFFFFFFu - b); atomicAdd(&n_valid, 1u); }
        }
        skey[b] = key;
    }
    __syncthreads();
    // bitonic sort, descending
    for (uint32_t size = 2; size <= n_pow2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < n_pow2 / 2; i += blockDim.x) {
                uint32_t lo = 2 * i - (i & (stride - 1));   // index with bit `stride` cleared
                uint32_t hi = lo + stride;
                bool desc = (lo & size) == 0;
                uint64_t a = skey[lo], b = skey[hi];
                if (desc ? (a < b) : (a > b)) { skey[lo] = b; skey[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t j = threadIdx.x; j < v.n_levels; j += blockDim.x) {
        const uint64_t key = skey[j];
        items[(size_t)q * v.n_levels + j] = key;
        if (key) {
            const uint32_t lv = 0xFFFFFFFFu - (uint32_t)key;
            uint2 e;
            e.x = (uint32_t)ent[lv] | ((uint32_t)ent[v.n_levels + lv] << 16);
            e.y = (uint32_t)ent[2 * v.n_levels + lv] | ((uint32_t)ent[3 * v.n_levels + lv] << 16);
            item_ent[(size_t)q * v.n_levels + j] = e;
        }
    }
    if (threadIdx.x == 0) {
        pl.n_items = n_valid;
        plans[q] = pl;
        atomicMax(&ctr[1], n_valid);
    }
}

// ================================================================= scoring kernel
struct TermRegs {   // generic path: lane t holds query term t of the current item
    uint32_t cnt; uint64_t off; uint32_t bmi; float idf; float ub;
};

// membership + rank probe of doc d in the list described by (cnt, off, bmi); posting word returned in `pw`
__device__ __forceinline__ bool probe(const LexView& v, uint32_t cnt, uint64_t off, uint32_t bmi, uint32_t d, uint32_t& rank) {
    if (bmi != NONE) {
        // both loads depend only on d: issue them together (one memory latency instead of two)
        const uint64_t w = __ldg(&v.bm_words[(size_t)bmi * 1024 + (d >> 6)]);
        const uint32_t r0 = (uint32_t)__ldg(&v.bm_rank[(size_t)bmi * 1024 + (d >> 6)]);
        rank = r0 + (uint32_t)__popcll(w & ((1ull << (d & 63)) - 1ull));
        return ((w >> (d & 63)) & 1ull) != 0;
    }
    uint32_t lo = 0, hi = cnt;
    const uint32_t* a = v.post + off;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if ((__ldg(&a[m]) & 0xFFFFu) < d) lo = m + 1; else hi = m; }
    rank = lo;
    return lo < cnt && (__ldg(&a[lo]) & 0xFFFFu) == d;
}

__device__ __forceinline__ float term_score(const LexView& v, float idf, uint64_t pos) {
    // idf * ((tf*(K+1)/(tf+comp)) + SIGMA), SIGMA = 0 (x + 0.0 == x)
    return __fmul_rn(idf, comp_of(v, __ldg(&v.post[pos]) >> 16, pos));
}

__device__ __forceinline__ void insert_candidates(uint64_t& L, uint32_t& thr, bool cand, float score, uint32_t doc,
                                                  uint32_t k, int lane, bool& dirty, uint64_t ceil) {
    // `ceil`: exclusive upper bound on keys (paging beyond 32 results: everything >= ceil was returned by an earlier page)
    const uint64_t key = pack_key(score, doc);
    unsigned m = __ballot_sync(FULL, cand && key < ceil);
    if (!m) return;
    while (m) {
        int src = __ffs(m) - 1; m &= m - 1;
        wl_insert(L, shfl64(key, src), lane);
    }
    dirty = true;
    uint32_t kth = (uint32_t)(shfl64(L, (int)k - 1) >> 32);
    if (kth > thr) thr = kth;
}

struct ItemCtx {
    uint32_t q, lv, n, k, docbase, bound_ord;
    uint64_t ceil;
    bool scoring, need_count, is_and;
};

struct FTerm { uint64_t off; uint32_t cnt, bmi; float idf, ub; uint32_t pos, cpos; };   // 32 B, one per (warp, query term)

// ---- fast path: n <= FAST_T live terms, per-term state in (warp-uniform) registers ----
__device__ __forceinline__ void process_item_fast(const LexView& v, const QueryPlan* pl, const ItemCtx& c, uint2 ient, int lane,
                                                  uint64_t& L, uint32_t& thr, bool& dirty, uint64_t& matches,
                                                  uint64_t& st_visited, uint64_t& st_probes) {
    uint32_t cnt[FAST_T], bmi[FAST_T]; uint64_t off[FAST_T]; float idf[FAST_T], ub[FAST_T];
    const uint32_t n = c.n;
#pragma unroll
    for (uint32_t t = 0; t < FAST_T; t++) {
        cnt[t] = 0; bmi[t] = NONE; off[t] = 0; idf[t] = 0.f; ub[t] = 0.f;
        const uint32_t er = (t < 2 ? (ient.x >> (16 * t)) : (ient.y >> (16 * (t - 2)))) & 0xFFFFu;
        if (t < n) {
            const QTerm qt = pl->t[t];
            idf[t] = qt.idf;
            if (er != ENT_NONE) {
                const uint32_t e = qt.first + er;
                cnt[t] = __ldg(&v.e_count[e]); off[t] = __ldg(&v.e_off[e]); bmi[t] = __ldg(&v.e_bitmap[e]);
                ub[t] = __fmul_rn(qt.idf, __ldg(&v.e_maxcomp[e]));
            }
        }
    }
#if SSB_LEX_AND_SMEM
    if (c.is_and) {
        // ---------------- AND, per-term state in shared memory (EXPERIMENT, default off: written after the round's GPU budget
        // was spent, never run; same transformation that took the OR path from 4.7 to 3.5 ms) ----------------
        __shared__ FTerm afts[8][FAST_T];
        FTerm* at = afts[(threadIdx.x >> 5) & 7];
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (uint32_t t = 0; t < FAST_T; t++) {
                FTerm f; f.off = off[t]; f.cnt = cnt[t]; f.bmi = bmi[t]; f.idf = idf[t]; f.ub = ub[t]; f.pos = 0; f.cpos = 0;
                at[t] = f;
            }
        }
        __syncwarp();
        uint32_t drv = 0, best = at[0].cnt;
        for (uint32_t t = 1; t < n; t++) { const uint32_t ct = at[t].cnt; if (ct < best) { best = ct; drv = t; } }
        const uint32_t dcnt = best; const uint64_t doff = at[drv].off; const float didf = at[drv].idf;
        st_visited += dcnt;
        for (uint32_t base = 0; base < dcnt; base += 32) {
            const uint32_t p = base + lane;
            const bool active = p < dcnt;
            const uint32_t pd = active ? __ldg(&v.post[doff + p]) : 0u;
            const uint32_t d = pd & 0xFFFFu;
            bool ok = active; float score = 0.f;
            for (uint32_t t = 0; t < n; t++) {               // query order
                if (t == drv) { if (ok && c.scoring) score = __fadd_rn(score, __fmul_rn(didf, comp_of(v, pd >> 16, doff + p))); continue; }
                if (!ok) continue;
                uint32_t rank; st_probes++;
                const uint64_t toff = at[t].off;
                if (!probe(v, at[t].cnt, toff, at[t].bmi, d, rank)) { ok = false; continue; }
                if (c.scoring) score = __fadd_rn(score, term_score(v, at[t].idf, toff + rank));
            }
            matches += __popc(__ballot_sync(FULL, ok));
            if (c.scoring) insert_candidates(L, thr, ok && ord_f32(score) >= thr, score, c.docbase | d, c.k, lane, dirty, c.ceil);
        }
        __syncwarp();
        return;
    }
#else
    if (c.is_and) {
        // ---------------- AND: drive with the shortest list (intersection.rs:258-273) ----------------
        uint32_t drv = 0, best = cnt[0];
#pragma unroll
        for (uint32_t t = 1; t < FAST_T; t++) if (t < n && cnt[t] < best) { best = cnt[t]; drv = t; }
        uint32_t dcnt = 0; uint64_t doff = 0;
#pragma unroll
        for (uint32_t t = 0; t < FAST_T; t++) if (t == drv) { dcnt = cnt[t]; doff = off[t]; }
        st_visited += dcnt;
        for (uint32_t base = 0; base < dcnt; base += 32) {
            const uint32_t p = base + lane;
            const bool active = p < dcnt;
            const uint32_t pd = active ? __ldg(&v.post[doff + p]) : 0u;
            const uint32_t d = pd & 0xFFFFu;
            bool ok = active; float score = 0.f;
#pragma unroll
            for (uint32_t t = 0; t < FAST_T; t++) {          // query order
                if (t >= n) continue;
                if (t == drv) { if (ok && c.scoring) score = __fadd_rn(score, __fmul_rn(idf[t], comp_of(v, pd >> 16, doff + p))); continue; }
                if (!ok) continue;
                uint32_t rank; st_probes++;
                if (!probe(v, cnt[t], off[t], bmi[t], d, rank)) { ok = false; continue; }
                if (c.scoring) score = __fadd_rn(score, term_score(v, idf[t], off[t] + rank));
            }
            matches += __popc(__ballot_sync(FULL, ok));
            if (c.scoring) insert_candidates(L, thr, ok && ord_f32(score) >= thr, score, c.docbase | d, c.k, lane, dirty, c.ceil);
        }
        return;
    }
#endif
    // ---------------- OR ----------------
    // Per-term state moves from registers to a per-warp shared-memory record here.  With 6 four-entry register arrays live
    // across the posting loop and a 48-register budget (5 CTAs/SM), the compiler re-evaluated the "t == drv" select chains
    // in every chunk iteration: ncu attributed 33 % of the kernel's instructions to the three per-driver setup lines.  Values
    // are warp-uniform, so every access below is one broadcast LDS with a dynamic index.
    __shared__ uint2 cbuf[8][64];                             // per-warp survivor queue (posting index, posting word)
    __shared__ FTerm fts[8][FAST_T];
    uint2* mybuf = cbuf[(threadIdx.x >> 5) & 7];
    FTerm* ft = fts[(threadIdx.x >> 5) & 7];
    {
        // MAXSCORE order: terms by block bound desc (present first); pos = rank of the term
        uint32_t pos[FAST_T];
#pragma unroll
        for (uint32_t t = 0; t < FAST_T; t++) {
            uint32_t r = 0;
#pragma unroll
            for (uint32_t u = 0; u < FAST_T; u++) {
                if (u == t || u >= n) continue;
                const bool before = (cnt[u] > 0 && cnt[t] == 0) || ((cnt[u] > 0) == (cnt[t] > 0) && (ub[u] > ub[t] || (ub[u] == ub[t] && u < t)));
                if (before) r++;
            }
            pos[t] = t < n ? r : 0xFFFFu;
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (uint32_t t = 0; t < FAST_T; t++) {
                FTerm f; f.off = off[t]; f.cnt = cnt[t]; f.bmi = bmi[t]; f.idf = idf[t]; f.ub = ub[t]; f.pos = pos[t]; f.cpos = 0xFFFFu;
                ft[t] = f;
            }
        }
        __syncwarp();
    }
    if (c.scoring) {
        for (uint32_t p = 0; p < n; p++) {
            uint32_t drv = 0;
            for (uint32_t t = 0; t < n; t++) if (ft[t].pos == p) drv = t;
            const uint32_t dcnt = ft[drv].cnt; const uint64_t doff = ft[drv].off; const float didf = ft[drv].idf;
            if (dcnt == 0) break;                         // absent terms sort last
            // a driver is essential while the in-query-order sum of the not-yet-driven bounds can reach theta
            float S = 0.f;
            for (uint32_t t = 0; t < n; t++) if (ft[t].pos >= p) S = __fadd_rn(S, ft[t].ub);
            if (ord_f32(S) < thr) break;
            st_visited += dcnt;
            // R = in-query-order sum of the bounds of the later-ranked (not yet driven) terms.  The per-posting filter is
            // (approx(cd) + R) * (1 + 2e-6) >= theta: cheap (no IEEE divide, no per-term loop) and still a strict upper bound
            // of the exact in-order score — the inflation covers the approximate reciprocal (<= 2 ulp) and the different
            // association of <= 4 additions (<= 3 ulp).  Survivors are re-scored exactly below.
            float R = 0.f;
            for (uint32_t t = 0; t < n; t++) if (t != drv && ft[t].pos > p) R = __fadd_rn(R, ft[t].ub);
            const float didf_k = didf * v.k1p;
            // exact contribution, probes, exact in-order score of one queued survivor per lane
            auto rescore = [&](uint32_t pp, uint32_t pd, bool alive) {
                const uint32_t d = pd & 0xFFFFu;
                const float cd = alive ? __fmul_rn(didf, comp_of(v, pd >> 16, doff + pp)) : 0.f;
                float score = 0.f;
                for (uint32_t t = 0; t < n; t++) {           // query order
                    if (t == drv) { score = __fadd_rn(score, cd); continue; }
                    const uint32_t tc = ft[t].cnt;
                    if (!alive || tc == 0) continue;
                    uint32_t rank; st_probes++;
                    const uint64_t toff = ft[t].off;
                    if (probe(v, tc, toff, ft[t].bmi, d, rank)) {
                        if (ft[t].pos < p) alive = false;     // already emitted when that term was the driver
                        else score = __fadd_rn(score, term_score(v, ft[t].idf, toff + rank));
                    }
                }
                insert_candidates(L, thr, alive && ord_f32(score) >= thr, score, c.docbase | d, c.k, lane, dirty, c.ceil);
            };
            uint32_t nbuf = 0;                            // queued survivors of this warp (warp-uniform, < 32 between chunks)
            uint32_t pd_next = (uint32_t)lane < dcnt ? __ldg(&v.post[doff + lane]) : 0u;
            for (uint32_t base = 0; base < dcnt; base += 32u) {
                // software pipelining: the next chunk's postings are requested before this chunk is processed
                const uint32_t pd = pd_next;
                { const uint32_t pn = base + 32u + lane; pd_next = pn < dcnt ? __ldg(&v.post[doff + pn]) : 0u; }
                const uint32_t pp = base + lane;
                const bool active = pp < dcnt;
                const uint32_t d = pd & 0xFFFFu;
                const uint32_t tf8 = (pd >> 16) & 255u;
                const float tfb = tf8 == 255u ? 65535.f : (float)tf8;                 // overflowed tf: bound with the u16 maximum
                const float cdb = didf_k * __fdividef(tfb, tfb + __ldg(&v.cache[pd >> 24]));
                bool alive = active && ord_f32((cdb + R) * 1.000002f) >= thr;
                if (!__any_sync(FULL, alive)) continue;
#if SSB_LEX_PRESENCE
                // ---- second filter: replace the level-wide bound R by what the bitmaps say about THIS doc.  One 8-byte load per
                // bitmap-backed term gives membership; a doc that is in an earlier-ranked list was already emitted there, and a term
                // the doc is not in contributes nothing.  No exact divide, no rank / payload load for postings that die here
                // (measured before this filter: 0.75 probes per enumerated posting, 0.10 after).  Lists without a bitmap count as
                // "maybe".
                {
                    float B = cdb; bool dead = false;
                    for (uint32_t t = 0; t < n; t++) {
                        if (t == drv || ft[t].cnt == 0) continue;
                        const uint32_t tb = ft[t].bmi, tp = ft[t].pos;
                        if (tb != NONE) {
                            const uint64_t w = alive ? __ldg(&v.bm_words[(size_t)tb * 1024 + (d >> 6)]) : 0ull;
                            if ((w >> (d & 63)) & 1ull) { if (tp < p) dead = true; else B += ft[t].ub; }
                        } else if (tp > p) B += ft[t].ub;
                    }
                    alive = alive && !dead && ord_f32(B * 1.000002f) >= thr;
                }
#endif
                // ---- compaction: the exact re-score is ~10x the cost of the filters and only a few lanes of a chunk survive them
                // (but most chunks have at least one survivor).  Survivors are queued per warp and re-scored 32 at a time, so the
                // expensive path runs with full lanes.  Order does not matter: keys are a total order, the top-k is a set.
                const unsigned am = __ballot_sync(FULL, alive);
                if (!am) continue;
                if (alive) mybuf[nbuf + __popc(am & ((1u << lane) - 1u))] = make_uint2(pp, pd);
                nbuf += __popc(am);
                __syncwarp();
                if (nbuf >= 32u) {
                    const uint2 e = mybuf[lane];
                    rescore(e.x, e.y, true);
                    const uint32_t rem = nbuf - 32u;
                    uint2 tmp = make_uint2(0u, 0u);
                    if ((uint32_t)lane < rem) tmp = mybuf[32 + lane];
                    __syncwarp();
                    if ((uint32_t)lane < rem) mybuf[lane] = tmp;
                    __syncwarp();
                    nbuf = rem;
                }
            }
            if (nbuf) {
                const bool act = (uint32_t)lane < nbuf;
                const uint2 e = act ? mybuf[lane] : make_uint2(0u, 0u);
                rescore(e.x, e.y, act);
                __syncwarp();
            }
        }
    }
    if (c.need_count) {
        // exact |union| of this block: sum of counts - duplicates; enumerate all but the longest list, probe longer ones
        if (lane == 0) {
            for (uint32_t t = 0; t < n; t++) {
                uint32_t r = 0;
                for (uint32_t u = 0; u < n; u++) if (u != t && (ft[u].cnt > ft[t].cnt || (ft[u].cnt == ft[t].cnt && u < t))) r++;
                ft[t].cpos = r;
            }
        }
        __syncwarp();
        for (uint32_t p = 0; p < n; p++) {
            uint32_t dcnt = 0; uint64_t doff = 0;
            for (uint32_t t = 0; t < n; t++) if (ft[t].cpos == p) { dcnt = ft[t].cnt; doff = ft[t].off; }
            if (dcnt == 0) break;
            if (p == 0) { matches += dcnt; continue; }
            st_visited += dcnt;
            for (uint32_t base = 0; base < dcnt; base += 32) {
                const uint32_t pp = base + lane;
                const bool active = pp < dcnt;
                const uint32_t d = active ? (__ldg(&v.post[doff + pp]) & 0xFFFFu) : 0u;
                bool dup = false;
                for (uint32_t t = 0; t < n; t++) {
                    const uint32_t tc = ft[t].cnt;
                    if (ft[t].cpos >= p || tc == 0 || !active || dup) continue;
                    uint32_t rank; st_probes++;
                    if (probe(v, tc, ft[t].off, ft[t].bmi, d, rank)) dup = true;
                }
                matches += __popc(__ballot_sync(FULL, active && !dup));
            }
        }
        __syncwarp();
    }
}

// ---- generic path: up to SSB_MAX_QUERY_TERMS live terms, lane t holds term t, values broadcast by shuffles ----
__device__ __noinline__ void process_item_generic(const LexView& v, const QueryPlan* pl, const ItemCtx& c, int lane,
                                                  uint64_t& L, uint32_t& thr, bool& dirty, uint64_t& matches,
                                                  uint64_t& st_visited, uint64_t& st_probes) {
    const uint32_t n = c.n, lv = c.lv;
    TermRegs tr; tr.cnt = 0; tr.off = 0; tr.bmi = NONE; tr.idf = 0.f; tr.ub = 0.f;
    if ((uint32_t)lane < n) {
        QTerm qt = pl->t[lane];
        uint32_t lo = 0, hi = qt.n;
        while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (__ldg(&v.e_level[qt.first + m]) < lv) lo = m + 1; else hi = m; }
        if (lo < qt.n && __ldg(&v.e_level[qt.first + lo]) == lv) {
            uint32_t e = qt.first + lo;
            tr.cnt = __ldg(&v.e_count[e]); tr.off = __ldg(&v.e_off[e]); tr.bmi = __ldg(&v.e_bitmap[e]);
            tr.ub = __fmul_rn(qt.idf, __ldg(&v.e_maxcomp[e]));
        }
        tr.idf = qt.idf;
    }
    if (c.is_and) {
        uint32_t cc = (uint32_t)lane < n ? tr.cnt : 0xFFFFFFFFu;
        uint32_t key = cc; int drv = lane;
        for (int s = 16; s; s >>= 1) {
            uint32_t ok = __shfl_xor_sync(FULL, key, s); int od = __shfl_xor_sync(FULL, drv, s);
            if (ok < key || (ok == key && od < drv)) { key = ok; drv = od; }
        }
        const uint32_t dcnt = __shfl_sync(FULL, tr.cnt, drv);
        const uint64_t doff = shfl64(tr.off, drv);
        st_visited += dcnt;
        for (uint32_t base = 0; base < dcnt; base += 32) {
            const uint32_t p = base + lane;
            const bool active = p < dcnt;
            const uint32_t d = active ? (__ldg(&v.post[doff + p]) & 0xFFFFu) : 0u;
            bool ok = active; float score = 0.f;
            for (uint32_t t = 0; t < n; t++) {          // query order
                const uint32_t tc = __shfl_sync(FULL, tr.cnt, t); const uint64_t to = shfl64(tr.off, t);
                const uint32_t tb = __shfl_sync(FULL, tr.bmi, t); const float ti = __shfl_sync(FULL, tr.idf, t);
                uint32_t rank = p; bool found = true;
                if ((int)t != drv) { found = ok && probe(v, tc, to, tb, d, rank); st_probes += ok ? 1 : 0; }
                ok = ok && found;
                if (ok && c.scoring) score = __fadd_rn(score, term_score(v, ti, to + rank));
            }
            matches += __popc(__ballot_sync(FULL, ok));
            if (c.scoring) insert_candidates(L, thr, ok && ord_f32(score) >= thr, score, c.docbase | d, c.k, lane, dirty, c.ceil);
        }
        return;
    }
    if (c.scoring) {
        uint32_t rk = 0;
        for (uint32_t t = 0; t < n; t++) {
            float ou = __shfl_sync(FULL, tr.ub, t); uint32_t oc = __shfl_sync(FULL, tr.cnt, t);
            bool mine = (uint32_t)lane < n;
            bool before = (oc > 0 && tr.cnt == 0) || ((oc > 0) == (tr.cnt > 0) && (ou > tr.ub || (ou == tr.ub && t < (uint32_t)lane)));
            if (mine && before && t != (uint32_t)lane) rk++;
        }
        if ((uint32_t)lane >= n) rk = 0xFFFFu;
        for (uint32_t p = 0; p < n; p++) {
            const int drv = __ffs(__ballot_sync(FULL, rk == p)) - 1;
            const uint32_t dcnt = __shfl_sync(FULL, tr.cnt, drv);
            if (dcnt == 0) break;
            float S = 0.f;
            for (uint32_t t = 0; t < n; t++) {
                float ou = __shfl_sync(FULL, tr.ub, t); uint32_t orr = __shfl_sync(FULL, rk, t);
                if (orr >= p) S = __fadd_rn(S, ou);
            }
            if (ord_f32(S) < thr) break;
            const uint64_t doff = shfl64(tr.off, drv);
            const float didf = __shfl_sync(FULL, tr.idf, drv);
            st_visited += dcnt;
            for (uint32_t base = 0; base < dcnt; base += 32) {
                const uint32_t pp = base + lane;
                const bool active = pp < dcnt;
                const uint32_t d = active ? (__ldg(&v.post[doff + pp]) & 0xFFFFu) : 0u;
                bool dup = false; float score = 0.f;
                for (uint32_t t = 0; t < n; t++) {      // query order
                    const uint32_t tc = __shfl_sync(FULL, tr.cnt, t); const uint64_t to = shfl64(tr.off, t);
                    const uint32_t tb = __shfl_sync(FULL, tr.bmi, t); const float ti = __shfl_sync(FULL, tr.idf, t);
                    const uint32_t trk = __shfl_sync(FULL, rk, t);
                    if ((int)t == drv) { if (active) score = __fadd_rn(score, term_score(v, didf, doff + pp)); continue; }
                    if (tc == 0 || !active || dup) continue;
                    uint32_t rank; st_probes++;
                    if (probe(v, tc, to, tb, d, rank)) {
                        if (trk < p) dup = true;
                        else score = __fadd_rn(score, term_score(v, ti, to + rank));
                    }
                }
                insert_candidates(L, thr, active && !dup && ord_f32(score) >= thr, score, c.docbase | d, c.k, lane, dirty, c.ceil);
            }
        }
    }
    if (c.need_count) {
        uint32_t crk = 0;
        for (uint32_t t = 0; t < n; t++) {
            uint32_t oc = __shfl_sync(FULL, tr.cnt, t);
            if ((uint32_t)lane < n && t != (uint32_t)lane && (oc > tr.cnt || (oc == tr.cnt && t < (uint32_t)lane))) crk++;
        }
        if ((uint32_t)lane >= n) crk = 0xFFFFu;
        for (uint32_t p = 0; p < n; p++) {
            const int drv = __ffs(__ballot_sync(FULL, crk == p)) - 1;
            const uint32_t dcnt = __shfl_sync(FULL, tr.cnt, drv);
            if (dcnt == 0) break;
            if (p == 0) { matches += dcnt; continue; }
            const uint64_t doff = shfl64(tr.off, drv);
            st_visited += dcnt;
            for (uint32_t base = 0; base < dcnt; base += 32) {
                const uint32_t pp = base + lane;
                const bool active = pp < dcnt;
                const uint32_t d = active ? (__ldg(&v.post[doff + pp]) & 0xFFFFu) : 0u;
                bool dup = false;
                for (uint32_t t = 0; t < n; t++) {
                    const uint32_t tc = __shfl_sync(FULL, tr.cnt, t); const uint64_t to = shfl64(tr.off, t);
                    const uint32_t tb = __shfl_sync(FULL, tr.bmi, t); const uint32_t trk = __shfl_sync(FULL, crk, t);
                    if (trk >= p || tc == 0 || !active || dup) continue;
                    uint32_t rank; st_probes++;
                    if (probe(v, tc, to, tb, d, rank)) dup = true;
                }
                matches += __popc(__ballot_sync(FULL, active && !dup));
            }
        }
    }
}

#ifndef SSB_LEX_MINB
#define SSB_LEX_MINB 5
#endif
__global__ void __launch_bounds__(256, SSB_LEX_MINB) lex_score(LexView v, const QueryPlan* __restrict__ plans, const uint64_t* __restrict__ items,
                                                 const uint2* __restrict__ item_ent, uint32_t nq, uint32_t query_type, uint32_t result_type,
                                                 uint32_t k, uint32_t* ctr, uint64_t* theta, int* lock, uint64_t* count, uint64_t* glist,
                                                 LexStats* stats, const uint64_t* __restrict__ ceil_keys) {
    const int lane = threadIdx.x & 31;
    const uint32_t max_items = *(volatile uint32_t*)&ctr[1];
    const uint64_t total = (uint64_t)max_items * nq;
    uint64_t st_visited = 0, st_probes = 0, st_done = 0, st_skipped = 0;
    const bool want_topk = result_type != SSB_RESULT_COUNT && k > 0;
    const bool need_count = result_type != SSB_RESULT_TOPK;

    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&ctr[0], 1u);
        i = __shfl_sync(FULL, i, 0);
        if ((uint64_t)i >= total) break;
        const uint32_t j = i / nq, q = i - j * nq;
        const QueryPlan* pl = &plans[q];
        if (j >= pl->n_items) continue;
        const uint64_t item = items[(size_t)q * v.n_levels + j];
        ItemCtx c;
        c.ceil = ceil_keys ? __ldg(&ceil_keys[q]) : ~0ull;
        if (c.ceil == 0) continue;                       // this query's result list is already exhausted
        c.q = q; c.n = pl->n_live; c.k = k; c.lv = 0xFFFFFFFFu - (uint32_t)item; c.bound_ord = (uint32_t)(item >> 32);
        uint32_t thr = (uint32_t)(__ldcg(&theta[q]) >> 32);
        c.scoring = want_topk && c.bound_ord >= thr;
        c.need_count = need_count; c.is_and = query_type == SSB_QUERY_INTERSECTION;
        if (!c.scoring && !need_count) { st_skipped++; continue; }
        st_done++;
        c.docbase = __ldg(&v.level_ids[c.lv]) << 16;
        uint64_t L = 0; bool dirty = false; uint64_t matches = 0;
        if (c.n <= FAST_T) process_item_fast(v, pl, c, item_ent[(size_t)q * v.n_levels + j], lane, L, thr, dirty, matches, st_visited, st_probes);
        else process_item_generic(v, pl, c, lane, L, thr, dirty, matches, st_visited, st_probes);

        // ---- publish: merge the warp list into the query's global list, raise theta ----
        if (dirty) {
            if (lane == 0) { while (atomicCAS(&lock[q], 0, 1) != 0) __nanosleep(40); }
            __syncwarp();
            __threadfence();
            uint64_t G = __ldcg(&glist[(size_t)q * LIST + lane]);
            uint64_t M = wl_merge(L, G, lane);
            __stcg(&glist[(size_t)q * LIST + lane], M);
            uint64_t nth = shfl64(M, (int)k - 1);
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                if (nth > __ldcg(&theta[q])) __stcg(&theta[q], nth);
                __threadfence();
                atomicExch(&lock[q], 0);
            }
        }
        if (need_count && lane == 0 && matches) atomicAdd((unsigned long long*)&count[q], (unsigned long long)matches);
    }
    if (lane == 0) {
        atomicAdd((unsigned long long*)&stats->postings_visited, (unsigned long long)st_visited);
        atomicAdd((unsigned long long*)&stats->items_processed, (unsigned long long)st_done);
        atomicAdd((unsigned long long*)&stats->items_skipped, (unsigned long long)st_skipped);
    }
    unsigned long long pr = st_probes;
    for (int s = 16; s; s >>= 1) pr += __shfl_xor_sync(FULL, pr, s);
    if (lane == 0) atomicAdd((unsigned long long*)&stats->probes, pr);
}

__global__ void copy_out(const uint64_t* __restrict__ glist, const uint64_t* __restrict__ count, uint32_t nq, uint32_t k,
                         uint64_t* keys_out, uint64_t* count_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * LIST) { uint32_t j = i & 31; keys_out[i] = j < k ? glist[i] : 0; }
    if (count_out && i < nq) count_out[i] = count[i];
}

// ================================================================= host side
LexIndex::~LexIndex() {
    for (auto& l : levels_) { cudaFree(l.d_term_keys); cudaFree(l.d_posting_offsets); }
    free_committed(); free_workspace();
    if (d_exc_count_) cudaFree(d_exc_count_);
}

void LexIndex::free_committed() {
    cudaFree(d_dict_keys_); cudaFree(d_term_first_); cudaFree(d_term_idf_); cudaFree(d_term_df_);
    cudaFree(d_e_level_); cudaFree(d_e_off_); cudaFree(d_e_count_); cudaFree(d_e_maxcomp_); cudaFree(d_e_bitmap_);
    cudaFree(d_bm_words_); cudaFree(d_bm_rank_); cudaFree(d_level_ids_); cudaFree(d_cache_);
    d_dict_keys_ = nullptr; d_term_first_ = nullptr; d_term_idf_ = nullptr; d_term_df_ = nullptr;
    d_e_level_ = nullptr; d_e_off_ = nullptr; d_e_count_ = nullptr; d_e_maxcomp_ = nullptr; d_e_bitmap_ = nullptr;
    d_bm_words_ = nullptr; d_bm_rank_ = nullptr; d_level_ids_ = nullptr; d_cache_ = nullptr;
    committed_ = false;
}

void LexIndex::free_workspace() {
    cudaFree(d_plans_); cudaFree(d_items_); cudaFree(d_item_ent_); cudaFree(d_theta_); cudaFree(d_lock_); cudaFree(d_count_); cudaFree(d_ctr_);
    cudaFree(d_qoff_); cudaFree(d_qkeys_); cudaFree(d_stats_);
    d_plans_ = nullptr; d_items_ = nullptr; d_item_ent_ = nullptr; d_theta_ = nullptr; d_lock_ = nullptr; d_count_ = nullptr; d_ctr_ = nullptr;
    d_qoff_ = nullptr; d_qkeys_ = nullptr; d_stats_ = nullptr; ws_nq_ = ws_terms_ = ws_levels_ = 0;
}

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// copy n bytes from a host-or-device pointer into device memory
static cudaError_t to_device(void* dst, const void* src, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, st);
}

int32_t LexIndex::add_level(const ssb_level_desc* d) {
    if (!d || d->n_docs == 0 || d->n_docs > 65536) { set_error("add_level: n_docs must be in 1..65536"); return SSB_E_INVALID; }
    if (levels_.size() >= MAX_LEVELS) { set_error("add_level: more than %u levels per GPU unsupported", MAX_LEVELS); return SSB_E_UNSUPPORTED; }
    for (auto& l : levels_) if (l.level_id == d->level_id) { set_error("add_level: duplicate level_id %u", d->level_id); return SSB_E_INVALID; }
    if (!levels_.empty() && d->level_id < levels_.back().level_id) { set_error("add_level: levels must be added in ascending level_id order"); return SSB_E_INVALID; }
    uint32_t np = 0;
    if (d->n_terms) {
        if (is_device_ptr(d->posting_offsets)) SSB_CUDA_TRY(cudaMemcpy(&np, d->posting_offsets + d->n_terms, 4, cudaMemcpyDeviceToHost));
        else np = d->posting_offsets[d->n_terms];
    }
    LexLevel l{};
    l.level_id = d->level_id; l.n_docs = d->n_docs; l.n_terms = d->n_terms; l.post_base = n_post_; l.n_post = np;
    SSB_CUDA_TRY(cudaMalloc(&l.d_term_keys, (size_t)(d->n_terms ? d->n_terms : 1) * 8));
    SSB_CUDA_TRY(cudaMalloc(&l.d_posting_offsets, ((size_t)d->n_terms + 1) * 4));
    SSB_CUDA_TRY(to_device(l.d_term_keys, d->term_keys, (size_t)d->n_terms * 8, st_));
    if (d->n_terms) SSB_CUDA_TRY(to_device(l.d_posting_offsets, d->posting_offsets, ((size_t)d->n_terms + 1) * 4, st_));
    else SSB_CUDA_TRY(cudaMemsetAsync(l.d_posting_offsets, 0, 4, st_));
    SSB_TRY(post_.reserve(n_post_ + np + 8, n_post_, st_));
    // posting word = id16 | tf8<<16 | len8<<24 (needs ids, tfs and the level's length bytes on the device)
    uint16_t* d_ids = nullptr; uint16_t* d_tfs = nullptr; uint8_t* d_len = nullptr;
    DevTmp<uint16_t> t_ids, t_tfs; DevTmp<uint8_t> t_len; DevTmp<uint32_t> t_bad;
    if (np) {
        if (is_device_ptr(d->doc_ids)) d_ids = const_cast<uint16_t*>(d->doc_ids);
        else { SSB_CUDA_TRY(t_ids.alloc(np)); d_ids = t_ids.p; SSB_CUDA_TRY(to_device(d_ids, d->doc_ids, (size_t)np * 2, st_)); }
        if (is_device_ptr(d->tfs)) d_tfs = const_cast<uint16_t*>(d->tfs);
        else { SSB_CUDA_TRY(t_tfs.alloc(np)); d_tfs = t_tfs.p; SSB_CUDA_TRY(to_device(d_tfs, d->tfs, (size_t)np * 2, st_)); }
        if (is_device_ptr(d->doc_len_bytes)) d_len = const_cast<uint8_t*>(d->doc_len_bytes);
        else { SSB_CUDA_TRY(t_len.alloc(d->n_docs)); d_len = t_len.p; SSB_CUDA_TRY(to_device(d_len, d->doc_len_bytes, d->n_docs, st_)); }
        // input contract: ids ascending and unique inside a term, < n_docs, tf >= 1 (a violated contract would make the
        // scoring kernel read out of bounds or mis-rank silently)
        SSB_CUDA_TRY(t_bad.alloc(1)); SSB_CUDA_TRY(cudaMemsetAsync(t_bad.p, 0, 4, st_));
        validate_level<<<(np + 255) / 256, 256, 0, st_>>>(d_ids, d_tfs, l.d_posting_offsets, d->n_terms, np, d->n_docs, t_bad.p);
        SSB_CUDA_TRY(cudaGetLastError());
        const uint32_t exc_cap = 1u << 20;
        if (!d_exc_count_) {
            SSB_CUDA_TRY(cudaMalloc(&d_exc_count_, 4)); SSB_CUDA_TRY(cudaMemsetAsync(d_exc_count_, 0, 4, st_));
            SSB_TRY(exc_pos_.reserve(exc_cap, 0, st_)); SSB_TRY(exc_tf_.reserve(exc_cap, 0, st_));
        }
        build_payload<<<(np + 255) / 256, 256, 0, st_>>>(d_ids, d_tfs, d_len, post_.p + n_post_, np, n_post_,
                                                         exc_pos_.p, exc_tf_.p, d_exc_count_, exc_cap);
        SSB_CUDA_TRY(cudaGetLastError());
    }
    SSB_CUDA_TRY(cudaStreamSynchronize(st_));
    if (np) {
        uint32_t bad = 0;
        SSB_CUDA_TRY(cudaMemcpy(&bad, t_bad.p, 4, cudaMemcpyDeviceToHost));
        if (bad) {
            cudaFree(l.d_term_keys); cudaFree(l.d_posting_offsets);
            set_error("add_level %u: malformed postings (%u violations: ids must ascend strictly inside a term and be < n_docs, tf >= 1)", d->level_id, bad);
            return SSB_E_INVALID;
        }
    }
    n_post_ += np;
    levels_.push_back(l);
    committed_ = false;
    return SSB_OK;
}

static void host_bm25_cache(uint64_t n_docs, uint64_t len_sum, float* cache) {
    // commit.rs:318-325; DOCUMENT_LENGTH_COMPRESSION = byte4_to_int (index.rs:4255-4279).  volatile keeps every
    // f32 operation individually rounded regardless of host compiler contraction settings.
    volatile float avgdl = (float)len_sum / (float)n_docs;
    const float K = 1.2f, B = 0.75f;
    for (int i = 0; i < 256; i++) {
        uint32_t b = (uint32_t)i, v;
        if (b < 24) v = b;
        else { uint32_t x = b - 24, bits = x & 7, shift = x >> 3; v = shift == 0 ? 24 + bits : 24 + ((bits | 8) << (shift - 1)); }
        volatile float quot = (float)v / avgdl;
        volatile float bq = B * quot;
        volatile float omb = 1.0f - B;
        volatile float inner = omb + bq;
        cache[i] = K * inner;
    }
}

static float host_idf(uint64_t n_docs, uint32_t df) {
    // search.rs:3225-3230
    volatile float a = (float)n_docs - (float)df;
    volatile float num = a + 0.5f;
    volatile float den = (float)df + 0.5f;
    volatile float r = num / den;
    volatile float r1 = r + 1.0f;
    return logf(r1);
}

int32_t LexIndex::commit(uint64_t n_docs, uint64_t len_sum) {
    if (n_docs == 0) { set_error("commit: n_docs must be > 0"); return SSB_E_INVALID; }
    free_committed();
    n_docs_ = n_docs; len_sum_ = len_sum;
    const uint32_t nlv = (uint32_t)levels_.size();
    uint64_t total64 = 0;
    for (auto& l : levels_) total64 += l.n_terms;
    if (total64 >= 0xFFFFFFFFull) { set_error("commit: too many (term, level) entries"); return SSB_E_UNSUPPORTED; }
    const uint32_t total = (uint32_t)total64;
    n_entries_ = total;
    auto pol = thrust::cuda::par.on(st_);

    float cache[256];
    host_bm25_cache(n_docs, len_sum, cache);
    SSB_CUDA_TRY(cudaMalloc(&d_cache_, 256 * 4));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_cache_, cache, 256 * 4, cudaMemcpyHostToDevice, st_));
    std::vector<uint32_t> lids(nlv ? nlv : 1); std::vector<uint64_t> lbase(nlv ? nlv : 1); std::vector<const uint32_t*> loffs(nlv ? nlv : 1);
    for (uint32_t i = 0; i < nlv; i++) { lids[i] = levels_[i].level_id; lbase[i] = levels_[i].post_base; loffs[i] = levels_[i].d_posting_offsets; }
    SSB_CUDA_TRY(cudaMalloc(&d_level_ids_, lids.size() * 4));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_level_ids_, lids.data(), lids.size() * 4, cudaMemcpyHostToDevice, st_));

    // exceptions sorted by posting position
    if (d_exc_count_) {
        SSB_CUDA_TRY(cudaMemcpyAsync(&n_exc_, d_exc_count_, 4, cudaMemcpyDeviceToHost, st_));
        SSB_CUDA_TRY(cudaStreamSynchronize(st_));
        if (n_exc_ > (1u << 20)) { set_error("commit: more than 2^20 postings with tf >= 255"); return SSB_E_UNSUPPORTED; }
        if (n_exc_) thrust::sort_by_key(pol, thrust::device_ptr<uint64_t>(exc_pos_.p), thrust::device_ptr<uint64_t>(exc_pos_.p + n_exc_),
                                        thrust::device_ptr<uint32_t>(exc_tf_.p));
    }

    size_t alloc_n = total ? total : 1;
    DevTmp<uint64_t> t_keys, t_vals, t_lbase, t_ukeys; DevTmp<const uint32_t*> t_loffs; DevTmp<uint32_t> t_epc, t_df;
    SSB_CUDA_TRY(t_keys.alloc(alloc_n)); SSB_CUDA_TRY(t_vals.alloc(alloc_n));
    SSB_CUDA_TRY(t_lbase.alloc(lbase.size())); SSB_CUDA_TRY(t_loffs.alloc(loffs.size()));
    uint64_t *d_keys = t_keys.p, *d_vals = t_vals.p, *d_lbase = t_lbase.p; const uint32_t** d_loffs = t_loffs.p;
    SSB_CUDA_TRY(cudaMemcpyAsync(d_lbase, lbase.data(), lbase.size() * 8, cudaMemcpyHostToDevice, st_));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_loffs, loffs.data(), loffs.size() * sizeof(void*), cudaMemcpyHostToDevice, st_));
    uint32_t pos = 0;
    for (uint32_t i = 0; i < nlv; i++) {
        if (levels_[i].n_terms) gather_dict<<<(levels_[i].n_terms + 255) / 256, 256, 0, st_>>>(levels_[i].d_term_keys, levels_[i].n_terms, i, d_keys + pos, d_vals + pos);
        pos += levels_[i].n_terms;
    }
    SSB_CUDA_TRY(cudaGetLastError());
    // stable: entries of one term stay in ascending level order
    thrust::stable_sort_by_key(pol, thrust::device_ptr<uint64_t>(d_keys), thrust::device_ptr<uint64_t>(d_keys + total), thrust::device_ptr<uint64_t>(d_vals));

    SSB_CUDA_TRY(cudaMalloc(&d_e_level_, alloc_n * 4)); SSB_CUDA_TRY(cudaMalloc(&d_e_off_, alloc_n * 8));
    SSB_CUDA_TRY(cudaMalloc(&d_e_count_, alloc_n * 4)); SSB_CUDA_TRY(cudaMalloc(&d_e_maxcomp_, alloc_n * 4));
    SSB_CUDA_TRY(cudaMalloc(&d_e_bitmap_, alloc_n * 4));
    if (total) build_entries<<<(total + 255) / 256, 256, 0, st_>>>(d_vals, total, d_loffs, d_lbase, d_e_level_, d_e_off_, d_e_count_);
    SSB_CUDA_TRY(cudaGetLastError());

    // dictionary: unique keys, entries per term, df per term
    SSB_CUDA_TRY(t_ukeys.alloc(alloc_n)); SSB_CUDA_TRY(t_epc.alloc(alloc_n + 1)); SSB_CUDA_TRY(t_df.alloc(alloc_n));
    uint64_t* d_ukeys = t_ukeys.p; uint32_t *d_epc = t_epc.p, *d_df = t_df.p;
    uint32_t nt = 0;
    if (total) {
        auto e1 = thrust::reduce_by_key(pol, thrust::device_ptr<uint64_t>(d_keys), thrust::device_ptr<uint64_t>(d_keys + total),
                                        thrust::constant_iterator<uint32_t>(1), thrust::device_ptr<uint64_t>(d_ukeys), thrust::device_ptr<uint32_t>(d_epc));
        nt = (uint32_t)(e1.first - thrust::device_ptr<uint64_t>(d_ukeys));
        thrust::reduce_by_key(pol, thrust::device_ptr<uint64_t>(d_keys), thrust::device_ptr<uint64_t>(d_keys + total),
                              thrust::device_ptr<uint32_t>(d_e_count_), thrust::make_discard_iterator(), thrust::device_ptr<uint32_t>(d_df));
    }
    n_terms_ = nt;
    size_t nt_alloc = nt ? nt : 1;
    SSB_CUDA_TRY(cudaMalloc(&d_dict_keys_, nt_alloc * 8)); SSB_CUDA_TRY(cudaMalloc(&d_term_first_, (nt_alloc + 1) * 4));
    SSB_CUDA_TRY(cudaMalloc(&d_term_idf_, nt_alloc * 4)); SSB_CUDA_TRY(cudaMalloc(&d_term_df_, nt_alloc * 4));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_dict_keys_, d_ukeys, (size_t)nt * 8, cudaMemcpyDeviceToDevice, st_));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_term_df_, d_df, (size_t)nt * 4, cudaMemcpyDeviceToDevice, st_));
    SSB_CUDA_TRY(cudaMemsetAsync(d_term_first_, 0, 4, st_));
    if (nt) thrust::inclusive_scan(pol, thrust::device_ptr<uint32_t>(d_epc), thrust::device_ptr<uint32_t>(d_epc + nt), thrust::device_ptr<uint32_t>(d_term_first_ + 1));
    h_dict_keys_.resize(nt); h_term_df_.resize(nt);
    SSB_CUDA_TRY(cudaMemcpyAsync(h_dict_keys_.data(), d_dict_keys_, (size_t)nt * 8, cudaMemcpyDeviceToHost, st_));
    SSB_CUDA_TRY(cudaMemcpyAsync(h_term_df_.data(), d_term_df_, (size_t)nt * 4, cudaMemcpyDeviceToHost, st_));
    SSB_CUDA_TRY(cudaStreamSynchronize(st_));
    {   // idf on the host (same libm as the oracle)
        std::vector<float> idf(nt_alloc);
        for (uint32_t t = 0; t < nt; t++) idf[t] = host_idf(n_docs, h_term_df_[t]);
        SSB_CUDA_TRY(cudaMemcpyAsync(d_term_idf_, idf.data(), (size_t)nt * 4, cudaMemcpyHostToDevice, st_));
        SSB_CUDA_TRY(cudaStreamSynchronize(st_));
    }

    // bitmaps for dense lists
    n_bitmaps_ = 0;
    if (total) {
        DevTmp<uint32_t> t_flags, t_scan, t_dense;
        SSB_CUDA_TRY(t_flags.alloc(alloc_n)); SSB_CUDA_TRY(t_scan.alloc(alloc_n));
        uint32_t *d_flags = t_flags.p, *d_scan = t_scan.p;
        mark_dense<<<(total + 255) / 256, 256, 0, st_>>>(d_e_count_, total, d_flags);
        thrust::exclusive_scan(pol, thrust::device_ptr<uint32_t>(d_flags), thrust::device_ptr<uint32_t>(d_flags + total), thrust::device_ptr<uint32_t>(d_scan));
        uint32_t last_flag = 0, last_scan = 0;
        SSB_CUDA_TRY(cudaMemcpyAsync(&last_flag, d_flags + total - 1, 4, cudaMemcpyDeviceToHost, st_));
        SSB_CUDA_TRY(cudaMemcpyAsync(&last_scan, d_scan + total - 1, 4, cudaMemcpyDeviceToHost, st_));
        SSB_CUDA_TRY(cudaStreamSynchronize(st_));
        n_bitmaps_ = last_flag + last_scan;
        assign_bitmap<<<(total + 255) / 256, 256, 0, st_>>>(d_e_count_, d_scan, total, d_e_bitmap_);
        SSB_CUDA_TRY(cudaGetLastError());
        if (n_bitmaps_) {
            SSB_CUDA_TRY(cudaMalloc(&d_bm_words_, (size_t)n_bitmaps_ * 1024 * 8));
            SSB_CUDA_TRY(cudaMalloc(&d_bm_rank_, (size_t)n_bitmaps_ * 1024 * 2));
            SSB_CUDA_TRY(t_dense.alloc(n_bitmaps_));
            uint32_t* d_dense = t_dense.p;
            compact_dense<<<(total + 255) / 256, 256, 0, st_>>>(d_e_bitmap_, total, d_dense);
            build_bitmaps<<<n_bitmaps_, 256, 0, st_>>>(d_e_bitmap_, d_e_off_, d_e_count_, total, post_.p, d_bm_words_, d_bm_rank_, d_dense);
            SSB_CUDA_TRY(cudaGetLastError());
            SSB_CUDA_TRY(cudaStreamSynchronize(st_));
        }
        SSB_CUDA_TRY(cudaStreamSynchronize(st_));
    }

    committed_ = true;   // view() is usable from here
    if (total) {
        LexView v{};
        v.e_off = d_e_off_; v.e_count = d_e_count_; v.post = post_.p; v.cache = d_cache_; v.exc_pos = exc_pos_.p; v.exc_tf = exc_tf_.p;
        v.n_exc = n_exc_; v.k1p = 1.2f + 1.0f;
        entry_maxcomp<<<(total + 7) / 8, 256, 0, st_>>>(v, total, d_e_maxcomp_);
        SSB_CUDA_TRY(cudaGetLastError());
    }
    SSB_CUDA_TRY(cudaStreamSynchronize(st_));
    free_workspace();
    return SSB_OK;
}

int32_t LexIndex::dict_export(uint64_t* keys, uint32_t* dfs, uint64_t cap) const {
    if (!committed_) { set_error("dict_export before commit"); return SSB_E_STATE; }
    if (cap < n_terms_) { set_error("dict_export: capacity too small"); return SSB_E_INVALID; }
    if (keys) memcpy(keys, h_dict_keys_.data(), (size_t)n_terms_ * 8);
    if (dfs) memcpy(dfs, h_term_df_.data(), (size_t)n_terms_ * 4);
    return SSB_OK;
}

int32_t LexIndex::set_global_df(const uint64_t* keys, const uint32_t* dfs, uint64_t n) {
    if (!committed_) { set_error("set_global_df before commit"); return SSB_E_STATE; }
    std::vector<float> idf(n_terms_ ? n_terms_ : 1);
    for (uint64_t i = 0; i < n; i++) {
        size_t lo = 0, hi = n_terms_;
        while (lo < hi) { size_t m = (lo + hi) / 2; if (h_dict_keys_[m] < keys[i]) lo = m + 1; else hi = m; }
        if (lo < n_terms_ && h_dict_keys_[lo] == keys[i]) h_term_df_[lo] = dfs[i];
    }
    for (uint32_t t = 0; t < n_terms_; t++) idf[t] = host_idf(n_docs_, h_term_df_[t]);
    SSB_CUDA_TRY(cudaMemcpyAsync(d_term_idf_, idf.data(), (size_t)n_terms_ * 4, cudaMemcpyHostToDevice, st_));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_term_df_, h_term_df_.data(), (size_t)n_terms_ * 4, cudaMemcpyHostToDevice, st_));
    SSB_CUDA_TRY(cudaStreamSynchronize(st_));
    return SSB_OK;
}

int32_t LexIndex::ensure_workspace(uint32_t nq, uint32_t total_terms) {
    const uint32_t nlv = (uint32_t)levels_.size();
    if (nq <= ws_nq_ && total_terms <= ws_terms_ && nlv == ws_levels_) return SSB_OK;
    free_workspace();
    uint32_t cq = nq > max_batch_ ? nq : max_batch_;
    uint32_t ct = total_terms > cq * 4 ? total_terms : cq * 4;
    SSB_CUDA_TRY(cudaMalloc(&d_plans_, (size_t)cq * sizeof(QueryPlan)));
    SSB_CUDA_TRY(cudaMalloc(&d_items_, (size_t)cq * (nlv ? nlv : 1) * 8));
    SSB_CUDA_TRY(cudaMalloc(&d_item_ent_, (size_t)cq * (nlv ? nlv : 1) * sizeof(uint2)));
    SSB_CUDA_TRY(cudaMalloc(&d_theta_, (size_t)cq * 8)); SSB_CUDA_TRY(cudaMalloc(&d_lock_, (size_t)cq * 4));
    SSB_CUDA_TRY(cudaMalloc(&d_count_, (size_t)cq * 8)); SSB_CUDA_TRY(cudaMalloc(&d_ctr_, 16));
    SSB_CUDA_TRY(cudaMalloc(&d_qoff_, ((size_t)cq + 1) * 4)); SSB_CUDA_TRY(cudaMalloc(&d_qkeys_, (size_t)ct * 8));
    SSB_CUDA_TRY(cudaMalloc(&d_stats_, sizeof(LexStats)));
    SSB_CUDA_TRY(cudaMemsetAsync(d_stats_, 0, sizeof(LexStats), st_));
    ws_nq_ = cq; ws_terms_ = ct; ws_levels_ = nlv;
    return SSB_OK;
}

int32_t LexIndex::search_keys(const ssb_lex_batch* q, uint32_t k, uint32_t result_type, uint64_t* keys_out_dev,
                              uint64_t* count_dev, uint64_t* launches, const uint64_t* ceil_dev) {
    if (!committed_) { set_error("search before ssb_lexical_commit"); return SSB_E_STATE; }
    if (!q || (q->n_queries && (!q->term_offsets || !keys_out_dev))) { set_error("search_lexical: null argument"); return SSB_E_INVALID; }
    if (k > SSB_K_MAX) { set_error("k=%u exceeds SSB_K_MAX=%u", k, SSB_K_MAX); return SSB_E_UNSUPPORTED; }
    if (result_type > SSB_RESULT_TOPKCOUNT || q->query_type > SSB_QUERY_INTERSECTION) { set_error("bad result_type/query_type"); return SSB_E_INVALID; }
    if (result_type != SSB_RESULT_COUNT && k == 0) result_type = SSB_RESULT_COUNT;   // search.rs:2472-2478
    const uint32_t nq = q->n_queries;
    if (nq == 0) return SSB_OK;
    uint32_t total_terms = 0;
    const bool off_dev = is_device_ptr(q->term_offsets);
    if (off_dev) SSB_CUDA_TRY(cudaMemcpy(&total_terms, q->term_offsets + nq, 4, cudaMemcpyDeviceToHost));
    else {
        total_terms = q->term_offsets[nq];
        for (uint32_t i = 0; i < nq; i++)
            if (q->term_offsets[i + 1] < q->term_offsets[i] || q->term_offsets[i + 1] - q->term_offsets[i] > SSB_MAX_QUERY_TERMS) {
                set_error("query %u has %u terms (max %u unique terms per query)", i, q->term_offsets[i + 1] - q->term_offsets[i], SSB_MAX_QUERY_TERMS);
                return SSB_E_UNSUPPORTED;
            }
    }
    if ((uint64_t)nq * (levels_.size() ? levels_.size() : 1) >= 0xFFFFFFFFull) { set_error("batch too large: n_queries * n_levels must be < 2^32"); return SSB_E_UNSUPPORTED; }
    SSB_TRY(ensure_workspace(nq, total_terms));
    SSB_CUDA_TRY(to_device(d_qoff_, q->term_offsets, ((size_t)nq + 1) * 4, st_));
    SSB_CUDA_TRY(to_device(d_qkeys_, q->term_keys, (size_t)total_terms * 8, st_));
    SSB_CUDA_TRY(cudaMemsetAsync(d_ctr_, 0, 16, st_));
    SSB_CUDA_TRY(cudaMemsetAsync(d_stats_, 0, sizeof(LexStats), st_));

    LexView v{};
    v.dict_keys = d_dict_keys_; v.n_terms = n_terms_; v.term_first = d_term_first_; v.term_idf = d_term_idf_; v.term_df = d_term_df_;
    v.e_level = d_e_level_; v.e_off = d_e_off_; v.e_count = d_e_count_; v.e_maxcomp = d_e_maxcomp_; v.e_bitmap = d_e_bitmap_;
    v.post = post_.p; v.bm_words = d_bm_words_; v.bm_rank = d_bm_rank_; v.level_ids = d_level_ids_;
    v.n_levels = (uint32_t)levels_.size(); v.cache = d_cache_; v.exc_pos = exc_pos_.p; v.exc_tf = exc_tf_.p; v.n_exc = n_exc_;
    v.k1p = 1.2f + 1.0f;

    uint32_t n_pow2 = 1; while (n_pow2 < v.n_levels) n_pow2 <<= 1;
    if (n_pow2 < 2) n_pow2 = 2;
    size_t plan_smem = (size_t)v.n_levels * (8 + 2 * FAST_T) + 16 + (size_t)n_pow2 * 8;
    // glist lives in keys_out_dev's shape: use a private list buffer = d_items_-adjacent? keep separate: reuse keys_out_dev
    // directly as the global list (32 u64 per query), then mask entries >= k in copy_out.
    uint64_t* glist = keys_out_dev;
    if (plan_smem > 48 * 1024) SSB_CUDA_TRY(cudaFuncSetAttribute(lex_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem));
    lex_plan<<<nq, 128, plan_smem, st_>>>(v, d_qoff_, d_qkeys_, q->query_type, d_plans_, d_items_, (uint2*)d_item_ent_, d_ctr_, d_theta_, d_lock_, d_count_, glist, n_pow2);
    SSB_CUDA_TRY(cudaGetLastError());
    int grid = n_sms_ * SSB_LEX_MINB;
    if (ev0_) cudaEventRecord(ev0_, st_);
    lex_score<<<grid, 256, 0, st_>>>(v, d_plans_, d_items_, (const uint2*)d_item_ent_, nq, q->query_type, result_type, k ? k : 1, d_ctr_, d_theta_, d_lock_, d_count_, glist, d_stats_, ceil_dev);
    if (ev1_) cudaEventRecord(ev1_, st_);
    SSB_CUDA_TRY(cudaGetLastError());
    copy_out<<<(nq * LIST + 255) / 256, 256, 0, st_>>>(glist, d_count_, nq, result_type == SSB_RESULT_COUNT ? 0 : k, keys_out_dev, count_dev);
    SSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 3;
    return SSB_OK;
}

LexStats LexIndex::last_stats() {
    LexStats s{};
    if (d_stats_) { cudaMemcpyAsync(&s, d_stats_, sizeof(s), cudaMemcpyDeviceToHost, st_); cudaStreamSynchronize(st_); }
    return s;
}

}  // namespace ssb
