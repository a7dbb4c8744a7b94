// bm25.h — lexical (BM25) index in HBM + batched query execution.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "common.cuh"

namespace ssb {

// growable device buffer (grow = new allocation + D2D copy of the used prefix)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    int32_t reserve(size_t n, size_t used, cudaStream_t st) {
        if (n <= cap) return SSB_OK;
        size_t nc = cap ? cap : 1024;
        while (nc < n) nc += nc / 2 + 1024;
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, nc * sizeof(T));
        if (e != cudaSuccess) {
            // fall back to the exact size
            nc = n;
            e = cudaMalloc(&q, nc * sizeof(T));
            if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", nc * sizeof(T), cudaGetErrorString(e)); return SSB_E_NOMEM; }
        }
        if (used && p) {
            e = cudaMemcpyAsync(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(q); set_error("DevBuf grow copy failed: %s", cudaGetErrorString(e)); return SSB_E_CUDA; }
        }
        if (p) cudaFree(p);
        p = q; cap = nc;
        return SSB_OK;
    }
};

// scoped device temporary (freed on every exit path, including SSB_CUDA_TRY early returns)
template <typename T>
struct DevTmp {
    T* p = nullptr;
    DevTmp() = default;
    DevTmp(const DevTmp&) = delete;
    DevTmp& operator=(const DevTmp&) = delete;
    ~DevTmp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
};

struct LexLevel {
    uint32_t level_id, n_docs, n_terms;
    uint64_t post_base;            // offset of this level's postings in the arenas
    uint32_t n_post;
    uint64_t* d_term_keys;         // [n_terms]
    uint32_t* d_posting_offsets;   // [n_terms+1]
};

// device view handed to the kernels (all pointers device)
struct LexView {
    const uint64_t* dict_keys; uint32_t n_terms;
    const uint32_t* term_first;   // [n_terms+1] entry ranges
    const float* term_idf;        // [n_terms]
    const uint32_t* term_df;      // [n_terms] (global df)
    const uint32_t* e_level;      // [n_entries] local level index, ascending within a term
    const uint64_t* e_off;        // posting offset in the arenas
    const uint32_t* e_count;
    const float* e_maxcomp;       // block-max basis: max tf*(K+1)/(tf+cache[len]) over the list
    const uint32_t* e_bitmap;     // index into bm_* or 0xFFFFFFFF
    const uint32_t* post;         // arena: id16 | tf8<<16 | len8<<24, one word per posting
    const uint64_t* bm_words;     // [n_bitmaps][1024]
    const uint16_t* bm_rank;      // [n_bitmaps][1024] postings before word w
    const uint32_t* level_ids;    // [n_levels]
    uint32_t n_levels;
    const float* cache;           // [256] bm25_component_cache
    const uint64_t* exc_pos; const uint32_t* exc_tf; uint32_t n_exc;  // tf >= 255 exceptions, sorted by pos
    float k1p;                    // K + 1
};

struct QTerm { uint32_t first, n; float idf; uint32_t df; };
struct QueryPlan { QTerm t[SSB_MAX_QUERY_TERMS]; uint32_t n_live, n_items, flags, pad; };

struct LexStats { uint64_t postings_visited, probes, items_processed, items_skipped; };

class LexIndex {
public:
    explicit LexIndex(cudaStream_t st, int n_sms, uint32_t max_batch) : st_(st), n_sms_(n_sms), max_batch_(max_batch) {}
    ~LexIndex();
    int32_t add_level(const ssb_level_desc* d);
    int32_t commit(uint64_t n_docs, uint64_t len_sum);
    int32_t dict_size(uint64_t* n) const { *n = n_terms_; return SSB_OK; }
    int32_t dict_export(uint64_t* keys, uint32_t* dfs, uint64_t cap) const;
    int32_t set_global_df(const uint64_t* keys, const uint32_t* dfs, uint64_t n);
    // keys_out_dev: [n_queries][32]; count_dev: [n_queries] or null. Asynchronous on the stream.
    // ceil_dev: optional [n_queries] exclusive key ceilings (paging: only hits ranked after that key; 0 = none left)
    int32_t search_keys(const ssb_lex_batch* q, uint32_t k, uint32_t result_type, uint64_t* keys_out_dev,
                        uint64_t* count_dev, uint64_t* launches, const uint64_t* ceil_dev = nullptr);
    bool committed() const { return committed_; }
    void set_stream(cudaStream_t st) { st_ = st; }
    void set_events(cudaEvent_t a, cudaEvent_t b) { ev0_ = a; ev1_ = b; }
    LexStats last_stats();
    uint64_t n_postings() const { return n_post_; }

private:
    int32_t ensure_workspace(uint32_t nq, uint32_t total_terms);
    cudaStream_t st_;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    int n_sms_;
    uint32_t max_batch_;
    bool committed_ = false;
    std::vector<LexLevel> levels_;
    DevBuf<uint32_t> post_;
    uint64_t n_post_ = 0;
    DevBuf<uint64_t> exc_pos_; DevBuf<uint32_t> exc_tf_; uint32_t* d_exc_count_ = nullptr; uint32_t n_exc_ = 0;
    // committed structures
    uint64_t n_docs_ = 0, len_sum_ = 0;
    uint32_t n_terms_ = 0, n_entries_ = 0, n_bitmaps_ = 0;
    uint64_t* d_dict_keys_ = nullptr; uint32_t* d_term_first_ = nullptr; float* d_term_idf_ = nullptr; uint32_t* d_term_df_ = nullptr;
    uint32_t* d_e_level_ = nullptr; uint64_t* d_e_off_ = nullptr; uint32_t* d_e_count_ = nullptr; float* d_e_maxcomp_ = nullptr; uint32_t* d_e_bitmap_ = nullptr;
    uint64_t* d_bm_words_ = nullptr; uint16_t* d_bm_rank_ = nullptr;
    uint32_t* d_level_ids_ = nullptr; float* d_cache_ = nullptr;
    std::vector<uint64_t> h_dict_keys_; std::vector<uint32_t> h_term_df_;
    void free_committed();
    // workspace
    uint32_t ws_nq_ = 0, ws_terms_ = 0, ws_levels_ = 0;
    QueryPlan* d_plans_ = nullptr; uint64_t* d_items_ = nullptr; void* d_item_ent_ = nullptr; uint64_t* d_theta_ = nullptr; int* d_lock_ = nullptr;
    uint64_t* d_count_ = nullptr; uint32_t* d_ctr_ = nullptr; /* [0]=work counter [1]=max_items */
    uint32_t* d_qoff_ = nullptr; uint64_t* d_qkeys_ = nullptr; LexStats* d_stats_ = nullptr;
    void free_workspace();
};

}  // namespace ssb
