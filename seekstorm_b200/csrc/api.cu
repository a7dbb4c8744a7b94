// api.cu — the extern "C" ABI of libseekstorm_b200.so (include/seekstorm_b200.h): handle, vector index
// storage, search entry points, hybrid RRF, key decoding.  No torch types; CUDA runtime only.
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <exception>
#include <memory>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <vector>

#include "bm25.h"
#include "comm.h"
#include "common.cuh"
#include "vec_scan.h"

namespace ssb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int32_t encode_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t inner_elems, uint64_t rows,
                           uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_rows, int swizzle128) {
    return encode_tmap_2d(out, base, 4, inner_elems, rows, row_pitch_bytes, box_inner, box_rows, swizzle128 ? 128 : 0);
}

int32_t encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner_elems, uint64_t rows,
                       uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
        if (e != cudaSuccess || !p || qr != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled unavailable"); return SSB_E_CUDA; }
        fn = (EncodeTiledFn)p;
    }
    cuuint64_t gdim[2] = {inner_elems, rows};
    cuuint64_t gstr[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return SSB_E_CUDA; }
    return SSB_OK;
}

static bool dev_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace ssb

using namespace ssb;

// One search context = one CUDA stream + every per-call workspace.  Concurrent ssb_search_* calls on one handle each take a
// context from the pool ("handles are thread-safe for concurrent search_* calls — a stream/workspace pool");
// the committed index data is immutable and shared.
struct SearchCtx {
    cudaStream_t st = nullptr;       // stream this context launches on (its own, or the caller's after ssb_set_stream)
    cudaStream_t own_st = nullptr;
    LexWorkspace lex;
    DevBuf<float> qpad, qstage, qhi, qlo, q_scale, q_norm; DevBuf<int8_t> q_i8; DevBuf<int> q_aff;
    DevBuf<uint64_t> ceil, scratch, keys_a, keys_b, counts, gather;
    DevBuf<float> ivf_scores; DevBuf<uint32_t> ivf_sel; DevBuf<uint64_t> ivf_obs; std::vector<uint64_t> h_obs;   // IVF probe (vec_ivf.cu)
    std::vector<uint64_t> h_ceil, h_keys_a, h_keys_b, h_counts;
    ssb_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool ev_used = false, last_lex = false;
    const uint32_t* fb_state = nullptr;   // filter scan of the current call: device count of queries that took the exact fallback
    ~SearchCtx() {
        if (own_st) cudaStreamSynchronize(own_st);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (own_st) cudaStreamDestroy(own_st);
    }
};

constexpr size_t SSB_MAX_CTX = 16;

struct ssb_index {
    ssb_config cfg;
    int n_sms = 0;
    cudaStream_t load_st = nullptr;   // load-time stream (add_level / commit)
    // searches hold `rw` shared, index mutation exclusive (mirrors the reference's RwLock around the shard, commit.rs:142)
    std::shared_mutex rw;
    std::mutex pool_mu; std::condition_variable pool_cv;
    std::vector<std::unique_ptr<SearchCtx>> pool; std::vector<SearchCtx*> free_ctx;
    bool ext_stream_set = false; cudaStream_t ext_stream = nullptr;   // ssb_set_stream: every search runs on the caller's stream (one context)
    std::mutex stats_mu; SearchCtx* last_ctx = nullptr; ssb_stats last_stats{};
    LexIndex* lex = nullptr;
    DeleteSet del;                    // shard.delete_hashset mirrored on the device (ssb_set_deleted)
    FacetSet facets;                  // the shard's facet file as one key column per facet (ssb_set_facets)
    ShardComm comm;                   // set: this handle is one shard of a `world`-way sharded index (one process per GPU)
    // vector index
    uint32_t dims = 0, dpad = 0, dpad8 = 0;
    bool quant_i8 = false;            // ScalarQuantizationI8 / TurboQuantI8: int8 corpus, exact int32 dot products
    bool turbo = false;               // TurboQuantI8: rows and queries are sign-flipped, FWHT-rotated and quantised at tq_dim = next_pow2(dims)
    uint32_t tq_dim = 0; float* tq_mask = nullptr;   // the index's seed mask (+-1), ssb_vector_set_turboquant_mask
    bool dup_docs = false;            // some doc id occurs on more than one vector row (multi-chunk documents): results are de-duplicated
    DevBuf<float> rows;
    DevBuf<uint16_t> rows_hi, rows_lo;   // bf16 planes of `rows` (hi = bf16_rn(x), lo = bf16_rn(x - hi)): what the tcgen05 bf16 scan streams
    DevBuf<int8_t> rows_i8;
    DevBuf<float> row_scale, row_norm;   // Dot / Euclidean + ScalarQuantizationI8: per-vector scale (and norm), QuantizedVector vector_similarity.rs:1340-1371
    // Euclidean + ScalarQuantizationI8 over integer-valued 0..255 data: the AFFINE quantiser (new_scale_norm_affine, vector_similarity.rs:1414-1463)
    bool affine = false; float aff_min = 3.402823466e+38f /* f32::MAX */, aff_max = -3.402823466e+38f /* f32::MIN */;   // shard.min / max_vector_value
    DevBuf<int> row_aff;                 // int2 per row: (zero_point, dims * zero_point - sum_q)
    DevBuf<uint32_t> doc_ids;
    DevBuf<uint16_t> rows_h16;        // filter scan: fp16 plane half_rn(rows * vec_scale)
    DevBuf<uint32_t> vec_err;         // filter scan: {max_r |a_r*scale - h_r|, max_r |h_r|, scratch} as f32 bits (launch_rows_f16_err)
    float vec_scale = 0.f;            // power of two; 0 = not chosen yet (first add_level)
    // IVF cluster tables (vector.rs:1066-1094; f32 indexes): one entry per add call ("level"), clusters numbered across levels
    DevBuf<float> medoids; DevBuf<uint32_t> row_cluster, cl_count, lvl_begin;
    std::vector<uint32_t> h_lvl_begin; uint32_t n_clusters = 0, max_level_clusters = 0;
    uint64_t n_rows = 0;
};

namespace {

// RAII lease of a search context
struct CtxLease {
    ssb_index* ix; SearchCtx* c = nullptr;
    explicit CtxLease(ssb_index* i) : ix(i) {}
    // block = false: give up (c stays null, SSB_OK) instead of waiting when every context is busy
    int32_t acquire(bool block = true) {
        std::unique_lock<std::mutex> g(ix->pool_mu);
        // one context when the caller owns the stream, and when searches contain collectives (every rank must issue them in the
        // same order: concurrent searches on one communicator would interleave them)
        const size_t cap = (ix->ext_stream_set || ix->comm.active()) ? 1 : SSB_MAX_CTX;
        for (;;) {
            if (!ix->free_ctx.empty()) { c = ix->free_ctx.back(); ix->free_ctx.pop_back(); break; }
            if (ix->pool.size() < cap) {
                std::unique_ptr<SearchCtx> n(new (std::nothrow) SearchCtx());
                if (!n) { set_error("out of host memory"); return SSB_E_NOMEM; }
                if (cudaStreamCreateWithFlags(&n->own_st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); set_error("stream create failed"); return SSB_E_CUDA; }
                cudaEventCreate(&n->ev0); cudaEventCreate(&n->ev1);
                n->lex.ev0 = n->ev0; n->lex.ev1 = n->ev1;
                c = n.get(); ix->pool.push_back(std::move(n));
                break;
            }
            if (!block) return SSB_OK;
            ix->pool_cv.wait(g);
        }
        c->st = ix->ext_stream_set ? ix->ext_stream : c->own_st;
        c->stats = ssb_stats{}; c->ev_used = false; c->last_lex = false; c->fb_state = nullptr;
        return SSB_OK;
    }
    ~CtxLease() {
        if (!c) return;
        { std::lock_guard<std::mutex> g(ix->stats_mu); ix->last_ctx = c; ix->last_stats = c->stats; }
        { std::lock_guard<std::mutex> g(ix->pool_mu); ix->free_ctx.push_back(c); }
        ix->pool_cv.notify_one();
    }
};

// every extern "C" body runs inside this guard: no exception (thrust::system_error, std::bad_alloc, ...) crosses the C boundary
#define SSB_API_BEGIN try {
#define SSB_API_END                                                                                              \
    } catch (const std::bad_alloc&) { cudaGetLastError(); set_error("out of memory (host or device)"); return SSB_E_NOMEM; \
    } catch (const std::exception& e) { cudaGetLastError(); set_error("internal error: %s", e.what()); return SSB_E_CUDA;   \
    } catch (...) { cudaGetLastError(); set_error("internal error: unknown exception"); return SSB_E_CUDA; }

}  // namespace

namespace {

struct IvfQuery { uint32_t mode, n_probe; float thr; };   // AnnMode of one call (thr pre-mapped, vector.rs:388-399)

int32_t vec_keys(ssb_index* ix, SearchCtx& c, const void* queries, bool queries_i8, uint32_t nq, uint32_t k, uint64_t* keys_out_dev /*[nq][32]*/,
                 const uint64_t* ceil_dev = nullptr /*[>= nq_pad] paging ceilings*/, const IvfQuery* ivf = nullptr /*null = AnnMode::All*/) {
    if (ivf && (ix->quant_i8 || !ix->medoids.p)) { set_error("AnnMode other than All needs an f32 vector index"); return SSB_E_UNSUPPORTED; }
    if (ix->dims == 0) { set_error("no vector index configured (vector_dims = 0)"); return SSB_E_STATE; }
    if (k == 0 || k > SSB_K_MAX) { set_error("k must be in 1..%u", SSB_K_MAX); return SSB_E_UNSUPPORTED; }
    if (queries_i8 && !ix->quant_i8) { set_error("int8 queries need a ScalarQuantizationI8 index"); return SSB_E_INVALID; }
    if (nq == 0) return SSB_OK;
    // AUTO (measured, 1M x 768): one FP32 pass of 16 queries takes ~0.5 ms, one tensor-core pass of up to 128 queries
    // ~0.6 ms -> FP32 scan for <= 16 queries, tensor-core scan above.  Euclidean always takes the FP32 scan.
    uint32_t kern = ix->cfg.vector_kernel;
    if (ix->quant_i8) kern = SSB_VEC_KERNEL_TCGEN05;   // one kernel for the int8 corpus: tcgen05 kind::i8, 128-query tile
    if (kern == SSB_VEC_KERNEL_AUTO) {
        // tensor-core scan above 16 queries; the 256-query tile (0.95 ms per pass vs 0.55 ms for 128 queries, measured) when it
        // needs fewer milliseconds for this batch: ceil(nq/256) * 0.95 < ceil(nq/128) * 0.55
        const uint32_t p128 = (nq + 127u) / 128u, p256 = (nq + 255u) / 256u;
        kern = nq <= 16 ? SSB_VEC_KERNEL_FFMA : (p256 * 95u < p128 * 55u ? SSB_VEC_KERNEL_TCGEN05_BF16_N256 : SSB_VEC_KERNEL_TCGEN05_BF16);
        // filter scan + exact refine: half the bytes and a third of the tensor work per pass
        // — at every batch size: one 128-query filter pass (0.25 ms on 1M x 768) also beats the FP32 scan's 0.5 ms pass for <= 16 queries
        const uint32_t exact_kern = kern;
        kern = nq <= 128 ? SSB_VEC_KERNEL_TCGEN05_FILTER : SSB_VEC_KERNEL_TCGEN05_FILTER_N256_PAIR;   // above 128 queries: 256 per pass on CTA pairs
        if (ix->quant_i8 || ix->cfg.vector_similarity == SSB_SIM_EUCLIDEAN || k > 16 || ceil_dev || !ix->rows_h16.p || !ix->vec_err.p) kern = exact_kern;
    }
    // the filter scan keeps a candidate set sized for k <= 16 in the 32-entry lists and has no paging (ceilings are exact keys): those
    // calls take the exact 3-product scan
    bool filter = (kern == SSB_VEC_KERNEL_TCGEN05_FILTER || kern == SSB_VEC_KERNEL_TCGEN05_FILTER_N256 || kern == SSB_VEC_KERNEL_TCGEN05_FILTER_N256_PAIR);
    if (filter && (ix->quant_i8 || ix->cfg.vector_similarity == SSB_SIM_EUCLIDEAN || k > 16 || ceil_dev || !ix->rows_h16.p || !ix->vec_err.p)) {
        kern = kern != SSB_VEC_KERNEL_TCGEN05_FILTER && (nq + 255u) / 256u * 95u < (nq + 127u) / 128u * 55u ? SSB_VEC_KERNEL_TCGEN05_BF16_N256 : SSB_VEC_KERNEL_TCGEN05_BF16;
        filter = false;
        if (ix->quant_i8) kern = SSB_VEC_KERNEL_TCGEN05;
    }
    const bool use_tc = ix->quant_i8 || (kern >= SSB_VEC_KERNEL_TCGEN05 && ix->cfg.vector_similarity != SSB_SIM_EUCLIDEAN);   // the int8 index is always scanned on the tensor cores
    const bool tc_bf16 = filter || kern == SSB_VEC_KERNEL_TCGEN05_BF16 || kern == SSB_VEC_KERNEL_TCGEN05_BF16_N64 || kern == SSB_VEC_KERNEL_TCGEN05_BF16_N256;
    const uint32_t qt = !use_tc ? vec::VEC_QT : (ix->quant_i8 ? 128u : (kern == SSB_VEC_KERNEL_TCGEN05_N64 || kern == SSB_VEC_KERNEL_TCGEN05_BF16_N64) ? 64u : ((kern == SSB_VEC_KERNEL_TCGEN05_BF16_N256 || kern == SSB_VEC_KERNEL_TCGEN05_FILTER_N256 || kern == SSB_VEC_KERNEL_TCGEN05_FILTER_N256_PAIR) ? 256u : 128u));
    const uint32_t nq_pad = (nq + qt - 1) / qt * qt;
    cudaStream_t st = c.st;
    if (!ix->quant_i8) SSB_TRY(c.qpad.reserve((size_t)nq_pad * ix->dpad, 0, st));
    const size_t qbytes = (size_t)nq * ix->dims * (queries_i8 ? 1 : 4);
    const void* qsrc = queries;
    if (!dev_ptr(queries)) {
        SSB_TRY(c.qstage.reserve(((size_t)nq * ix->dims + 3) / (queries_i8 ? 4 : 1) + 1, 0, st));
        SSB_CUDA_TRY(cudaMemcpyAsync(c.qstage.p, queries, qbytes, cudaMemcpyHostToDevice, st));
        c.stats.h2d_bytes += qbytes;
        qsrc = c.qstage.p;
    }
    if (ix->quant_i8) {
        SSB_TRY(c.q_i8.reserve((size_t)nq_pad * ix->dpad8, 0, st));
        if (queries_i8 && (ix->turbo || ix->cfg.vector_similarity != SSB_SIM_COSINE)) { set_error("int8 query codes are accepted for Cosine + ScalarQuantizationI8 only (the other quantisers need the query scale)"); return SSB_E_UNSUPPORTED; }
        if (ix->turbo) {
            // the query goes through the same TurboQuant as the rows (search.rs:1545-1556, 1592-1602).  Dot / Cosine: the reference's score is
            // -(dot * query_scale * row_scale) (vector_similarity.rs:161-176): the NEGATED query scale through the scaled epilogue is exactly that
            SSB_TRY(c.q_scale.reserve(nq_pad, 0, st)); SSB_TRY(c.q_norm.reserve(nq_pad, 0, st));
            SSB_TRY(vec::launch_quantize_rows_turbo_i8((const float*)qsrc, ix->dims, nq, nq_pad, ix->dims, ix->tq_dim, ix->tq_mask, c.q_i8.p, ix->dpad8, c.q_scale.p,
                                                       c.q_norm.p, ix->cfg.vector_similarity == SSB_SIM_COSINE, ix->cfg.vector_similarity != SSB_SIM_EUCLIDEAN, st));
        } else if (ix->affine) {
            // affine Euclidean: the query is quantised with a COPY of the shard's (min, max) state (search.rs:1514-1530, 1562-1580)
            SSB_TRY(c.q_scale.reserve(nq_pad, 0, st)); SSB_TRY(c.q_norm.reserve(nq_pad, 0, st)); SSB_TRY(c.q_aff.reserve((size_t)nq_pad * 2, 0, st));
            SSB_TRY(vec::launch_quantize_rows_affine_i8((const float*)qsrc, ix->dims, nq, nq_pad, ix->dims, nullptr, nullptr, ix->aff_min, ix->aff_max, c.q_i8.p, ix->dpad8,
                                                        c.q_scale.p, c.q_norm.p, c.q_aff.p, 1, st));
        } else if (ix->cfg.vector_similarity != SSB_SIM_COSINE) {
            // Dot / Euclidean: the query goes through the same QuantizedVector::new_scale[_norm] as the rows (search.rs:1499-1530)
            SSB_TRY(c.q_scale.reserve(nq_pad, 0, st)); SSB_TRY(c.q_norm.reserve(nq_pad, 0, st));
            SSB_TRY(vec::launch_quantize_rows_scale_i8((const float*)qsrc, ix->dims, nq, nq_pad, ix->dims, c.q_i8.p, ix->dpad8, c.q_scale.p, c.q_norm.p,
                                                       ix->cfg.vector_similarity == SSB_SIM_EUCLIDEAN, st));
        } else if (queries_i8) {
            // the caller already ran normalize + quantize_f32_to_i8 (what the reference's server holds after search.rs:1477-1490): pad only
            SSB_CUDA_TRY(cudaMemsetAsync(c.q_i8.p, 0, (size_t)nq_pad * ix->dpad8, st));
            SSB_CUDA_TRY(cudaMemcpy2DAsync(c.q_i8.p, ix->dpad8, qsrc, ix->dims, ix->dims, nq, cudaMemcpyDeviceToDevice, st));
        } else {
            // the query is normalised and quantised exactly like the corpus (search.rs:1464-1475, vector_similarity.rs:1226-1232)
            SSB_TRY(vec::launch_quantize_rows_i8((const float*)qsrc, ix->dims, nq, nq_pad, ix->dims, c.q_i8.p, ix->dpad8, st));
        }
    } else if (use_tc && tc_bf16) {
        // one launch: pad + normalise + bf16 hi/lo split (the scan reads only the split parts)
        SSB_TRY(c.qhi.reserve((size_t)nq_pad * ix->dpad, 0, st));
        SSB_TRY(c.qlo.reserve((size_t)nq_pad * ix->dpad, 0, st));
        if (filter) SSB_TRY(c.q_scale.reserve(nq_pad, 0, st));   // per-query margins 2 eps_q
        SSB_TRY(vec::launch_prep_split_queries_bf16((const float*)qsrc, nq, ix->dims, ix->dims, c.qhi.p, c.qlo.p, nq_pad, ix->dpad,
                                                    ix->cfg.vector_similarity == SSB_SIM_COSINE, st, (filter || ivf) ? c.qpad.p : nullptr,
                                                    filter ? c.q_scale.p : nullptr, filter ? ix->vec_err.p : nullptr));
    } else
    SSB_TRY(vec::launch_prep_queries((const float*)qsrc, nq, ix->dims, ix->dims, c.qpad.p, nq_pad, ix->dpad,
                                     ix->cfg.vector_similarity == SSB_SIM_COSINE, st));
    c.stats.kernel_launches += 1;
    if (ix->n_rows == 0) { SSB_CUDA_TRY(cudaMemsetAsync(keys_out_dev, 0, (size_t)nq * LIST * 8, st)); return SSB_OK; }
    size_t sb = use_tc ? vec::scan_tc_scratch_bytes(ix->n_sms, nq_pad) : vec::scan_scratch_bytes(ix->n_sms, nq_pad);
    const size_t head_words = sb / 8 + (size_t)nq_pad * LIST + (nq_pad + 1) / 2;
    SSB_TRY(c.scratch.reserve(head_words + (filter ? vec::refine_scratch_words(ix->n_sms, nq_pad) : 0), 0, st));
    vec::ScanArgs a{};
    a.rows = ix->rows.p; a.rows_hi = ix->rows_hi.p; a.rows_lo = ix->rows_lo.p; a.rows_h16 = ix->rows_h16.p; a.doc_ids = ix->doc_ids.p; a.n_rows = ix->n_rows; a.dpad = ix->dpad; a.queries_padded = c.qpad.p;
    a.nq_pad = nq_pad; a.nq_valid = nq; a.k = k; a.similarity = ix->cfg.vector_similarity; a.n_sms = ix->n_sms;
    a.scratch = c.scratch.p; a.scratch_bytes = sb;
    uint64_t* merged = c.scratch.p + sb / 8;   // [nq_pad][32]
    a.keys_out = merged; a.ev0 = c.ev0; a.ev1 = c.ev1; c.ev_used = true;
    a.thr_buf = reinterpret_cast<uint32_t*>(merged + (size_t)nq_pad * LIST);
    a.ceil_keys = ceil_dev;
    if (ix->del.n) { a.del_slot = ix->del.d_slot; a.del_words = ix->del.d_words; }
    a.launches = &c.stats.kernel_launches;
    if (ivf) {
        // cluster probe: medoid scores, per-(query, level) selection -> one bit per (query, cluster); the scans test it per candidate
        vec::IvfArgs v{};
        v.medoids = ix->medoids.p; v.lvl_begin = ix->lvl_begin.p; v.cl_count = ix->cl_count.p;
        v.n_clusters = ix->n_clusters; v.n_levels = (uint32_t)ix->h_lvl_begin.size(); v.max_level_clusters = ix->max_level_clusters;
        v.queries_padded = c.qpad.p; v.nq = nq; v.nq_pad = nq_pad; v.dpad = ix->dpad; v.similarity = ix->cfg.vector_similarity;
        v.ann_mode = ivf->mode; v.n_probe = ivf->n_probe; v.cluster_threshold = ivf->thr;
        v.words = (ix->n_clusters + 31) / 32;
        SSB_TRY(c.ivf_scores.reserve((size_t)nq * ix->n_clusters, 0, st));
        SSB_TRY(c.ivf_sel.reserve((size_t)nq_pad * v.words, 0, st));
        SSB_TRY(c.ivf_obs.reserve(nq, 0, st));
        v.scores = c.ivf_scores.p; v.sel = c.ivf_sel.p; v.observed = c.ivf_obs.p; v.launches = &c.stats.kernel_launches;
        SSB_TRY(vec::launch_ivf_select(v, st));
        a.ivf_sel = c.ivf_sel.p; a.ivf_words = v.words; a.row_cluster = ix->row_cluster.p;
    }
    if (ix->quant_i8) {
        a.rows_i8 = ix->rows_i8.p; a.queries_i8 = c.q_i8.p; a.dpad8 = ix->dpad8;
        if (ix->turbo || ix->cfg.vector_similarity != SSB_SIM_COSINE) {
            a.i8_scaled = ix->cfg.vector_similarity == SSB_SIM_EUCLIDEAN ? (ix->affine ? 3 : 2) : 1;
            a.row_aff = ix->row_aff.p; a.q_aff = c.q_aff.p;
            a.row_scale = ix->row_scale.p; a.row_norm = ix->row_norm.p; a.q_scale = c.q_scale.p; a.q_norm = c.q_norm.p;
        }
        SSB_TRY(vec::launch_scan_tc(a, 128, 2, st));
    } else if (use_tc) {
        SSB_TRY(c.qhi.reserve((size_t)nq_pad * ix->dpad, 0, st));
        SSB_TRY(c.qlo.reserve((size_t)nq_pad * ix->dpad, 0, st));
        a.q_hi = c.qhi.p; a.q_lo = c.qlo.p;
        if (filter) a.q_scale = c.q_scale.p;
        SSB_TRY(vec::launch_scan_tc(a, qt, filter ? (kern == SSB_VEC_KERNEL_TCGEN05_FILTER_N256_PAIR ? 4 : 3) : (tc_bf16 ? 1 : 0), st));
        if (filter) {
            vec::RefineArgs r{};
            r.rows = ix->rows.p; r.doc_ids = ix->doc_ids.p; r.n_rows = ix->n_rows; r.dpad = ix->dpad; r.queries_padded = c.qpad.p; r.margin = c.q_scale.p;
            r.keys = merged; r.keys_out = keys_out_dev;   // the refine step writes the caller's buffer directly
            r.nq = nq; r.nq_pad = nq_pad; r.k = k;
            r.fb_lists = c.scratch.p + head_words;
            r.fb_state = reinterpret_cast<uint32_t*>(r.fb_lists + (size_t)nq_pad * ix->n_sms * LIST);
            r.del_slot = a.del_slot; r.del_words = a.del_words; r.ivf_sel = a.ivf_sel; r.ivf_words = a.ivf_words; r.row_cluster = a.row_cluster; r.n_sms = ix->n_sms; r.launches = &c.stats.kernel_launches;
            SSB_TRY(vec::launch_refine(r, st));
            c.fb_state = r.fb_state;
        }
    } else {
        SSB_TRY(vec::launch_scan_ffma(a, st));
    }
    if (!filter) SSB_CUDA_TRY(cudaMemcpyAsync(keys_out_dev, merged, (size_t)nq * LIST * 8, cudaMemcpyDeviceToDevice, st));
    c.stats.algorithmic_bytes += (uint64_t)(nq_pad / qt) * ix->n_rows * ix->dims * (ix->quant_i8 ? 1 : 4);
    // bytes the scan kernel actually streams per call: the filter scan reads the 2-byte hi plane, the refine step <= 32 f32 rows per query
    c.stats.scan_bytes_read += filter ? (uint64_t)(nq_pad / qt) * ix->n_rows * ix->dims * 2 + (uint64_t)nq * LIST * ix->dims * 4
                                      : (uint64_t)(nq_pad / qt) * ix->n_rows * ix->dims * (ix->quant_i8 ? 1 : 4);
    return SSB_OK;
}

// keys of one query -> hits; `dedup`: keep only the best-scoring row of a doc id (TopK::push, vector.rs:436-470)
uint32_t decode_list(const uint64_t* keys, uint32_t k, ssb_hit* hits, bool dedup) {
    uint32_t n = 0;
    for (uint32_t j = 0; j < LIST && n < k; j++) {
        const uint64_t key = keys[j];
        if (!key) break;
        const uint64_t doc = key_doc(key);
        if (dedup) { bool seen = false; for (uint32_t i = 0; i < n; i++) seen = seen || hits[i].doc_id == doc; if (seen) continue; }
        hits[n].doc_id = doc; hits[n].score = key_score(key); hits[n].pad = 0;
        n++;
    }
    return n;
}

void decode_keys(const uint64_t* keys, uint32_t nq, uint32_t k, ssb_hit* hits, uint32_t* n_hits, bool dedup = false) {
    for (uint32_t q = 0; q < nq; q++) {
        const uint32_t n = decode_list(keys + (size_t)q * LIST, k, hits + (size_t)q * k, dedup);
        for (uint32_t j = n; j < k; j++) { hits[(size_t)q * k + j].doc_id = 0; hits[(size_t)q * k + j].score = 0.f; hits[(size_t)q * k + j].pad = 0; }
        if (n_hits) n_hits[q] = n;
    }
}

// Sharded index: all-gather every rank's [nq][32] key lists and merge them (G*k -> k with the canonical tie rule; the reference
// concatenates the shard results and sorts, search.rs:1875-1928, 2097-2106).  In place; identical result on every rank.
int32_t shard_merge(ssb_index* ix, SearchCtx& c, uint64_t* keys_dev, uint32_t nq) {
    if (!ix->comm.active() || nq == 0) return SSB_OK;
    const size_t n = (size_t)nq * LIST;
    SSB_TRY(c.gather.reserve(n * ix->comm.world, 0, c.st));
    SSB_TRY(comm_all_gather_u64(ix->comm, keys_dev, c.gather.p, n, c.st));
    SSB_TRY(vec::launch_merge_lists(c.gather.p, ix->comm.world, nq, keys_dev, c.st));
    c.stats.kernel_launches += 2;
    return SSB_OK;
}
int32_t shard_sum_counts(ssb_index* ix, SearchCtx& c, uint64_t* counts_dev, uint32_t nq) {
    if (!ix->comm.active() || nq == 0 || !counts_dev) return SSB_OK;
    SSB_TRY(comm_all_reduce_sum_u64(ix->comm, counts_dev, nq, c.st));   // result_count_total = sum over shards (search.rs:1875-1940)
    c.stats.kernel_launches += 1;
    return SSB_OK;
}

// Paging state of the host-facing search calls (k > SSB_K_MAX, or de-duplication of multi-chunk documents).
struct PageState {
    SearchCtx& c; uint32_t nq, k; ssb_hit* hits; uint32_t* n_hits; std::vector<uint32_t> cnt; std::vector<uint8_t> open; bool dedup;
    PageState(SearchCtx& c_, uint32_t nq_, uint32_t k_, ssb_hit* h, uint32_t* n, bool dedup_ = false)
        : c(c_), nq(nq_), k(k_), hits(h), n_hits(n), cnt(nq_, 0), open(nq_, 1), dedup(dedup_) {
        c.h_ceil.assign((size_t)nq_ + 256, 0);
    }
    // consume one [nq][32] page that was fetched with `kk` results per query; returns true if any query wants another page
    bool append(const uint64_t* keys, uint32_t kk) {
        bool more = false;
        for (uint32_t q = 0; q < nq; q++) {
            if (!open[q]) { c.h_ceil[q] = 0; continue; }                    // exhausted or complete on an earlier page
            uint32_t seen = 0; uint64_t last = 0;
            for (uint32_t j = 0; j < kk && cnt[q] < k; j++) {
                const uint64_t key = keys[(size_t)q * LIST + j];
                if (!key) break;
                seen++; last = key;
                const uint64_t doc = key_doc(key);
                if (dedup) {
                    bool dup = false;
                    for (uint32_t i = 0; i < cnt[q]; i++) dup = dup || hits[(size_t)q * k + i].doc_id == doc;
                    if (dup) continue;
                }
                ssb_hit& h = hits[(size_t)q * k + cnt[q]];
                h.doc_id = doc; h.score = key_score(key); h.pad = 0;
                cnt[q]++;
            }
            // another page only if this one was full (else the list is exhausted) and the query still lacks results
            open[q] = seen == kk && cnt[q] < k;
            c.h_ceil[q] = open[q] ? last : 0;                                // 0 = nothing left below
            more = more || open[q];
        }
        return more;
    }
    uint32_t next_page_k() const {   // results to fetch per query on the next page
        if (dedup) return SSB_K_MAX;
        uint32_t need = 0;
        for (uint32_t q = 0; q < nq; q++) if (open[q]) need = std::max(need, k - cnt[q]);
        return need < SSB_K_MAX ? need : SSB_K_MAX;
    }
    int32_t upload_ceilings() {
        SSB_TRY(c.ceil.reserve((size_t)nq + 256, 0, c.st));
        SSB_CUDA_TRY(cudaMemcpyAsync(c.ceil.p, c.h_ceil.data(), ((size_t)nq + 256) * 8, cudaMemcpyHostToDevice, c.st));
        c.stats.h2d_bytes += (uint64_t)nq * 8;
        return SSB_OK;
    }
    void finish() {
        for (uint32_t q = 0; q < nq; q++) {
            for (uint32_t j = cnt[q]; j < k; j++) { ssb_hit& h = hits[(size_t)q * k + j]; h.doc_id = 0; h.score = 0.f; h.pad = 0; }
            if (n_hits) n_hits[q] = cnt[q];
        }
    }
};

inline bool hit_better(const ssb_hit& a, const ssb_hit& b) { return a.score > b.score || (a.score == b.score && a.doc_id < b.doc_id); }

void finish_stats(ssb_index* ix, SearchCtx& c) {
    if (c.ev_used) {
        float ms = 0.f;
        if (cudaEventSynchronize(c.ev1) == cudaSuccess && cudaEventElapsedTime(&ms, c.ev0, c.ev1) == cudaSuccess)
            c.stats.dominant_kernel_ns = (uint64_t)((double)ms * 1e6);
        else cudaGetLastError();
    }
}

// host-facing vector search: paging beyond 32 results, de-duplication, optional threshold
int32_t search_vector_host(ssb_index* ix, SearchCtx& c, const void* queries, bool queries_i8, uint32_t nq, uint32_t k, ssb_hit* hits, uint32_t* n_hits,
                           const IvfQuery* ivf = nullptr) {
    SSB_TRY(c.keys_a.reserve((size_t)nq * LIST, 0, c.st));
    c.h_keys_a.resize((size_t)nq * LIST);
    PageState ps(c, nq, k, hits, n_hits, ix->dup_docs);
    // the fused kernels keep 32 results per query; longer result lists are produced page by page, each page restricted to
    // keys strictly below the last key of the previous one (keys are a total order on (score desc, doc id asc))
    uint32_t kk = ix->dup_docs ? SSB_K_MAX : (k < SSB_K_MAX ? k : SSB_K_MAX);
    for (uint32_t page = 0; page < 4096; page++) {
        SSB_TRY(vec_keys(ix, c, queries, queries_i8, nq, kk, c.keys_a.p, page ? c.ceil.p : nullptr, ivf));
        SSB_TRY(shard_merge(ix, c, c.keys_a.p, nq));
        if (ivf && page == 0) {   // observed_vector_count = the vectors of the selected clusters (summed over the shards)
            SSB_TRY(shard_sum_counts(ix, c, c.ivf_obs.p, nq));
            c.h_obs.resize(nq);
            SSB_CUDA_TRY(cudaMemcpyAsync(c.h_obs.data(), c.ivf_obs.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, c.st));
        }
        SSB_CUDA_TRY(cudaMemcpyAsync(c.h_keys_a.data(), c.keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, c.st));
        uint32_t n_fb = 0;
        if (c.fb_state) SSB_CUDA_TRY(cudaMemcpyAsync(&n_fb, c.fb_state, 4, cudaMemcpyDeviceToHost, c.st));
        SSB_CUDA_TRY(cudaStreamSynchronize(c.st));
        c.stats.filter_fallbacks += n_fb;
        c.stats.d2h_bytes += (uint64_t)nq * LIST * 8;
        if (!ps.append(c.h_keys_a.data(), kk)) break;
        kk = ps.next_page_k();
        SSB_TRY(ps.upload_ceilings());
    }
    ps.finish();
    return SSB_OK;
}

}  // namespace

extern "C" {

uint32_t ssb_abi_version(void) { return SSB_ABI_VERSION; }
const char* ssb_last_error(void) { return g_err; }

int32_t ssb_create(const ssb_config* cfg, ssb_index** out) {
    SSB_API_BEGIN
    if (!cfg || !out) { set_error("ssb_create: null argument"); return SSB_E_INVALID; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device visible: libseekstorm_b200 has no CPU fallback"); return SSB_E_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("device %d out of range (%d visible)", cfg->device, ndev); return SSB_E_INVALID; }
    if (cfg->vector_similarity > SSB_SIM_EUCLIDEAN) { set_error("bad vector_similarity"); return SSB_E_INVALID; }
    if (cfg->vector_quantization > SSB_QUANT_TURBO_I8) { set_error("bad vector_quantization"); return SSB_E_INVALID; }
    if (cfg->vector_quantization == SSB_QUANT_TURBO_I8 && cfg->vector_dims > 16384) { set_error("TurboQuantI8: vector_dims above 16384 unsupported"); return SSB_E_UNSUPPORTED; }
    SSB_CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    SSB_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor); return SSB_E_UNSUPPORTED; }
    std::unique_ptr<ssb_index> ix(new (std::nothrow) ssb_index());
    if (!ix) { set_error("out of host memory"); return SSB_E_NOMEM; }
    ix->cfg = *cfg;
    if (ix->cfg.max_batch == 0) ix->cfg.max_batch = 4096;
    ix->n_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ix->load_st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); set_error("stream create failed"); return SSB_E_CUDA; }
    ix->lex = new (std::nothrow) LexIndex(ix->load_st, ix->n_sms, ix->cfg.max_batch);
    if (!ix->lex) { cudaStreamDestroy(ix->load_st); set_error("out of host memory"); return SSB_E_NOMEM; }
    ix->lex->set_deleted(&ix->del);
    ix->lex->set_facets(&ix->facets);
    ix->dims = cfg->vector_dims;
    ix->dpad = (cfg->vector_dims + 31) / 32 * 32;
    ix->dpad8 = (cfg->vector_dims + 127) / 128 * 128;
    ix->quant_i8 = cfg->vector_quantization == SSB_QUANT_SCALAR_I8 || cfg->vector_quantization == SSB_QUANT_TURBO_I8;
    ix->turbo = cfg->vector_quantization == SSB_QUANT_TURBO_I8;
    if (ix->turbo) {
        // TurboQuant::new: dim = next power of two >= vector_dims (vector_similarity.rs:1836-1859); the codes of a row span tq_dim bytes
        uint32_t d = 1; while (d < cfg->vector_dims) d <<= 1;
        ix->tq_dim = d;
        ix->dpad8 = (d + 127) / 128 * 128;
    }
    *out = ix.release();
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_destroy(ssb_index* ix) {
    SSB_API_BEGIN
    if (!ix) return SSB_OK;
    cudaSetDevice(ix->cfg.device);
    {
        std::unique_lock<std::shared_mutex> g(ix->rw);    // waits for searches in flight
        cudaStreamSynchronize(ix->load_st);
        ix->pool.clear();                                  // ~SearchCtx synchronises its stream
        delete ix->lex; ix->lex = nullptr;
        comm_destroy(ix->comm);
        ix->del.release();
        ix->facets.release();
        cudaFree(ix->tq_mask); ix->tq_mask = nullptr;
        ix->rows.release(); ix->rows_hi.release(); ix->rows_lo.release(); ix->rows_i8.release(); ix->row_scale.release(); ix->row_norm.release(); ix->row_aff.release(); ix->doc_ids.release();
        cudaStreamDestroy(ix->load_st);
    }
    delete ix;
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_lexical_add_level(ssb_index* ix, const ssb_level_desc* level) {
    SSB_API_BEGIN
    if (!ix) { set_error("null index"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (level && (dev_ptr(level->doc_ids) || dev_ptr(level->term_keys))) SSB_CUDA_TRY(cudaDeviceSynchronize());   // inputs produced on another stream
    return ix->lex->add_level(level);
    SSB_API_END
}

int32_t ssb_lexical_set_field_boosts(ssb_index* ix, uint32_t n_fields, const float* boosts) {
    SSB_API_BEGIN
    if (!ix) { set_error("null index"); return SSB_E_INVALID; }
    if (boosts && dev_ptr(boosts)) { set_error("boosts must be host memory"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    return ix->lex->set_fields(n_fields, boosts);
    SSB_API_END
}

int32_t ssb_lexical_commit(ssb_index* ix, uint64_t n_docs, uint64_t len_sum) {
    SSB_API_BEGIN
    if (!ix) { set_error("null index"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    return ix->lex->commit(n_docs, len_sum);
    SSB_API_END
}

int32_t ssb_lexical_dict_size(const ssb_index* ix, uint64_t* n) { if (!ix || !n) return SSB_E_INVALID; return ix->lex->dict_size(n); }
int32_t ssb_lexical_dict_export(const ssb_index* ix, uint64_t* keys, uint32_t* dfs, uint64_t cap) { if (!ix) return SSB_E_INVALID; return ix->lex->dict_export(keys, dfs, cap); }
int32_t ssb_lexical_set_global_df(ssb_index* ix, const uint64_t* keys, const uint32_t* dfs, uint64_t n) {
    SSB_API_BEGIN
    if (!ix || (n && (!keys || !dfs))) return SSB_E_INVALID;
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    return ix->lex->set_global_df(keys, dfs, n);
    SSB_API_END
}

// capacity hint: allocate the vector arenas for n_rows rows once instead of growing them level by level (growth = new
// allocation + device copy of everything loaded so far)
int32_t ssb_vector_reserve(ssb_index* ix, uint64_t n_rows) {
    SSB_API_BEGIN
    if (!ix) { set_error("null index"); return SSB_E_INVALID; }
    if (ix->dims == 0) { set_error("no vector index configured (vector_dims = 0)"); return SSB_E_STATE; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    cudaStream_t st = ix->load_st;
    if (n_rows <= ix->n_rows) return SSB_OK;
    SSB_TRY(ix->doc_ids.reserve(n_rows, ix->n_rows, st, true));
    if (ix->quant_i8) {
        SSB_TRY(ix->rows_i8.reserve(n_rows * ix->dpad8, ix->n_rows * ix->dpad8, st, true));
        if (ix->cfg.vector_similarity != SSB_SIM_COSINE) { SSB_TRY(ix->row_scale.reserve(n_rows, ix->n_rows, st, true)); SSB_TRY(ix->row_norm.reserve(n_rows, ix->n_rows, st, true)); }
    } else {
        SSB_TRY(ix->rows.reserve(n_rows * ix->dpad, ix->n_rows * ix->dpad, st, true));
        if (ix->cfg.vector_similarity != SSB_SIM_EUCLIDEAN) {
            SSB_TRY(ix->rows_hi.reserve(n_rows * ix->dpad, ix->n_rows * ix->dpad, st, true));
            SSB_TRY(ix->rows_lo.reserve(n_rows * ix->dpad, ix->n_rows * ix->dpad, st, true));
            SSB_TRY(ix->rows_h16.reserve(n_rows * ix->dpad, ix->n_rows * ix->dpad, st, true));
        }
    }
    return SSB_OK;
    SSB_API_END
}

}  // extern "C"

// one level (= one add call) of the vector index; cluster_counts = the level's IVF cluster table or null (one cluster).  n is only
// bounded by the loader (a level may hold one record per chunk, i.e. more than 64K).
static int32_t vector_add_level_impl(ssb_index* ix, uint32_t level_id, const float* rows, uint64_t row_stride, const uint16_t* local_ids,
                                     uint32_t n, uint32_t dims, const uint32_t* cluster_counts, uint32_t n_clusters) {
    SSB_API_BEGIN
    if (!ix || (n && !rows)) { set_error("ssb_vector_add_level: null argument"); return SSB_E_INVALID; }
    if (ix->dims == 0 || dims != ix->dims) { set_error("dims %u != configured vector_dims %u", dims, ix->dims); return SSB_E_INVALID; }
    if (cluster_counts) {
        if (dev_ptr(cluster_counts)) { set_error("cluster_counts must be host memory"); return SSB_E_INVALID; }
        uint64_t sum = 0;
        for (uint32_t c = 0; c < n_clusters; c++) { if (cluster_counts[c] == 0) { set_error("empty cluster %u", c); return SSB_E_INVALID; } sum += cluster_counts[c]; }
        if (sum != n || (n && n_clusters == 0)) { set_error("cluster table covers %llu of %u rows", (unsigned long long)sum, n); return SSB_E_INVALID; }
        if (ix->quant_i8 && n_clusters > 1) { set_error("IVF cluster tables need an f32 vector index"); return SSB_E_UNSUPPORTED; }
    }
    if (level_id >= 65536) { set_error("level_id must be < 65536 (doc id = level_id << 16 | local)"); return SSB_E_INVALID; }
    if (row_stride == 0) row_stride = dims;
    if (row_stride < dims) { set_error("row stride < dims"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (n == 0) return SSB_OK;
    cudaStream_t st = ix->load_st;
    // device-resident inputs may still be in flight on the caller's stream (the load stream is non-blocking): load time is not
    // hot, wait for the device once
    if (dev_ptr(rows) || (local_ids && dev_ptr(local_ids))) SSB_CUDA_TRY(cudaDeviceSynchronize());
    // multi-chunk documents: several rows may share a local id (one vector per chunk, vector.rs:62-73); the reference's TopK keeps
    // the best chunk per doc id (vector.rs:436-470) — remember that this index needs the de-duplicating result path
    std::vector<uint16_t> h_ids;
    if (local_ids) {
        h_ids.resize(n);
        SSB_CUDA_TRY(cudaMemcpy(h_ids.data(), local_ids, (size_t)n * 2, cudaMemcpyDefault));
        std::vector<bool> seen(65536, false);
        for (uint32_t i = 0; i < n; i++) { if (seen[h_ids[i]]) ix->dup_docs = true; seen[h_ids[i]] = true; }
    }
    SSB_TRY(ix->doc_ids.reserve(ix->n_rows + n, ix->n_rows, st));
    DevTmp<float> stage;
    if (ix->quant_i8) {
        // index-time normalise + quantise (vector.rs:585-640); the f32 rows are only staged
        SSB_TRY(ix->rows_i8.reserve((ix->n_rows + n) * ix->dpad8, ix->n_rows * ix->dpad8, st));
        SSB_CUDA_TRY(stage.alloc((size_t)n * dims));
        SSB_CUDA_TRY(cudaMemcpy2DAsync(stage.p, (size_t)dims * 4, rows, row_stride * 4, (size_t)dims * 4, n, cudaMemcpyDefault, st));
        if (ix->turbo) {
            // TurboQuant::quantize_f32_i8 for every similarity (vector.rs:684-695, 729-740), after normalize_f32 for Cosine (:585-596)
            if (!ix->tq_mask) { set_error("TurboQuantI8: call ssb_vector_set_turboquant_mask before adding vectors"); return SSB_E_STATE; }
            SSB_TRY(ix->row_scale.reserve(ix->n_rows + n, ix->n_rows, st));
            SSB_TRY(ix->row_norm.reserve(ix->n_rows + n, ix->n_rows, st));
            SSB_TRY(vec::launch_quantize_rows_turbo_i8(stage.p, dims, n, n, dims, ix->tq_dim, ix->tq_mask, ix->rows_i8.p + ix->n_rows * ix->dpad8, ix->dpad8,
                                                       ix->row_scale.p + ix->n_rows, ix->row_norm.p + ix->n_rows,
                                                       ix->cfg.vector_similarity == SSB_SIM_COSINE, 0, st));
        } else if (ix->cfg.vector_similarity == SSB_SIM_COSINE)
            SSB_TRY(vec::launch_quantize_rows_i8(stage.p, dims, n, n, dims, ix->rows_i8.p + ix->n_rows * ix->dpad8, ix->dpad8, st));
        else {
            // Dot: QuantizedVector::new_scale; Euclidean: new_scale_norm — the NON-AFFINE variant the reference picks when the first
            // vector is not all integers in 0..255 (vector.rs:651-660); integer-valued data (affine quantisation) is not built
            if (ix->cfg.vector_similarity == SSB_SIM_EUCLIDEAN && ix->n_rows == 0) {
                std::vector<float> first(dims);
                SSB_CUDA_TRY(cudaMemcpyAsync(first.data(), stage.p, (size_t)dims * 4, cudaMemcpyDeviceToHost, st));
                SSB_CUDA_TRY(cudaStreamSynchronize(st));
                bool non_affine = false;
                for (float x : first) non_affine = non_affine || x != floorf(x) || x < 0.0f || x > 255.0f;
                ix->affine = !non_affine;      // decided by the FIRST vector of the shard, for its whole life (vector.rs:657-664)
            }
            if (ix->affine) {
                // new_scale_norm_affine: every vector is quantised with the running (min, max) of everything indexed before it — the state is
                // walked on the host over the level's per-row (min, max) (64K rows), the codes are written by one more kernel
                DevTmp<float> mm, d_scale; DevTmp<int> d_zp;
                SSB_CUDA_TRY(mm.alloc((size_t)n * 2)); SSB_CUDA_TRY(d_scale.alloc(n)); SSB_CUDA_TRY(d_zp.alloc(n));
                SSB_TRY(vec::launch_rows_minmax(stage.p, dims, n, dims, mm.p, st));
                std::vector<float> h_mm((size_t)n * 2), h_scale(n); std::vector<int> h_zp(n);
                SSB_CUDA_TRY(cudaMemcpyAsync(h_mm.data(), mm.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
                SSB_CUDA_TRY(cudaStreamSynchronize(st));
                float smin = ix->aff_min, smax = ix->aff_max;
                vec::affine_walk_rows(h_mm.data(), n, &smin, &smax, h_scale.data(), h_zp.data());
                SSB_CUDA_TRY(cudaMemcpyAsync(d_scale.p, h_scale.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
                SSB_CUDA_TRY(cudaMemcpyAsync(d_zp.p, h_zp.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
                SSB_TRY(ix->row_scale.reserve(ix->n_rows + n, ix->n_rows, st));
                SSB_TRY(ix->row_norm.reserve(ix->n_rows + n, ix->n_rows, st));
                SSB_TRY(ix->row_aff.reserve((ix->n_rows + n) * 2, ix->n_rows * 2, st));
                SSB_TRY(vec::launch_quantize_rows_affine_i8(stage.p, dims, n, n, dims, d_scale.p, d_zp.p, 0.f, 0.f, ix->rows_i8.p + ix->n_rows * ix->dpad8, ix->dpad8,
                                                            ix->row_scale.p + ix->n_rows, ix->row_norm.p + ix->n_rows, ix->row_aff.p + ix->n_rows * 2, 0, st));
                SSB_CUDA_TRY(cudaStreamSynchronize(st));
                ix->aff_min = smin; ix->aff_max = smax;
            } else {
            SSB_TRY(ix->row_scale.reserve(ix->n_rows + n, ix->n_rows, st));
            SSB_TRY(ix->row_norm.reserve(ix->n_rows + n, ix->n_rows, st));
            SSB_TRY(vec::launch_quantize_rows_scale_i8(stage.p, dims, n, n, dims, ix->rows_i8.p + ix->n_rows * ix->dpad8, ix->dpad8,
                                                       ix->row_scale.p + ix->n_rows, ix->row_norm.p + ix->n_rows,
                                                       ix->cfg.vector_similarity == SSB_SIM_EUCLIDEAN, st));
            }
        }
    } else {
        SSB_TRY(ix->rows.reserve((ix->n_rows + n) * ix->dpad, ix->n_rows * ix->dpad, st));
        float* dst = ix->rows.p + ix->n_rows * ix->dpad;
        SSB_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ix->dpad * 4, rows, row_stride * 4, (size_t)dims * 4, n, cudaMemcpyDefault, st));
        SSB_TRY(vec::launch_normalize_rows(dst, n, dims, ix->dpad, ix->cfg.vector_similarity == SSB_SIM_COSINE, st));
        if (ix->cfg.vector_similarity != SSB_SIM_EUCLIDEAN) {
            // the tensor-core scan reads the corpus as two bf16 planes (same 4 bytes per element as the f32 rows, which stay for
            // the FP32 scan): split once here instead of per stage in shared memory
            SSB_TRY(ix->rows_hi.reserve((ix->n_rows + n) * ix->dpad, ix->n_rows * ix->dpad, st));
            SSB_TRY(ix->rows_lo.reserve((ix->n_rows + n) * ix->dpad, ix->n_rows * ix->dpad, st));
            SSB_TRY(vec::launch_split_rows_bf16(dst, ix->rows_hi.p + ix->n_rows * ix->dpad, ix->rows_lo.p + ix->n_rows * ix->dpad, (size_t)n * ix->dpad, st));
            // filter scan: scaled fp16 plane + its index-wide error bounds.  The scale (a power of two, fixed for the life of the index)
            // puts Cosine's unit rows below 256 and a Dot index's first level into [128, 256): later rows may be 255x larger before fp16
            // overflows — an overflowing row makes the error bound infinite and every filter query takes the exact fallback.
            if (!ix->vec_err.p) { SSB_TRY(ix->vec_err.reserve(4, 0, st, true)); SSB_CUDA_TRY(cudaMemsetAsync(ix->vec_err.p, 0, 16, st)); }
            if (ix->vec_scale == 0.f) {
                if (ix->cfg.vector_similarity == SSB_SIM_COSINE) ix->vec_scale = 256.f;
                else {
                    uint32_t bits = 0;
                    SSB_TRY(vec::launch_max_abs_f32(dst, (size_t)n * ix->dpad, ix->vec_err.p + 2, st));
                    SSB_CUDA_TRY(cudaMemcpyAsync(&bits, ix->vec_err.p + 2, 4, cudaMemcpyDeviceToHost, st));
                    SSB_CUDA_TRY(cudaStreamSynchronize(st));
                    float mx; memcpy(&mx, &bits, 4);
                    ix->vec_scale = mx > 0.f ? exp2f((float)(7 - ilogbf(mx))) : 1.f;
                }
            }
            SSB_TRY(ix->rows_h16.reserve((ix->n_rows + n) * ix->dpad, ix->n_rows * ix->dpad, st));
            SSB_TRY(vec::launch_rows_f16_err(dst, ix->rows_h16.p + ix->n_rows * ix->dpad, n, ix->dpad, ix->vec_scale, ix->vec_err.p, st));
        }
    }
    DevTmp<uint16_t> tmp;
    const uint16_t* lid = local_ids;
    if (local_ids && !dev_ptr(local_ids)) {
        SSB_CUDA_TRY(tmp.alloc(n));
        SSB_CUDA_TRY(cudaMemcpyAsync(tmp.p, h_ids.data(), (size_t)n * 2, cudaMemcpyHostToDevice, st));
        lid = tmp.p;
    }
    SSB_TRY(vec::launch_fill_doc_ids(ix->doc_ids.p + ix->n_rows, lid, level_id, n, st));
    if (!ix->quant_i8) {
        // IVF tables: clusters are numbered across levels; a cluster's medoid is its first row (vector.rs:1316-1320)
        const uint32_t one = n;
        if (!cluster_counts) { cluster_counts = &one; n_clusters = 1; }
        std::vector<uint32_t> rc(n), mrow(n_clusters);
        uint32_t r = 0;
        for (uint32_t c = 0; c < n_clusters; c++) {
            mrow[c] = (uint32_t)ix->n_rows + r;
            for (uint32_t i = 0; i < cluster_counts[c]; i++) rc[r++] = ix->n_clusters + c;
        }
        SSB_TRY(ix->row_cluster.reserve(ix->n_rows + n, ix->n_rows, st));
        SSB_TRY(ix->cl_count.reserve(ix->n_clusters + n_clusters, ix->n_clusters, st));
        SSB_TRY(ix->medoids.reserve((size_t)(ix->n_clusters + n_clusters) * ix->dpad, (size_t)ix->n_clusters * ix->dpad, st));
        SSB_TRY(ix->lvl_begin.reserve(ix->h_lvl_begin.size() + 2, 0, st));
        DevTmp<uint32_t> midx; SSB_CUDA_TRY(midx.alloc(n_clusters));
        SSB_CUDA_TRY(cudaMemcpyAsync(ix->row_cluster.p + ix->n_rows, rc.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        SSB_CUDA_TRY(cudaMemcpyAsync(ix->cl_count.p + ix->n_clusters, cluster_counts, (size_t)n_clusters * 4, cudaMemcpyHostToDevice, st));
        SSB_CUDA_TRY(cudaMemcpyAsync(midx.p, mrow.data(), (size_t)n_clusters * 4, cudaMemcpyHostToDevice, st));
        SSB_TRY(vec::launch_gather_rows(ix->rows.p, midx.p, n_clusters, ix->dpad, ix->medoids.p + (size_t)ix->n_clusters * ix->dpad, st));
        std::vector<uint32_t> lb = ix->h_lvl_begin; lb.push_back(ix->n_clusters); lb.push_back(ix->n_clusters + n_clusters);
        SSB_CUDA_TRY(cudaMemcpyAsync(ix->lvl_begin.p, lb.data(), lb.size() * 4, cudaMemcpyHostToDevice, st));
        SSB_CUDA_TRY(cudaStreamSynchronize(st));
        ix->h_lvl_begin.push_back(ix->n_clusters);
        ix->n_clusters += n_clusters;
        if (n_clusters > ix->max_level_clusters) ix->max_level_clusters = n_clusters;
    }
    SSB_CUDA_TRY(cudaStreamSynchronize(st));
    ix->n_rows += n;
    return SSB_OK;
    SSB_API_END
}

extern "C" {

int32_t ssb_vector_add_level(ssb_index* ix, uint32_t level_id, const float* rows, uint64_t row_stride, const uint16_t* local_ids,
                             uint32_t n, uint32_t dims) {
    if (n > 65536) { set_error("a level holds at most 65536 vectors"); return SSB_E_INVALID; }
    return vector_add_level_impl(ix, level_id, rows, row_stride, local_ids, n, dims, nullptr, 0);
}

int32_t ssb_vector_add_level_clustered(ssb_index* ix, uint32_t level_id, const float* rows, uint64_t row_stride, const uint16_t* local_ids,
                                       uint32_t n, uint32_t dims, const uint32_t* cluster_counts, uint32_t n_clusters) {
    if (n > 65536) { set_error("a level holds at most 65536 vectors"); return SSB_E_INVALID; }
    if (n && !cluster_counts) { set_error("ssb_vector_add_level_clustered: null cluster table"); return SSB_E_INVALID; }
    return vector_add_level_impl(ix, level_id, rows, row_stride, local_ids, n, dims, n ? cluster_counts : nullptr, n_clusters);
}

int32_t ssb_load_index_bin(ssb_index* ix, const void* bytes, uint64_t len, const ssb_index_bin_params* params, uint64_t* n_docs_out) {
    SSB_API_BEGIN
    if (!ix || !bytes || !params) { set_error("ssb_load_index_bin: null argument"); return SSB_E_INVALID; }
    if (dev_ptr(bytes)) { set_error("ssb_load_index_bin: bytes must be host memory"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    return load_index_bin(ix->lex, (const uint8_t*)bytes, len, params, n_docs_out);
    SSB_API_END
}

int32_t ssb_index_bin_inspect(const void* bytes, uint64_t len, const ssb_index_bin_params* params, uint64_t out[8]) {
    SSB_API_BEGIN
    if (!bytes || !params || !out) { set_error("ssb_index_bin_inspect: null argument"); return SSB_E_INVALID; }
    return inspect_index_bin((const uint8_t*)bytes, len, params, out);
    SSB_API_END
}

int32_t ssb_load_vector_bin(ssb_index* ix, const void* bytes, uint64_t len, uint64_t* n_vectors_out) {
    SSB_API_BEGIN
    if (!ix || !bytes) { set_error("ssb_load_vector_bin: null argument"); return SSB_E_INVALID; }
    if (dev_ptr(bytes)) { set_error("ssb_load_vector_bin: bytes must be host memory"); return SSB_E_INVALID; }
    if (ix->dims == 0) { set_error("ssb_load_vector_bin: the index has no vector_dims"); return SSB_E_STATE; }
    std::vector<VectorLevel> levels;
    SSB_TRY(parse_vector_bin((const uint8_t*)bytes, len, ix->dims, levels));
    uint64_t total = 0;
    for (auto& vl : levels) {
        // a level may hold more than 64K records (one per chunk); its cluster table (IVF, vector.rs:1066-1094) rides along.  Empty clusters
        // cannot be probed (their medoid would be another cluster's record): such a table is dropped, the level becomes one cluster.
        bool ok = !vl.cluster_counts.empty() && !ix->quant_i8;
        for (uint32_t c : vl.cluster_counts) ok = ok && c != 0;
        SSB_TRY(vector_add_level_impl(ix, vl.level_id, vl.rows.data(), ix->dims, vl.ids.data(), (uint32_t)vl.ids.size(), ix->dims,
                                      ok ? vl.cluster_counts.data() : nullptr, ok ? (uint32_t)vl.cluster_counts.size() : 0));
        total += vl.ids.size();
    }
    if (n_vectors_out) *n_vectors_out = total;
    return SSB_OK;
    SSB_API_END
}

// shard.delete_hashset (index.rs:1594, filled by delete_document index.rs:5110): deleted docs are neither scored nor counted
// (add_result.rs:3435, vector.rs:1450-1451, union_count union.rs:975-1000).  Replaces the current set; n = 0 clears it.
int32_t ssb_set_deleted(ssb_index* ix, const uint64_t* doc_ids, uint64_t n) {
    SSB_API_BEGIN
    if (!ix || (n && !doc_ids)) { set_error("ssb_set_deleted: null argument"); return SSB_E_INVALID; }
    if (n >= (1ull << 31)) { set_error("ssb_set_deleted: too many doc ids"); return SSB_E_UNSUPPORTED; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    std::vector<uint32_t> docs(n);
    for (uint64_t i = 0; i < n; i++) {
        if (doc_ids[i] >> 32) { set_error("ssb_set_deleted: doc id %llu out of range", (unsigned long long)doc_ids[i]); return SSB_E_INVALID; }
        docs[i] = (uint32_t)doc_ids[i];
    }
    std::sort(docs.begin(), docs.end());
    docs.erase(std::unique(docs.begin(), docs.end()), docs.end());
    for (auto& c : ix->pool) cudaStreamSynchronize(c->own_st);
    ix->del.release();
    if (docs.empty()) return SSB_OK;
    std::vector<uint32_t> slot(65536, 0xFFFFFFFFu);
    uint32_t n_slots = 0;
    for (uint32_t d : docs) if (slot[d >> 16] == 0xFFFFFFFFu) slot[d >> 16] = n_slots++;
    std::vector<uint64_t> words((size_t)n_slots * 1024, 0ull);
    for (uint32_t d : docs) words[(size_t)slot[d >> 16] * 1024 + ((d & 0xFFFFu) >> 6)] |= 1ull << (d & 63u);
    SSB_CUDA_TRY(cudaMalloc(&ix->del.d_slot, 65536 * 4));
    SSB_CUDA_TRY(cudaMalloc(&ix->del.d_words, words.size() * 8));
    SSB_CUDA_TRY(cudaMalloc(&ix->del.d_docs, docs.size() * 4));
    SSB_CUDA_TRY(cudaMemcpy(ix->del.d_slot, slot.data(), 65536 * 4, cudaMemcpyHostToDevice));
    SSB_CUDA_TRY(cudaMemcpy(ix->del.d_words, words.data(), words.size() * 8, cudaMemcpyHostToDevice));
    SSB_CUDA_TRY(cudaMemcpy(ix->del.d_docs, docs.data(), docs.size() * 4, cudaMemcpyHostToDevice));
    ix->del.n = (uint32_t)docs.size();
    return SSB_OK;
    SSB_API_END
}

// facets_file_mmap (is_facet_filter, add_result.rs:340-478, reads `facets_size_sum * docid + facet.offset`): every value becomes an
// order-preserving 64-bit key, one column per facet, so the kernels test any FilterSparse range with two unsigned compares.
int32_t ssb_set_facets(ssb_index* ix, const void* rows, uint64_t first_doc_id, uint64_t n_docs, uint32_t row_bytes,
                       const ssb_facet_field* fields, uint32_t n_fields) {
    SSB_API_BEGIN
    if (!ix) { set_error("ssb_set_facets: null index"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    SSB_CUDA_TRY(cudaDeviceSynchronize());        // searches may run on a caller-owned stream (ssb_set_stream): nothing may still read the old columns
    ix->facets.release();
    if (n_docs == 0 || n_fields == 0) return SSB_OK;
    if (!rows || !fields) { set_error("ssb_set_facets: null argument"); return SSB_E_INVALID; }
    if (n_fields > SSB_MAX_FACETS) { set_error("ssb_set_facets: more than %u facets", SSB_MAX_FACETS); return SSB_E_UNSUPPORTED; }
    if (first_doc_id + n_docs > (1ull << 32)) { set_error("ssb_set_facets: doc ids must be < 2^32"); return SSB_E_INVALID; }
    for (uint32_t f = 0; f < n_fields; f++) {
        const uint32_t w = facet_type_bytes(fields[f].type);
        if (!w) { set_error("ssb_set_facets: field %u has unsupported type %u (Point facets are not built)", f, fields[f].type); return SSB_E_UNSUPPORTED; }
        if ((uint64_t)fields[f].offset + w > row_bytes) { set_error("ssb_set_facets: field %u does not fit a %u-byte row", f, row_bytes); return SSB_E_INVALID; }
    }
    std::vector<uint64_t> keys((size_t)n_fields * n_docs);
    const uint8_t* base = (const uint8_t*)rows;
    for (uint32_t f = 0; f < n_fields; f++) {
        uint64_t* col = keys.data() + (size_t)f * n_docs;
        const uint32_t type = fields[f].type, off = fields[f].offset;
        for (uint64_t d = 0; d < n_docs; d++) col[d] = facet_value_key(type, base + d * row_bytes + off);
    }
    SSB_CUDA_TRY(cudaMalloc(&ix->facets.d_keys, keys.size() * 8));
    SSB_CUDA_TRY(cudaMemcpy(ix->facets.d_keys, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice));
    ix->facets.n_rows = n_docs; ix->facets.first_doc = (uint32_t)first_doc_id; ix->facets.n_facets = n_fields;
    for (uint32_t f = 0; f < n_fields; f++) ix->facets.types[f] = (uint8_t)fields[f].type;
    return SSB_OK;
    SSB_API_END
}

// TurboQuant.seed_mask (vector_similarity.rs:1845-1859): the reference draws the +-1 mask once per index from ChaCha8Rng::seed_from_u64(1234)
// (index.rs:2215-2216) — a third-party generator this library does not restate; the host hands over the mask it holds.
int32_t ssb_vector_set_turboquant_mask(ssb_index* ix, const float* seed_mask, uint32_t dim) {
    SSB_API_BEGIN
    if (!ix || !seed_mask) { set_error("ssb_vector_set_turboquant_mask: null argument"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    if (!ix->turbo) { set_error("ssb_vector_set_turboquant_mask: the index was not created with SSB_QUANT_TURBO_I8"); return SSB_E_STATE; }
    if (ix->n_rows) { set_error("ssb_vector_set_turboquant_mask: call it before the first vector level"); return SSB_E_STATE; }
    if (dim != ix->tq_dim) { set_error("ssb_vector_set_turboquant_mask: dim %u, expected next_power_of_two(vector_dims) = %u", dim, ix->tq_dim); return SSB_E_INVALID; }
    std::vector<float> m(dim);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    SSB_CUDA_TRY(cudaMemcpy(m.data(), seed_mask, (size_t)dim * 4, cudaMemcpyDefault));
    for (float x : m) if (x != 1.0f && x != -1.0f) { set_error("ssb_vector_set_turboquant_mask: the mask must hold +1 / -1"); return SSB_E_INVALID; }
    if (!ix->tq_mask) SSB_CUDA_TRY(cudaMalloc(&ix->tq_mask, (size_t)dim * 4));
    SSB_CUDA_TRY(cudaMemcpy(ix->tq_mask, m.data(), (size_t)dim * 4, cudaMemcpyHostToDevice));
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_set_vector_kernel(ssb_index* ix, uint32_t kernel) {
    SSB_API_BEGIN
    if (!ix || kernel > SSB_VEC_KERNEL_TCGEN05_FILTER_N256_PAIR) { set_error("bad vector kernel"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    ix->cfg.vector_kernel = kernel;
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_vector_count(const ssb_index* ix, uint64_t* n) { if (!ix || !n) return SSB_E_INVALID; *n = ix->n_rows; return SSB_OK; }

int32_t ssb_search_vector_keys(ssb_index* ix, const float* queries, uint32_t nq, uint32_t k, uint64_t* keys_out_dev) {
    SSB_API_BEGIN
    if (!ix || (nq && (!queries || !keys_out_dev))) { set_error("ssb_search_vector_keys: null argument"); return SSB_E_INVALID; }
    std::shared_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    CtxLease l(ix); SSB_TRY(l.acquire());
    SSB_TRY(vec_keys(ix, *l.c, queries, false, nq, k, keys_out_dev));
    return shard_merge(ix, *l.c, keys_out_dev, nq);
    SSB_API_END
}

int32_t ssb_search_vector(ssb_index* ix, const float* queries, uint32_t nq, uint32_t k, ssb_hit* hits, uint32_t* n_hits) {
    SSB_API_BEGIN
    if (!ix || (nq && (!queries || !hits))) { set_error("ssb_search_vector: null argument"); return SSB_E_INVALID; }
    std::shared_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (nq == 0) return SSB_OK;
    if (k == 0 || k > SSB_K_LIMIT) { set_error("k must be in 1..%u", SSB_K_LIMIT); return SSB_E_UNSUPPORTED; }
    CtxLease l(ix); SSB_TRY(l.acquire());
    SSB_TRY(search_vector_host(ix, *l.c, queries, false, nq, k, hits, n_hits));
    finish_stats(ix, *l.c);
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_search_vector_ex(ssb_index* ix, const ssb_vec_query* vq, ssb_hit* hits, uint32_t* n_hits, ssb_hit_ext* ext, uint64_t* observed) {
    SSB_API_BEGIN
    if (!ix || !vq || (vq->n_queries && (!vq->queries || !hits))) { set_error("ssb_search_vector_ex: null argument"); return SSB_E_INVALID; }
    if (vq->query_format > SSB_QFMT_I8) { set_error("bad query_format"); return SSB_E_INVALID; }
    std::shared_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    const uint32_t nq = vq->n_queries, k = vq->k;
    if (nq == 0) return SSB_OK;
    if (k == 0 || k > SSB_K_LIMIT) { set_error("k must be in 1..%u", SSB_K_LIMIT); return SSB_E_UNSUPPORTED; }
    CtxLease l(ix); SSB_TRY(l.acquire());
    std::vector<uint32_t> nh(nq, 0);
    const bool euclid = ix->cfg.vector_similarity == SSB_SIM_EUCLIDEAN;
    if (vq->ann_mode > SSB_ANN_NPROBE_SIMILARITY_THRESHOLD) { set_error("bad ann_mode"); return SSB_E_INVALID; }
    IvfQuery ivf{vq->ann_mode, vq->n_probe, 0.f};
    {   // the cluster threshold goes through the same pre-map as the result threshold (TopK::new, vector.rs:388-399)
        volatile float c2 = vq->cluster_threshold * 2.0f; volatile float c21 = c2 - 1.0f;
        ivf.thr = euclid ? -vq->cluster_threshold : c21 / (1.0f / 16129.0f);
    }
    const bool use_ivf = vq->ann_mode != SSB_ANN_ALL;
    SSB_TRY(search_vector_host(ix, *l.c, vq->queries, vq->query_format == SSB_QFMT_I8, nq, k, hits, nh.data(), use_ivf ? &ivf : nullptr));
    finish_stats(ix, *l.c);
    // TopK::new (vector.rs:388-399): threshold pre-map (2t-1)*16129 for Dot/Cosine, -t for Euclidean; TopK::push (:421) rejects
    // score < threshold.  The hits are sorted by score, so dropping the tail is the same filter.
    volatile float t2 = vq->similarity_threshold * 2.0f; volatile float t21 = t2 - 1.0f;
    const float cut = euclid ? -vq->similarity_threshold : t21 / (1.0f / 16129.0f);
    for (uint32_t q = 0; q < nq; q++) {
        uint32_t n = nh[q];
        if (vq->has_threshold) {
            uint32_t m = 0;
            while (m < n && !(hits[(size_t)q * k + m].score < cut)) m++;
            for (uint32_t j = m; j < n; j++) hits[(size_t)q * k + j] = ssb_hit{0, 0.f, 0};
            n = m;
        }
        if (n_hits) n_hits[q] = n;
        if (observed) observed[q] = use_ivf ? l.c->h_obs[q] : ix->n_rows;   // AnnMode::All scores every record (observed_vector_count, vector.rs:420); else: the selected clusters' vectors
        if (ext) for (uint32_t j = 0; j < k; j++) {
            ssb_hit_ext& e = ext[(size_t)q * k + j];
            memset(&e, 0, sizeof(e));
            if (j >= n) continue;
            const ssb_hit& h = hits[(size_t)q * k + j];
            e.level_id = (uint32_t)(h.doc_id >> 16);           // vector.rs:1448 doc id = level << 16 | local
            volatile float sn = h.score * (1.0f / 16129.0f); volatile float s1 = sn + 1.0f;
            e.vector_score = euclid ? -h.score : s1 * 0.5f;     // vector.rs:1495-1499 (SIMILARITY_NORMALIZATION_64_I8 regardless of precision)
            e.cluster_score = euclid ? 0.f : 0.5f;              // no clustering (AnnMode::All, Clustering::None): cluster_score = 0 -> post-map 0.5 / -0
            e.source = SSB_SOURCE_VECTOR;
        }
    }
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_search_lexical_keys(ssb_index* ix, const ssb_lex_batch* q, uint32_t k, uint32_t result_type, uint64_t* keys_out_dev,
                                uint64_t* count_dev) {
    SSB_API_BEGIN
    if (!ix || !q) { set_error("ssb_search_lexical_keys: null argument"); return SSB_E_INVALID; }
    std::shared_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    CtxLease l(ix); SSB_TRY(l.acquire());
    l.c->ev_used = true; l.c->last_lex = true;
    SSB_TRY(ix->lex->search_keys(l.c->lex, l.c->st, q, k, result_type, keys_out_dev, count_dev, &l.c->stats.kernel_launches));
    SSB_TRY(shard_merge(ix, *l.c, keys_out_dev, q->n_queries));
    return shard_sum_counts(ix, *l.c, count_dev, q->n_queries);
    SSB_API_END
}

int32_t ssb_search_lexical(ssb_index* ix, const ssb_lex_batch* q, uint32_t k, uint32_t result_type, ssb_hit* hits, uint32_t* n_hits,
                           uint64_t* count_total) {
    SSB_API_BEGIN
    if (!ix || !q || (q->n_queries && k && result_type != SSB_RESULT_COUNT && !hits)) { set_error("ssb_search_lexical: null argument"); return SSB_E_INVALID; }
    std::shared_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    const uint32_t nq = q->n_queries;
    if (nq == 0) return SSB_OK;
    if (k > SSB_K_LIMIT) { set_error("k=%u exceeds SSB_K_LIMIT=%u", k, SSB_K_LIMIT); return SSB_E_UNSUPPORTED; }
    CtxLease l(ix); SSB_TRY(l.acquire());
    SearchCtx& c = *l.c;
    SSB_TRY(c.keys_a.reserve((size_t)nq * LIST, 0, c.st));
    SSB_TRY(c.counts.reserve(nq, 0, c.st));
    c.h_keys_a.resize((size_t)nq * LIST); c.h_counts.resize(nq);
    const bool want_hits = hits && k && result_type != SSB_RESULT_COUNT;
    const uint32_t k1 = k < SSB_K_MAX ? k : SSB_K_MAX;
    SSB_TRY(ix->lex->search_keys(c.lex, c.st, q, k1, result_type, c.keys_a.p, c.counts.p, &c.stats.kernel_launches));
    SSB_TRY(shard_merge(ix, c, c.keys_a.p, nq));
    SSB_TRY(shard_sum_counts(ix, c, c.counts.p, nq));
    SSB_CUDA_TRY(cudaMemcpyAsync(c.h_keys_a.data(), c.keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, c.st));
    SSB_CUDA_TRY(cudaMemcpyAsync(c.h_counts.data(), c.counts.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, c.st));
    SSB_CUDA_TRY(cudaStreamSynchronize(c.st));
    c.stats.d2h_bytes += (uint64_t)nq * (LIST * 8 + 8);
    c.ev_used = true;
    finish_stats(ix, c);                                        // the first page's scoring kernel is the one reported
    const LexStats ls = LexIndex::read_stats(c.lex, c.st);
    if (want_hits) {
        PageState ps(c, nq, k, hits, n_hits);
        bool more = ps.append(c.h_keys_a.data(), k1);
        // pages beyond the first 32 results: Topk search restricted to keys below the previous page's last key
        while (more) {
            const uint32_t kk = ps.next_page_k();
            SSB_TRY(ps.upload_ceilings());
            SSB_TRY(ix->lex->search_keys(c.lex, c.st, q, kk, SSB_RESULT_TOPK, c.keys_a.p, nullptr, &c.stats.kernel_launches, c.ceil.p));
            SSB_TRY(shard_merge(ix, c, c.keys_a.p, nq));
            SSB_CUDA_TRY(cudaMemcpyAsync(c.h_keys_a.data(), c.keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, c.st));
            SSB_CUDA_TRY(cudaStreamSynchronize(c.st));
            c.stats.d2h_bytes += (uint64_t)nq * LIST * 8;
            more = ps.append(c.h_keys_a.data(), kk);
        }
        ps.finish();
    } else if (n_hits) for (uint32_t i = 0; i < nq; i++) n_hits[i] = 0;
    if (count_total) for (uint32_t i = 0; i < nq; i++) count_total[i] = c.h_counts[i];
    c.stats.postings_visited = ls.postings_visited;
    // SURVEY.md §8(d) accounting with this layout's sizes: 4 B per streamed posting word, 16 B per probe (8 B bitmap word + 4 B
    // rank / word bound + 4 B component), 128 B per (query, level) record read, 8 B per bitmap word of the word-wise count paths;
    // the 1-byte coarse-table lookups of the stream filter are not counted
    c.stats.algorithmic_bytes = ls.postings_visited * 4 + ls.probes * 16 + ls.recs_processed * 128 + ls.dense_words * 8;
    c.stats.probes = ls.probes; c.stats.items_processed = ls.items_processed; c.stats.items_skipped = ls.items_skipped;
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_rrf_fuse(const ssb_hit* lex, uint32_t n_lex, const ssb_hit* vec, uint32_t n_vec, ssb_hit* out, uint32_t* n_out) {
    SSB_API_BEGIN
    // search.rs:1962-2035: k = 0.6, rank from 0 over each list sorted by score desc; then :2097-2106 sort desc.
    if ((n_lex && !lex) || (n_vec && !vec) || !out) { set_error("ssb_rrf_fuse: null argument"); return SSB_E_INVALID; }
    const float kf = 0.6f;
    uint32_t n = 0;
    for (uint32_t i = 0; i < n_lex; i++) {
        volatile float denom = kf + (float)i; float s = 1.0f / denom;
        uint32_t j = 0; for (; j < n; j++) if (out[j].doc_id == lex[i].doc_id) break;
        if (j == n) { out[n].doc_id = lex[i].doc_id; out[n].score = s; out[n].pad = 0; n++; } else out[j].score = s;
    }
    for (uint32_t i = 0; i < n_vec; i++) {
        volatile float denom = kf + (float)i; float s = 1.0f / denom;
        uint32_t j = 0; for (; j < n; j++) if (out[j].doc_id == vec[i].doc_id) break;
        if (j == n) { out[n].doc_id = vec[i].doc_id; out[n].score = s; out[n].pad = 0; n++; }
        else { volatile float sum = out[j].score + s; out[j].score = sum; }
    }
    std::stable_sort(out, out + n, hit_better);
    if (n_out) *n_out = n;
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_search_hybrid(ssb_index* ix, const ssb_lex_batch* q, const float* queries, uint32_t k, ssb_hit* hits, uint32_t* n_hits) {
    SSB_API_BEGIN
    if (!ix || !q || !queries || !hits) { set_error("ssb_search_hybrid: null argument"); return SSB_E_INVALID; }
    if (k == 0 || k > SSB_K_MAX) { set_error("k must be in 1..%u", SSB_K_MAX); return SSB_E_UNSUPPORTED; }
    std::shared_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    const uint32_t nq = q->n_queries;
    if (nq == 0) return SSB_OK;
    // the two per-shard searches are independent until the fusion: the lexical one runs on this context's stream, the vector one
    // on a second context's stream, so lex_score and the scan share the GPU instead of running back to back
    CtxLease l(ix); SSB_TRY(l.acquire());
    SearchCtx& c = *l.c;
    CtxLease l2(ix);
    SearchCtx* cv = &c;
    if (!ix->ext_stream_set) { SSB_TRY(l2.acquire(false)); if (l2.c) cv = l2.c; }   // never wait for a second context (no hold-and-wait)
    SSB_TRY(c.keys_a.reserve((size_t)nq * LIST, 0, c.st));
    SSB_TRY(cv->keys_b.reserve((size_t)nq * LIST, 0, cv->st));
    c.h_keys_a.resize((size_t)nq * LIST); cv->h_keys_b.resize((size_t)nq * LIST);
    SSB_TRY(ix->lex->search_keys(c.lex, c.st, q, k, SSB_RESULT_TOPK, c.keys_a.p, nullptr, &c.stats.kernel_launches));
    // sharded: both lists are merged over the shards FIRST and fused afterwards — RRF ranks are positions in the merged lists
    // (search.rs:1962-2035 runs after the shard results were concatenated and sorted)
    SSB_TRY(shard_merge(ix, c, c.keys_a.p, nq));
    SSB_CUDA_TRY(cudaMemcpyAsync(c.h_keys_a.data(), c.keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, c.st));
    // multi-chunk documents: fetch the full 32-list so that k distinct docs survive the per-doc de-duplication
    SSB_TRY(vec_keys(ix, *cv, queries, false, nq, ix->dup_docs ? SSB_K_MAX : k, cv->keys_b.p));
    SSB_TRY(shard_merge(ix, *cv, cv->keys_b.p, nq));
    SSB_CUDA_TRY(cudaMemcpyAsync(cv->h_keys_b.data(), cv->keys_b.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, cv->st));
    SSB_CUDA_TRY(cudaStreamSynchronize(c.st));
    if (cv != &c) SSB_CUDA_TRY(cudaStreamSynchronize(cv->st));
    c.stats.kernel_launches += cv != &c ? cv->stats.kernel_launches : 0;
    c.stats.h2d_bytes += cv != &c ? cv->stats.h2d_bytes : 0;
    c.stats.d2h_bytes += (uint64_t)nq * LIST * 16;
    std::vector<ssb_hit> a(k), b(k), f(2 * (size_t)k);
    for (uint32_t i = 0; i < nq; i++) {
        uint32_t nf = 0;
        const uint32_t na = decode_list(c.h_keys_a.data() + (size_t)i * LIST, k, a.data(), false);
        const uint32_t nb = decode_list(cv->h_keys_b.data() + (size_t)i * LIST, k, b.data(), ix->dup_docs);
        ssb_rrf_fuse(a.data(), na, b.data(), nb, f.data(), &nf);
        uint32_t n = nf < k ? nf : k;     // search.rs:2117-2119 truncate(length)
        for (uint32_t j = 0; j < k; j++) hits[(size_t)i * k + j] = j < n ? f[j] : ssb_hit{0, 0.f, 0};
        if (n_hits) n_hits[i] = n;
    }
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_merge_keys(ssb_index* ix, const uint64_t* keys_dev, uint32_t n_lists, uint32_t nq, uint32_t k, ssb_hit* hits, uint32_t* n_hits) {
    SSB_API_BEGIN
    if (!ix || !keys_dev || !hits) { set_error("ssb_merge_keys: null argument"); return SSB_E_INVALID; }
    if (k == 0 || k > SSB_K_MAX) { set_error("k must be in 1..%u", SSB_K_MAX); return SSB_E_UNSUPPORTED; }
    std::shared_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (nq == 0) return SSB_OK;
    CtxLease l(ix); SSB_TRY(l.acquire());
    SearchCtx& c = *l.c;
    SSB_TRY(c.keys_b.reserve((size_t)nq * LIST, 0, c.st));
    SSB_TRY(vec::launch_merge_lists(keys_dev, n_lists, nq, c.keys_b.p, c.st));
    c.stats.kernel_launches += 1;
    c.h_keys_b.resize((size_t)nq * LIST);
    SSB_CUDA_TRY(cudaMemcpyAsync(c.h_keys_b.data(), c.keys_b.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, c.st));
    SSB_CUDA_TRY(cudaStreamSynchronize(c.st));
    decode_keys(c.h_keys_b.data(), nq, k, hits, n_hits, ix->dup_docs);
    return SSB_OK;
    SSB_API_END
}

// ---- sharded index: one process per GPU, NCCL communicator owned by (or lent to) the handle -------------------------------
int32_t ssb_comm_unique_id(uint8_t* id128) {
    SSB_API_BEGIN
    if (!id128) { set_error("ssb_comm_unique_id: null argument"); return SSB_E_INVALID; }
    return comm_unique_id(id128);
    SSB_API_END
}

int32_t ssb_comm_init(ssb_index* ix, const uint8_t* id128, uint32_t rank, uint32_t world) {
    SSB_API_BEGIN
    if (!ix || !id128) { set_error("ssb_comm_init: null argument"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (ix->comm.comm) { set_error("ssb_comm_init: the index already has a communicator"); return SSB_E_STATE; }
    SSB_TRY(comm_init(ix->comm, id128, rank, world));
    std::lock_guard<std::mutex> g2(ix->pool_mu);
    if (ix->pool.size() > 1) { ix->pool.resize(1); ix->free_ctx.clear(); ix->free_ctx.push_back(ix->pool[0].get()); ix->last_ctx = nullptr; }
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_comm_attach(ssb_index* ix, void* nccl_comm, uint32_t rank, uint32_t world) {
    SSB_API_BEGIN
    if (!ix || !nccl_comm || world == 0 || rank >= world) { set_error("ssb_comm_attach: bad argument"); return SSB_E_INVALID; }
    std::unique_lock<std::shared_mutex> g(ix->rw);
    if (ix->comm.comm) { set_error("ssb_comm_attach: the index already has a communicator"); return SSB_E_STATE; }
    ix->comm.comm = nccl_comm; ix->comm.rank = rank; ix->comm.world = world; ix->comm.owned = false;
    std::lock_guard<std::mutex> g2(ix->pool_mu);
    if (ix->pool.size() > 1) { ix->pool.resize(1); ix->free_ctx.clear(); ix->free_ctx.push_back(ix->pool[0].get()); ix->last_ctx = nullptr; }
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_comm_destroy(ssb_index* ix) {
    SSB_API_BEGIN
    if (!ix) return SSB_E_INVALID;
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    comm_destroy(ix->comm);
    return SSB_OK;
    SSB_API_END
}

// Global document frequencies of a sharded index (idf must use the df of the WHOLE index, search.rs:3225-3230): every rank
// contributes its dictionary (sorted keys + local df), the sum per key is installed on every rank.  Collective.
int32_t ssb_lexical_sync_df(ssb_index* ix) {
    SSB_API_BEGIN
    if (!ix) return SSB_E_INVALID;
    std::unique_lock<std::shared_mutex> g(ix->rw);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (!ix->comm.active()) return SSB_OK;
    if (!ix->lex->committed()) { set_error("ssb_lexical_sync_df before ssb_lexical_commit"); return SSB_E_STATE; }
    cudaStream_t st = ix->load_st;
    const std::vector<uint64_t>& keys = ix->lex->host_keys();
    const std::vector<uint32_t>& dfs = ix->lex->host_local_df();
    const uint32_t world = ix->comm.world;
    DevTmp<uint64_t> d_n; SSB_CUDA_TRY(d_n.alloc(1));
    uint64_t n_max = keys.size();
    SSB_CUDA_TRY(cudaMemcpyAsync(d_n.p, &n_max, 8, cudaMemcpyHostToDevice, st));
    SSB_TRY(comm_all_reduce_max_u64(ix->comm, d_n.p, 1, st));
    SSB_CUDA_TRY(cudaMemcpyAsync(&n_max, d_n.p, 8, cudaMemcpyDeviceToHost, st));
    SSB_CUDA_TRY(cudaStreamSynchronize(st));
    if (n_max == 0) return SSB_OK;
    // send [n_max] keys (padded with ~0, which sorts last and matches nothing) and [n_max] dfs
    std::vector<uint64_t> sk(n_max, ~0ull), sd(n_max, 0);
    for (size_t i = 0; i < keys.size(); i++) { sk[i] = keys[i]; sd[i] = dfs[i]; }
    DevTmp<uint64_t> d_sk, d_sd, d_rk, d_rd;
    SSB_CUDA_TRY(d_sk.alloc(n_max)); SSB_CUDA_TRY(d_sd.alloc(n_max)); SSB_CUDA_TRY(d_rk.alloc(n_max * world)); SSB_CUDA_TRY(d_rd.alloc(n_max * world));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_sk.p, sk.data(), n_max * 8, cudaMemcpyHostToDevice, st));
    SSB_CUDA_TRY(cudaMemcpyAsync(d_sd.p, sd.data(), n_max * 8, cudaMemcpyHostToDevice, st));
    SSB_TRY(comm_all_gather_u64(ix->comm, d_sk.p, d_rk.p, n_max, st));
    SSB_TRY(comm_all_gather_u64(ix->comm, d_sd.p, d_rd.p, n_max, st));
    std::vector<uint64_t> rk(n_max * world), rd(n_max * world);
    SSB_CUDA_TRY(cudaMemcpyAsync(rk.data(), d_rk.p, n_max * world * 8, cudaMemcpyDeviceToHost, st));
    SSB_CUDA_TRY(cudaMemcpyAsync(rd.data(), d_rd.p, n_max * world * 8, cudaMemcpyDeviceToHost, st));
    SSB_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<uint32_t> total(keys.size(), 0);
    for (uint32_t r = 0; r < world; r++) {          // merge-join of two sorted key arrays
        const uint64_t* k2 = rk.data() + (size_t)r * n_max; const uint64_t* d2 = rd.data() + (size_t)r * n_max;
        size_t j = 0;
        for (size_t i = 0; i < keys.size(); i++) {
            while (j < n_max && k2[j] < keys[i]) j++;
            if (j < n_max && k2[j] == keys[i]) total[i] += (uint32_t)d2[j];
        }
    }
    return ix->lex->set_global_df(keys.data(), total.data(), keys.size());
    SSB_API_END
}

int32_t ssb_sync(ssb_index* ix) {
    SSB_API_BEGIN
    if (!ix) return SSB_E_INVALID;
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    std::vector<cudaStream_t> sts;
    {
        std::lock_guard<std::mutex> g(ix->pool_mu);
        for (auto& c : ix->pool) sts.push_back(ix->ext_stream_set ? ix->ext_stream : c->own_st);
    }
    for (cudaStream_t s : sts) SSB_CUDA_TRY(cudaStreamSynchronize(s));
    return SSB_OK;
    SSB_API_END
}

void* ssb_stream(ssb_index* ix) {
    if (!ix) return nullptr;
    if (ix->ext_stream_set) return (void*)ix->ext_stream;
    std::lock_guard<std::mutex> g(ix->pool_mu);
    return ix->pool.empty() ? (void*)ix->load_st : (void*)ix->pool[0]->own_st;
}

int32_t ssb_set_stream(ssb_index* ix, void* stream) {
    SSB_API_BEGIN
    if (!ix) return SSB_E_INVALID;
    std::unique_lock<std::shared_mutex> g(ix->rw);       // no search in flight
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    {
        std::lock_guard<std::mutex> g2(ix->pool_mu);
        for (auto& c : ix->pool) SSB_CUDA_TRY(cudaStreamSynchronize(ix->ext_stream_set ? ix->ext_stream : c->own_st));
        if (stream == SSB_OWN_STREAM) { ix->ext_stream_set = false; ix->ext_stream = nullptr; }
        else { ix->ext_stream_set = true; ix->ext_stream = (cudaStream_t)stream; }
        // with a caller-owned stream every search runs on that one stream: keep a single context
        if (ix->ext_stream_set && ix->pool.size() > 1) {
            ix->pool.resize(1); ix->free_ctx.clear(); ix->free_ctx.push_back(ix->pool[0].get()); ix->last_ctx = nullptr;
        }
    }
    return SSB_OK;
    SSB_API_END
}

int32_t ssb_last_stats(const ssb_index* cix, ssb_stats* out) {
    SSB_API_BEGIN
    if (!cix || !out) return SSB_E_INVALID;
    ssb_index* ix = const_cast<ssb_index*>(cix);
    SearchCtx* c = nullptr;
    { std::lock_guard<std::mutex> g(ix->stats_mu); *out = ix->last_stats; c = ix->last_ctx; }
    if (c && c->ev_used && out->dominant_kernel_ns == 0) {   // asynchronous *_keys call: wait for its kernel's events now
        cudaSetDevice(ix->cfg.device);
        float ms = 0.f;
        if (cudaEventSynchronize(c->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess)
            out->dominant_kernel_ns = (uint64_t)((double)ms * 1e6);
        else cudaGetLastError();
        if (c->last_lex && out->postings_visited == 0 && c->lex.stats) {
            const LexStats ls = LexIndex::read_stats(c->lex, ix->ext_stream_set ? ix->ext_stream : c->own_st);
            out->postings_visited = ls.postings_visited; out->probes = ls.probes; out->items_processed = ls.items_processed; out->items_skipped = ls.items_skipped;
            if (ls.postings_visited) out->algorithmic_bytes = ls.postings_visited * 4 + ls.probes * 16 + ls.recs_processed * 128 + ls.dense_words * 8;
        }
    }
    return SSB_OK;
    SSB_API_END
}

}  // extern "C"
