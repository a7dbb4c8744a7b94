// api.cu — the extern "C" ABI of libseekstorm_b200.so (include/seekstorm_b200.h): handle, vector index
// storage, search entry points, hybrid RRF, key decoding.  No torch types; CUDA runtime only.
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "bm25.h"
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

struct ssb_index {
    ssb_config cfg;
    int n_sms = 0;
    cudaStream_t st = nullptr;       // stream in use
    cudaStream_t own_st = nullptr;   // the index's own stream
    std::mutex mu;
    LexIndex* lex = nullptr;
    // vector index
    uint32_t dims = 0, dpad = 0, dpad8 = 0;
    bool quant_i8 = false;            // Cosine + ScalarQuantizationI8: int8 corpus, exact int32 dot products
    DevBuf<float> rows;
    DevBuf<int8_t> rows_i8, q_i8;
    DevBuf<uint32_t> doc_ids;
    uint64_t n_rows = 0;
    // query workspace
    DevBuf<float> qpad; DevBuf<float> qstage; DevBuf<float> qhi, qlo; DevBuf<uint64_t> ceil;
    std::vector<uint64_t> h_ceil; DevBuf<uint64_t> scratch; DevBuf<uint64_t> keys_a, keys_b, counts;
    std::vector<uint64_t> h_keys_a, h_keys_b, h_counts;
    ssb_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool ev_used = false;
};

namespace {

int32_t vec_keys(ssb_index* ix, const float* queries, uint32_t nq, uint32_t k, uint64_t* keys_out_dev /*[nq_pad][32] min*/,
                 const uint64_t* ceil_dev = nullptr /*[>= nq_pad] paging ceilings*/) {
    if (ix->dims == 0) { set_error("no vector index configured (vector_dims = 0)"); return SSB_E_STATE; }
    if (k == 0 || k > SSB_K_MAX) { set_error("k must be in 1..%u", SSB_K_MAX); return SSB_E_UNSUPPORTED; }
    if (nq == 0) return SSB_OK;
    // AUTO (measured, 1M x 768): one FP32 pass of 16 queries takes 0.53 ms, one tensor-core pass of up to 128 queries
    // 0.81 ms -> FP32 scan for <= 16 queries, tensor-core scan above.  Euclidean always takes the FP32 scan.
    uint32_t kern = ix->cfg.vector_kernel;
    if (ix->quant_i8) kern = SSB_VEC_KERNEL_TCGEN05;   // one kernel for the int8 corpus: tcgen05 kind::i8, 128-query tile
    if (kern == SSB_VEC_KERNEL_AUTO) kern = nq > 16 ? SSB_VEC_KERNEL_TCGEN05_BF16 : SSB_VEC_KERNEL_FFMA;
    const bool use_tc = kern >= SSB_VEC_KERNEL_TCGEN05 && ix->cfg.vector_similarity != SSB_SIM_EUCLIDEAN;
    const bool tc_bf16 = kern == SSB_VEC_KERNEL_TCGEN05_BF16 || kern == SSB_VEC_KERNEL_TCGEN05_BF16_N64;
    const uint32_t qt = !use_tc ? vec::VEC_QT : ((kern == SSB_VEC_KERNEL_TCGEN05_N64 || kern == SSB_VEC_KERNEL_TCGEN05_BF16_N64) ? 64u : 128u);
    const uint32_t nq_pad = (nq + qt - 1) / qt * qt;
    if (!ix->quant_i8) SSB_TRY(ix->qpad.reserve((size_t)nq_pad * ix->dpad, 0, ix->st));
    const float* qsrc = queries;
    if (!dev_ptr(queries)) {
        SSB_TRY(ix->qstage.reserve((size_t)nq * ix->dims, 0, ix->st));
        SSB_CUDA_TRY(cudaMemcpyAsync(ix->qstage.p, queries, (size_t)nq * ix->dims * 4, cudaMemcpyHostToDevice, ix->st));
        ix->stats.h2d_bytes += (uint64_t)nq * ix->dims * 4;
        qsrc = ix->qstage.p;
    }
    if (ix->quant_i8) {
        // the query is normalised and quantised exactly like the corpus (search.rs:1464-1475, vector_similarity.rs:1226-1232)
        SSB_TRY(ix->q_i8.reserve((size_t)nq_pad * ix->dpad8, 0, ix->st));
        SSB_TRY(vec::launch_quantize_rows_i8(qsrc, ix->dims, nq, nq_pad, ix->dims, ix->q_i8.p, ix->dpad8, ix->st));
    } else
    SSB_TRY(vec::launch_prep_queries(qsrc, nq, ix->dims, ix->dims, ix->qpad.p, nq_pad, ix->dpad,
                                     ix->cfg.vector_similarity == SSB_SIM_COSINE, ix->st));
    ix->stats.kernel_launches += 1;
    if (ix->n_rows == 0) { SSB_CUDA_TRY(cudaMemsetAsync(keys_out_dev, 0, (size_t)nq * LIST * 8, ix->st)); return SSB_OK; }
    size_t sb = use_tc ? vec::scan_tc_scratch_bytes(ix->n_sms, nq_pad) : vec::scan_scratch_bytes(ix->n_sms, nq_pad);
    SSB_TRY(ix->scratch.reserve(sb / 8 + (size_t)nq_pad * LIST + (nq_pad + 1) / 2, 0, ix->st));
    vec::ScanArgs a{};
    a.rows = ix->rows.p; a.doc_ids = ix->doc_ids.p; a.n_rows = ix->n_rows; a.dpad = ix->dpad; a.queries_padded = ix->qpad.p;
    a.nq_pad = nq_pad; a.nq_valid = nq; a.k = k; a.similarity = ix->cfg.vector_similarity; a.n_sms = ix->n_sms;
    a.scratch = ix->scratch.p; a.scratch_bytes = sb;
    uint64_t* merged = ix->scratch.p + sb / 8;   // [nq_pad][32]
    a.keys_out = merged; a.ev0 = ix->ev0; a.ev1 = ix->ev1; ix->ev_used = true;
    a.thr_buf = reinterpret_cast<uint32_t*>(merged + (size_t)nq_pad * LIST);
    a.ceil_keys = ceil_dev;
    a.launches = &ix->stats.kernel_launches;
    if (ix->quant_i8) {
        a.rows_i8 = ix->rows_i8.p; a.queries_i8 = ix->q_i8.p; a.dpad8 = ix->dpad8;
        SSB_TRY(vec::launch_scan_tc(a, 128, 2, ix->st));
    } else if (use_tc) {
        SSB_TRY(ix->qhi.reserve((size_t)nq_pad * ix->dpad, 0, ix->st));
        SSB_TRY(ix->qlo.reserve((size_t)nq_pad * ix->dpad, 0, ix->st));
        a.q_hi = ix->qhi.p; a.q_lo = ix->qlo.p;
        SSB_TRY(vec::launch_scan_tc(a, qt, tc_bf16 ? 1 : 0, ix->st));
    } else {
        SSB_TRY(vec::launch_scan_ffma(a, ix->st));
    }
    SSB_CUDA_TRY(cudaMemcpyAsync(keys_out_dev, merged, (size_t)nq * LIST * 8, cudaMemcpyDeviceToDevice, ix->st));
    ix->stats.algorithmic_bytes += (uint64_t)(nq_pad / qt) * ix->n_rows * ix->dims * (ix->quant_i8 ? 1 : 4);
    return SSB_OK;
}

void decode_keys(const uint64_t* keys, uint32_t nq, uint32_t k, ssb_hit* hits, uint32_t* n_hits) {
    for (uint32_t q = 0; q < nq; q++) {
        uint32_t n = 0;
        for (uint32_t j = 0; j < k && j < LIST; j++) {
            uint64_t key = keys[(size_t)q * LIST + j];
            if (!key) break;
            hits[(size_t)q * k + n].doc_id = key_doc(key);
            hits[(size_t)q * k + n].score = key_score(key);
            hits[(size_t)q * k + n].pad = 0;
            n++;
        }
        for (uint32_t j = n; j < k; j++) { hits[(size_t)q * k + j].doc_id = 0; hits[(size_t)q * k + j].score = 0.f; hits[(size_t)q * k + j].pad = 0; }
        if (n_hits) n_hits[q] = n;
    }
}

// Paging state of the host-facing search calls (k > SSB_K_MAX).
struct PageState {
    ssb_index* ix; uint32_t nq, k; ssb_hit* hits; uint32_t* n_hits; std::vector<uint32_t> cnt;
    PageState(ssb_index* ix_, uint32_t nq_, uint32_t k_, ssb_hit* h, uint32_t* n) : ix(ix_), nq(nq_), k(k_), hits(h), n_hits(n), cnt(nq_, 0) {
        ix->h_ceil.assign((size_t)nq_ + 256, 0);
    }
    // append up to kk keys per query from a [nq][32] page; returns true if any query may have more results
    bool append(const uint64_t* keys, uint32_t done, uint32_t kk) {
        bool more = false;
        for (uint32_t q = 0; q < nq; q++) {
            if (cnt[q] < done) { ix->h_ceil[q] = 0; continue; }         // exhausted on an earlier page
            uint32_t n = 0; uint64_t last = 0;
            for (uint32_t j = 0; j < kk; j++) {
                const uint64_t key = keys[(size_t)q * LIST + j];
                if (!key) break;
                ssb_hit& h = hits[(size_t)q * k + cnt[q] + n];
                h.doc_id = key_doc(key); h.score = key_score(key); h.pad = 0;
                last = key; n++;
            }
            cnt[q] += n;
            ix->h_ceil[q] = n == kk ? last : 0;                           // 0 = nothing left below
            more = more || n == kk;
        }
        return more;
    }
    int32_t upload_ceilings() {
        SSB_TRY(ix->ceil.reserve((size_t)nq + 256, 0, ix->st));
        SSB_CUDA_TRY(cudaMemcpyAsync(ix->ceil.p, ix->h_ceil.data(), ((size_t)nq + 256) * 8, cudaMemcpyHostToDevice, ix->st));
        ix->stats.h2d_bytes += (uint64_t)nq * 8;
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

}  // namespace

extern "C" {

uint32_t ssb_abi_version(void) { return SSB_ABI_VERSION; }
const char* ssb_last_error(void) { return g_err; }

int32_t ssb_create(const ssb_config* cfg, ssb_index** out) {
    if (!cfg || !out) { set_error("ssb_create: null argument"); return SSB_E_INVALID; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device visible: libseekstorm_b200 has no CPU fallback"); return SSB_E_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("device %d out of range (%d visible)", cfg->device, ndev); return SSB_E_INVALID; }
    if (cfg->vector_similarity > SSB_SIM_EUCLIDEAN) { set_error("bad vector_similarity"); return SSB_E_INVALID; }
    if (cfg->vector_quantization > SSB_QUANT_SCALAR_I8) { set_error("bad vector_quantization"); return SSB_E_INVALID; }
    if (cfg->vector_quantization == SSB_QUANT_SCALAR_I8 && cfg->vector_similarity != SSB_SIM_COSINE) {
        // Dot / Euclidean + SQ carry per-vector scale / zero point (vector.rs:597-660): not built
        set_error("ScalarQuantizationI8 is built for Cosine similarity only"); return SSB_E_UNSUPPORTED;
    }
    SSB_CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    SSB_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor); return SSB_E_UNSUPPORTED; }
    ssb_index* ix = new (std::nothrow) ssb_index();
    if (!ix) return SSB_E_NOMEM;
    ix->cfg = *cfg;
    if (ix->cfg.max_batch == 0) ix->cfg.max_batch = 4096;
    ix->n_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ix->st, cudaStreamNonBlocking) != cudaSuccess) { delete ix; set_error("stream create failed"); return SSB_E_CUDA; }
    ix->own_st = ix->st;
    cudaEventCreate(&ix->ev0); cudaEventCreate(&ix->ev1);
    ix->lex = new LexIndex(ix->st, ix->n_sms, ix->cfg.max_batch);
    ix->lex->set_events(ix->ev0, ix->ev1);
    ix->dims = cfg->vector_dims;
    ix->dpad = (cfg->vector_dims + 31) / 32 * 32;
    ix->dpad8 = (cfg->vector_dims + 127) / 128 * 128;
    ix->quant_i8 = cfg->vector_quantization == SSB_QUANT_SCALAR_I8;
    *out = ix;
    return SSB_OK;
}

int32_t ssb_destroy(ssb_index* ix) {
    if (!ix) return SSB_OK;
    cudaSetDevice(ix->cfg.device);
    cudaStreamSynchronize(ix->st);
    delete ix->lex;
    ix->rows.release(); ix->rows_i8.release(); ix->q_i8.release(); ix->doc_ids.release(); ix->qpad.release(); ix->qstage.release(); ix->qhi.release(); ix->qlo.release(); ix->ceil.release(); ix->scratch.release();
    ix->keys_a.release(); ix->keys_b.release(); ix->counts.release();
    cudaEventDestroy(ix->ev0); cudaEventDestroy(ix->ev1);
    cudaStreamDestroy(ix->own_st);
    delete ix;
    return SSB_OK;
}

int32_t ssb_lexical_add_level(ssb_index* ix, const ssb_level_desc* level) {
    if (!ix) { set_error("null index"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    return ix->lex->add_level(level);
}

int32_t ssb_lexical_commit(ssb_index* ix, uint64_t n_docs, uint64_t len_sum) {
    if (!ix) { set_error("null index"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    return ix->lex->commit(n_docs, len_sum);
}

int32_t ssb_lexical_dict_size(const ssb_index* ix, uint64_t* n) { if (!ix || !n) return SSB_E_INVALID; return ix->lex->dict_size(n); }
int32_t ssb_lexical_dict_export(const ssb_index* ix, uint64_t* keys, uint32_t* dfs, uint64_t cap) { if (!ix) return SSB_E_INVALID; return ix->lex->dict_export(keys, dfs, cap); }
int32_t ssb_lexical_set_global_df(ssb_index* ix, const uint64_t* keys, const uint32_t* dfs, uint64_t n) {
    if (!ix || (n && (!keys || !dfs))) return SSB_E_INVALID;
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    return ix->lex->set_global_df(keys, dfs, n);
}

int32_t ssb_vector_add_level(ssb_index* ix, uint32_t level_id, const float* rows, uint64_t row_stride, const uint16_t* local_ids,
                             uint32_t n, uint32_t dims) {
    if (!ix || (n && !rows)) { set_error("ssb_vector_add_level: null argument"); return SSB_E_INVALID; }
    if (ix->dims == 0 || dims != ix->dims) { set_error("dims %u != configured vector_dims %u", dims, ix->dims); return SSB_E_INVALID; }
    if (n > 65536) { set_error("a level holds at most 65536 vectors"); return SSB_E_INVALID; }
    if (row_stride == 0) row_stride = dims;
    if (row_stride < dims) { set_error("row stride < dims"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (n == 0) return SSB_OK;
    SSB_TRY(ix->doc_ids.reserve(ix->n_rows + n, ix->n_rows, ix->st));
    if (ix->quant_i8) {
        // index-time normalise + quantise (vector.rs:585-640); the f32 rows are only staged
        SSB_TRY(ix->rows_i8.reserve((ix->n_rows + n) * ix->dpad8, ix->n_rows * ix->dpad8, ix->st));
        SSB_TRY(ix->qstage.reserve((size_t)n * dims, 0, ix->st));
        SSB_CUDA_TRY(cudaMemcpy2DAsync(ix->qstage.p, (size_t)dims * 4, rows, row_stride * 4, (size_t)dims * 4, n, cudaMemcpyDefault, ix->st));
        SSB_TRY(vec::launch_quantize_rows_i8(ix->qstage.p, dims, n, n, dims, ix->rows_i8.p + ix->n_rows * ix->dpad8, ix->dpad8, ix->st));
    } else {
        SSB_TRY(ix->rows.reserve((ix->n_rows + n) * ix->dpad, ix->n_rows * ix->dpad, ix->st));
        float* dst = ix->rows.p + ix->n_rows * ix->dpad;
        SSB_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ix->dpad * 4, rows, row_stride * 4, (size_t)dims * 4, n, cudaMemcpyDefault, ix->st));
        SSB_TRY(vec::launch_normalize_rows(dst, n, dims, ix->dpad, ix->cfg.vector_similarity == SSB_SIM_COSINE, ix->st));
    }
    const uint16_t* lid = local_ids; uint16_t* tmp = nullptr;
    if (local_ids && !dev_ptr(local_ids)) {
        SSB_CUDA_TRY(cudaMalloc(&tmp, (size_t)n * 2));
        SSB_CUDA_TRY(cudaMemcpyAsync(tmp, local_ids, (size_t)n * 2, cudaMemcpyHostToDevice, ix->st));
        lid = tmp;
    }
    SSB_TRY(vec::launch_fill_doc_ids(ix->doc_ids.p + ix->n_rows, lid, level_id, n, ix->st));
    SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
    if (tmp) cudaFree(tmp);
    ix->n_rows += n;
    return SSB_OK;
}

int32_t ssb_set_vector_kernel(ssb_index* ix, uint32_t kernel) {
    if (!ix || kernel > SSB_VEC_KERNEL_TCGEN05_BF16_N64) { set_error("bad vector kernel"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    ix->cfg.vector_kernel = kernel;
    return SSB_OK;
}

int32_t ssb_vector_count(const ssb_index* ix, uint64_t* n) { if (!ix || !n) return SSB_E_INVALID; *n = ix->n_rows; return SSB_OK; }

int32_t ssb_search_vector_keys(ssb_index* ix, const float* queries, uint32_t nq, uint32_t k, uint64_t* keys_out_dev) {
    if (!ix || (nq && (!queries || !keys_out_dev))) { set_error("ssb_search_vector_keys: null argument"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    ix->stats = ssb_stats{};
    return vec_keys(ix, queries, nq, k, keys_out_dev);
}

int32_t ssb_search_vector(ssb_index* ix, const float* queries, uint32_t nq, uint32_t k, ssb_hit* hits, uint32_t* n_hits) {
    if (!ix || (nq && (!queries || !hits))) { set_error("ssb_search_vector: null argument"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    ix->stats = ssb_stats{};
    if (nq == 0) return SSB_OK;
    if (k == 0 || k > SSB_K_LIMIT) { set_error("k must be in 1..%u", SSB_K_LIMIT); return SSB_E_UNSUPPORTED; }
    SSB_TRY(ix->keys_a.reserve((size_t)nq * LIST, 0, ix->st));
    ix->h_keys_a.resize((size_t)nq * LIST);
    PageState ps(ix, nq, k, hits, n_hits);
    // the fused kernels keep 32 results per query; longer result lists are produced page by page, each page restricted to
    // keys strictly below the last key of the previous one (keys are a total order on (score desc, doc id asc))
    for (uint32_t done = 0; done < k; done += SSB_K_MAX) {
        const uint32_t kk = k - done < SSB_K_MAX ? k - done : SSB_K_MAX;
        SSB_TRY(vec_keys(ix, queries, nq, kk, ix->keys_a.p, done ? ix->ceil.p : nullptr));
        SSB_CUDA_TRY(cudaMemcpyAsync(ix->h_keys_a.data(), ix->keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, ix->st));
        SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
        ix->stats.d2h_bytes += (uint64_t)nq * LIST * 8;
        if (!ps.append(ix->h_keys_a.data(), done, kk) || done + kk >= k) break;
        SSB_TRY(ps.upload_ceilings());
    }
    ps.finish();
    return SSB_OK;
}

int32_t ssb_search_lexical_keys(ssb_index* ix, const ssb_lex_batch* q, uint32_t k, uint32_t result_type, uint64_t* keys_out_dev,
                                uint64_t* count_dev) {
    if (!ix || !q) { set_error("ssb_search_lexical_keys: null argument"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    ix->stats = ssb_stats{};
    ix->ev_used = true;
    return ix->lex->search_keys(q, k, result_type, keys_out_dev, count_dev, &ix->stats.kernel_launches);
}

int32_t ssb_search_lexical(ssb_index* ix, const ssb_lex_batch* q, uint32_t k, uint32_t result_type, ssb_hit* hits, uint32_t* n_hits,
                           uint64_t* count_total) {
    if (!ix || !q || (q->n_queries && k && result_type != SSB_RESULT_COUNT && !hits)) { set_error("ssb_search_lexical: null argument"); return SSB_E_INVALID; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    ix->stats = ssb_stats{};
    const uint32_t nq = q->n_queries;
    if (nq == 0) return SSB_OK;
    if (k > SSB_K_LIMIT) { set_error("k=%u exceeds SSB_K_LIMIT=%u", k, SSB_K_LIMIT); return SSB_E_UNSUPPORTED; }
    SSB_TRY(ix->keys_a.reserve((size_t)nq * LIST, 0, ix->st));
    SSB_TRY(ix->counts.reserve(nq, 0, ix->st));
    ix->h_keys_a.resize((size_t)nq * LIST); ix->h_counts.resize(nq);
    const bool want_hits = hits && k && result_type != SSB_RESULT_COUNT;
    const uint32_t k1 = k < SSB_K_MAX ? k : SSB_K_MAX;
    SSB_TRY(ix->lex->search_keys(q, k1, result_type, ix->keys_a.p, ix->counts.p, &ix->stats.kernel_launches));
    SSB_CUDA_TRY(cudaMemcpyAsync(ix->h_keys_a.data(), ix->keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, ix->st));
    SSB_CUDA_TRY(cudaMemcpyAsync(ix->h_counts.data(), ix->counts.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, ix->st));
    SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
    ix->stats.d2h_bytes += (uint64_t)nq * (LIST * 8 + 8);
    if (want_hits) {
        PageState ps(ix, nq, k, hits, n_hits);
        bool more = ps.append(ix->h_keys_a.data(), 0, k1);
        // pages beyond the first 32 results: Topk search restricted to keys below the previous page's last key
        for (uint32_t done = k1; more && done < k; done += SSB_K_MAX) {
            const uint32_t kk = k - done < SSB_K_MAX ? k - done : SSB_K_MAX;
            SSB_TRY(ps.upload_ceilings());
            SSB_TRY(ix->lex->search_keys(q, kk, SSB_RESULT_TOPK, ix->keys_a.p, nullptr, &ix->stats.kernel_launches, ix->ceil.p));
            SSB_CUDA_TRY(cudaMemcpyAsync(ix->h_keys_a.data(), ix->keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, ix->st));
            SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
            ix->stats.d2h_bytes += (uint64_t)nq * LIST * 8;
            more = ps.append(ix->h_keys_a.data(), done, kk);
        }
        ps.finish();
    } else if (n_hits) for (uint32_t i = 0; i < nq; i++) n_hits[i] = 0;
    if (count_total) for (uint32_t i = 0; i < nq; i++) count_total[i] = ix->h_counts[i];
    LexStats ls = ix->lex->last_stats();
    ix->stats.postings_visited = ls.postings_visited;
    ix->stats.algorithmic_bytes = ls.postings_visited * 4 + ls.probes * 12 + ls.items_processed * 24;
    ix->stats.probes = ls.probes; ix->stats.items_processed = ls.items_processed; ix->stats.items_skipped = ls.items_skipped;
    ix->ev_used = true;
    return SSB_OK;
}

int32_t ssb_rrf_fuse(const ssb_hit* lex, uint32_t n_lex, const ssb_hit* vec, uint32_t n_vec, ssb_hit* out, uint32_t* n_out) {
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
}

int32_t ssb_search_hybrid(ssb_index* ix, const ssb_lex_batch* q, const float* queries, uint32_t k, ssb_hit* hits, uint32_t* n_hits) {
    if (!ix || !q || !queries || !hits) { set_error("ssb_search_hybrid: null argument"); return SSB_E_INVALID; }
    if (k == 0 || k > SSB_K_MAX) { set_error("k must be in 1..%u", SSB_K_MAX); return SSB_E_UNSUPPORTED; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    ix->stats = ssb_stats{};
    const uint32_t nq = q->n_queries;
    if (nq == 0) return SSB_OK;
    SSB_TRY(ix->keys_a.reserve((size_t)nq * LIST, 0, ix->st));
    SSB_TRY(ix->keys_b.reserve((size_t)nq * LIST, 0, ix->st));
    SSB_TRY(ix->lex->search_keys(q, k, SSB_RESULT_TOPK, ix->keys_a.p, nullptr, &ix->stats.kernel_launches));
    SSB_TRY(vec_keys(ix, queries, nq, k, ix->keys_b.p));
    ix->h_keys_a.resize((size_t)nq * LIST); ix->h_keys_b.resize((size_t)nq * LIST);
    SSB_CUDA_TRY(cudaMemcpyAsync(ix->h_keys_a.data(), ix->keys_a.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, ix->st));
    SSB_CUDA_TRY(cudaMemcpyAsync(ix->h_keys_b.data(), ix->keys_b.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, ix->st));
    SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
    ix->stats.d2h_bytes += (uint64_t)nq * LIST * 16;
    std::vector<ssb_hit> a(k), b(k), f(2 * (size_t)k);
    for (uint32_t i = 0; i < nq; i++) {
        uint32_t na = 0, nb = 0, nf = 0;
        decode_keys(ix->h_keys_a.data() + (size_t)i * LIST, 1, k, a.data(), &na);
        decode_keys(ix->h_keys_b.data() + (size_t)i * LIST, 1, k, b.data(), &nb);
        ssb_rrf_fuse(a.data(), na, b.data(), nb, f.data(), &nf);
        uint32_t n = nf < k ? nf : k;     // search.rs:2117-2119 truncate(length)
        for (uint32_t j = 0; j < k; j++) hits[(size_t)i * k + j] = j < n ? f[j] : ssb_hit{0, 0.f, 0};
        if (n_hits) n_hits[i] = n;
    }
    return SSB_OK;
}

int32_t ssb_merge_keys(ssb_index* ix, const uint64_t* keys_dev, uint32_t n_lists, uint32_t nq, uint32_t k, ssb_hit* hits, uint32_t* n_hits) {
    if (!ix || !keys_dev || !hits) { set_error("ssb_merge_keys: null argument"); return SSB_E_INVALID; }
    if (k == 0 || k > SSB_K_MAX) { set_error("k must be in 1..%u", SSB_K_MAX); return SSB_E_UNSUPPORTED; }
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    if (nq == 0) return SSB_OK;
    SSB_TRY(ix->keys_b.reserve((size_t)nq * LIST, 0, ix->st));
    SSB_TRY(vec::launch_merge_lists(keys_dev, n_lists, nq, ix->keys_b.p, ix->st));
    ix->stats.kernel_launches += 1;
    ix->h_keys_b.resize((size_t)nq * LIST);
    SSB_CUDA_TRY(cudaMemcpyAsync(ix->h_keys_b.data(), ix->keys_b.p, (size_t)nq * LIST * 8, cudaMemcpyDeviceToHost, ix->st));
    SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
    decode_keys(ix->h_keys_b.data(), nq, k, hits, n_hits);
    return SSB_OK;
}

int32_t ssb_sync(ssb_index* ix) {
    if (!ix) return SSB_E_INVALID;
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
    return SSB_OK;
}

void* ssb_stream(ssb_index* ix) { return ix ? (void*)ix->st : nullptr; }

int32_t ssb_set_stream(ssb_index* ix, void* stream) {
    if (!ix) return SSB_E_INVALID;
    std::lock_guard<std::mutex> g(ix->mu);
    SSB_CUDA_TRY(cudaSetDevice(ix->cfg.device));
    SSB_CUDA_TRY(cudaStreamSynchronize(ix->st));
    ix->st = stream == SSB_OWN_STREAM ? ix->own_st : (cudaStream_t)stream;
    ix->lex->set_stream(ix->st);
    return SSB_OK;
}

int32_t ssb_last_stats(const ssb_index* ix, ssb_stats* out) {
    if (!ix || !out) return SSB_E_INVALID;
    *out = ix->stats;
    if (ix->ev_used) {
        cudaSetDevice(ix->cfg.device);
        float ms = 0.f;
        if (cudaEventSynchronize(ix->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ix->ev0, ix->ev1) == cudaSuccess)
            out->dominant_kernel_ns = (uint64_t)((double)ms * 1e6);
        else cudaGetLastError();
    }
    return SSB_OK;
}

}  // extern "C"
