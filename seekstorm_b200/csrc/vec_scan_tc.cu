// vec_scan_tc.cu — tensor-core variant of the brute-force vector scan (sm_100a, tcgen05 + TMEM + TMA).
//
// Same contract as scan_ffma (vec_scan.cu): corpus [n_rows, Dpad] f32 x a block of NQ queries -> per-CTA
// top-32 lists, but the query x corpus contraction runs on the 5th-gen tensor cores:
//   D[128 corpus rows, NQ queries] (f32, TMEM) += A[128 x 8] . B[NQ x 8]^T      tcgen05.mma kind::tf32
// Each smem stage holds MT=2 corpus tiles (256 rows) against ONE query chunk, so the query operand, which is
// re-fetched from L2 for every stage, costs half the L2->SM bandwidth (the measured limiter of the 1-tile version).
// f32 accuracy (north-star tolerance 1e-4 on cosine scores) is kept with the 3xTF32 split
//   a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo,   x_hi = x & 0xFFFFE000 (exactly representable in tf32),
//   x_lo = x - x_hi (exact in f32); the dropped a_lo.b_lo term is ~2^-22 relative.
// B_hi / B_lo are prepared once per batch in global memory; A_hi / A_lo are produced per stage in shared
// memory by 4 "splitter" warps (the corpus is stored once, as f32 — algorithmic bytes stay n_rows*dims*4).
//
// Warp roles (384 threads, 1 CTA/SM): w0 TMA producer | w1 TMEM alloc + MMA issuer | w4-7 epilogue (TMEM
// lane quadrant = warp%4) | w8-11 splitters.  Pipelines: full/split/empty per smem stage, tfull/tempty per
// TMEM accumulator buffer (double-buffered: the epilogue of tile t overlaps the MMAs of tile t+1).
// Epilogue: tcgen05.ld 8 query columns at a time, filter against the per-query threshold, warp-aggregated
// push into per-query buckets, per-query sorted lists (smem) updated by warp-shuffle insertion.
#include "common.cuh"
#include "vec_scan.h"

namespace ssb {
namespace vec {

namespace tc {
constexpr int KC = 32;                 // floats per k-chunk = one 128-byte swizzle row
constexpr int TM = 128;                // UMMA M
constexpr int MT = 2;                  // M-tiles per stage: the B (query) chunk is fetched once per MT*128 corpus rows
constexpr int TROWS = TM * MT;         // corpus rows per stage
constexpr int A1_BYTES = TM * KC * 4;  // one 128-row swizzled tile, 16 KB
constexpr int A_BYTES = MT * A1_BYTES; // 32 KB
constexpr int THREADS = 384;
constexpr int CHUNK = 8;               // query columns per epilogue step

template <int NQ> struct Cfg {
    static constexpr int B_BYTES = NQ * KC * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;          // A, A_lo, B_hi, B_lo
    static constexpr int STAGES = 2;                                       // 2 x (64 KB A/A_lo + 2*B) 
    static constexpr int TX_BYTES = A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + NQ * 4 + 256;
    static constexpr int TMEM_COLS = 2 * MT * NQ;                          // double-buffered MT accumulators (256 / 512 columns)
};

__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
    // K-major, SWIZZLE_128B canonical layout ((8,n),2):((8,SBO),1) in 16-byte units: LBO = 1, SBO = 1024 B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int NQ>
__global__ void __launch_bounds__(THREADS, 1)
scan_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
        const __grid_constant__ CUtensorMap tmBl, uint32_t n_rows, uint32_t n_kchunks, uint32_t n_tiles, uint32_t k,
        const uint32_t* __restrict__ doc_ids, uint64_t* __restrict__ scratch /*[gridDim.y][gridDim.x*4][NQ][32]*/,
        const uint32_t* __restrict__ thr_init /*[gridDim.y*NQ] or null*/) {
    using C = Cfg<NQ>;
    constexpr int STAGES = C::STAGES;
    // no static shared memory: the dynamic segment starts at offset 0 of the CTA window (1024-aligned for SWIZZLE_128B)
    // and pointers derived from it stay in the shared address space (LDS/STS instead of generic LD/ST)
    extern __shared__ __align__(1024) uint8_t base[];
    uint8_t* stage0 = base;
    // per-query sorted lists live directly in this CTA's slice of the output scratch (global, L2-resident): they are
    // touched only on the rare candidate insert, and each query is always owned by the same warp
    uint64_t* lists = scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 * NQ * LIST;   // [4 epilogue warps][NQ][32]
    float* thr_s = (float*)(base + STAGES * C::STAGE_BYTES);                  // [NQ] (used as ordered-uint thresholds)
    uint64_t* bars = (uint64_t*)(thr_s + NQ);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* split = full + STAGES;       // [STAGES]
    uint64_t* empty = split + STAGES;      // [STAGES]
    uint64_t* tfull = empty + STAGES;      // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t group = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&split[s], 4); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < NQ; i += THREADS)   // ordered-uint thresholds, seeded by the pre-sample pass when present
        reinterpret_cast<uint32_t*>(thr_s)[i] = thr_init ? __ldg(&thr_init[blockIdx.y * NQ + i]) : 0u;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *(volatile uint32_t*)tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                    uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    uint8_t* st = stage0 + s * C::STAGE_BYTES;
                    mbar_wait(&empty[s], ph ^ 1u);
#ifdef TC_DBG_NOB
                    mbar_arrive_expect_tx(&full[s], it < (uint32_t)STAGES ? C::TX_BYTES : A_BYTES);
#else
                    mbar_arrive_expect_tx(&full[s], C::TX_BYTES);
#endif
                    tma_load_2d(st, &tmA, (int)(kc * KC), (int)(tile * TROWS), &full[s]);   // 256-row box = MT swizzled tiles
#ifdef TC_DBG_NOB
                    if (it < (uint32_t)STAGES) {
#endif
                    tma_load_2d(st + 2 * A_BYTES, &tmBh, (int)(kc * KC), (int)(group * NQ), &full[s]);
                    tma_load_2d(st + 2 * A_BYTES + C::B_BYTES, &tmBl, (int)(kc * KC), (int)(group * NQ), &full[s]);
#ifdef TC_DBG_NOB
                    }
#endif
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NQ >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            uint32_t it = 0, ti = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
                const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
                mbar_wait(&tempty[buf], tph ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * (MT * NQ);
                for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                    uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    mbar_wait(&split[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage0 + s * C::STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_k128(sa), a_lo = umma_desc_k128(sa + A_BYTES);
                    const uint64_t b_hi = umma_desc_k128(sa + 2 * A_BYTES), b_lo = umma_desc_k128(sa + 2 * A_BYTES + C::B_BYTES);
#pragma unroll
                    for (uint32_t m = 0; m < MT; m++) {
                        const uint64_t am = (uint64_t)((m * A1_BYTES) >> 4);   // next 128-row tile of the stage
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; kk++) {        // 4 x K=8 (32 bytes) inside the 128-byte swizzle row
                            const uint64_t o = (uint64_t)(kk * 2);  // +32 bytes in 16-byte units
                            umma_tf32(d + m * NQ, a_hi + am + o, b_hi + o, idesc, (kc | kk) != 0);
#ifndef TC_DBG_1MMA
                            umma_tf32(d + m * NQ, a_lo + am + o, b_hi + o, idesc, 1);
                            umma_tf32(d + m * NQ, a_hi + am + o, b_lo + o, idesc, 1);
#endif
                        }
                    }
                    umma_commit(&empty[s]);                      // stage reusable once these MMAs retire
                    if (kc + 1 == n_kchunks) umma_commit(&tfull[buf]);
                }
            }
        }
    } else if (warp >= 8) {
        // ===================== splitters: A -> (A_hi in place, A_lo) =====================
        const int t = threadIdx.x - 256;   // 0..127
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                uint4* A = (uint4*)(stage0 + s * C::STAGE_BYTES);
                uint4* Al = (uint4*)(stage0 + s * C::STAGE_BYTES + A_BYTES);
#ifdef TC_DBG_NOSPLIT
                if (false)
#endif
#pragma unroll
                for (int j = 0; j < (A_BYTES / 16) / 128; j++) {
                    uint4 x = A[t + 128 * j], h, l;
                    h.x = x.x & 0xFFFFE000u; h.y = x.y & 0xFFFFE000u; h.z = x.z & 0xFFFFE000u; h.w = x.w & 0xFFFFE000u;
                    l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x));
                    l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y));
                    l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z));
                    l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w));
                    A[t + 128 * j] = h;
                    Al[t + 128 * j] = l;
                }
                fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> filter -> per-warp per-query top-k (no CTA-level barriers) =====================
        // Each epilogue warp owns the 32 TMEM lanes (= corpus rows) of its quadrant and keeps its own sorted list per
        // query in this CTA's slice of the output scratch.  The per-query threshold (ordered-uint score of the best
        // k-th entry any warp of the CTA has seen) is shared through smem with atomicMax: monotone, so stale reads only
        // cost an extra insert.  (An earlier version synchronised the 4 warps with two named barriers per 8-query chunk;
        // measured cost ~800 cycles per barrier — 40 % of the kernel.)
        const int ew = warp - 4;                          // == warp % 4 == TMEM lane quadrant
        uint32_t* thr_u = reinterpret_cast<uint32_t*>(thr_s);
        uint64_t* mylists = lists + (size_t)ew * NQ * LIST;
        for (int i = lane; i < NQ * LIST; i += 32) mylists[i] = 0;
        __syncwarp();
        uint32_t ti = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
            const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
            mbar_wait(&tfull[buf], tph);
            tc_fence_after();
            for (int mc = 0; mc < MT * (NQ / CHUNK); mc++) {
                const int m = mc / (NQ / CHUNK), c = mc % (NQ / CHUNK);
                const uint32_t row = tile * TROWS + (uint32_t)(m * TM) + (uint32_t)(ew * 32 + lane);
                const bool valid = row < n_rows;
                uint32_t v[CHUNK];
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * (MT * NQ) + m * NQ + c * CHUNK;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < CHUNK; j++) {
                    const int q = c * CHUNK + j;
                    const float sc = __uint_as_float(v[j]);
                    const uint32_t so = ord_f32(sc);
                    const bool pass = valid && sc == sc && so >= thr_u[q];
                    unsigned pm = __ballot_sync(FULL, pass);
                    if (pm) {                                           // rare after warm-up
                        uint64_t key = 0;
                        if (pass) key = ((uint64_t)so << 32) | (uint64_t)(0xFFFFFFFFu - (doc_ids ? __ldg(&doc_ids[row]) : row));
                        uint64_t L = mylists[q * LIST + lane];
                        if (__popc(pm) > 3) {
                            // bulk (warm-up tiles: every row passes): sort the 32 keys, bitonic-merge into the list
                            L = wl_merge(L, wl_sort_desc(key, lane), lane);
                        } else {
                            while (pm) { const int src = __ffs(pm) - 1; pm &= pm - 1; wl_insert(L, shfl64(key, src), lane); }
                        }
                        mylists[q * LIST + lane] = L;
                        const uint32_t kth = (uint32_t)(shfl64(L, (int)k - 1) >> 32);
                        if (lane == 0 && kth > thr_u[q]) atomicMax(&thr_u[q], kth);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// [nq_pad][dpad] f32 -> hi / lo tf32 parts
__global__ void split_queries(const float* __restrict__ q, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = __float_as_uint(q[i]), h = x & 0xFFFFE000u;
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(x) - __uint_as_float(h);
}

}  // namespace tc

template <int NQ>
static int32_t launch_tc_n(const ScanArgs& a, cudaStream_t st) {
    using C = tc::Cfg<NQ>;
    CUtensorMap tmA, tmBh, tmBl;
    uint32_t n_tiles = (uint32_t)((a.n_rows + tc::TROWS - 1) / tc::TROWS);
    uint32_t n_groups = a.nq_pad / NQ;
    SSB_TRY(encode_tmap_2d_f32(&tmA, a.rows, a.dpad, a.n_rows, (uint64_t)a.dpad * 4, tc::KC, tc::TROWS, 1));
    SSB_TRY(encode_tmap_2d_f32(&tmBh, a.q_hi, a.dpad, a.nq_pad, (uint64_t)a.dpad * 4, tc::KC, NQ, 1));
    SSB_TRY(encode_tmap_2d_f32(&tmBl, a.q_lo, a.dpad, a.nq_pad, (uint64_t)a.dpad * 4, tc::KC, NQ, 1));
    uint32_t gx = n_tiles < (uint32_t)a.n_sms ? n_tiles : (uint32_t)a.n_sms;
    if ((size_t)n_groups * gx * 4 * NQ * LIST * 8 > a.scratch_bytes) { set_error("vector scan scratch too small"); return SSB_E_STATE; }
    static bool attr_set = false;
    if (!attr_set) {
        SSB_CUDA_TRY(cudaFuncSetAttribute(tc::scan_tc<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    size_t nel = (size_t)a.nq_pad * a.dpad;
    tc::split_queries<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(a.queries_padded, a.q_hi, a.q_lo, nel);
    if (a.ev0) cudaEventRecord(a.ev0, st);
    tc::scan_tc<NQ><<<dim3(gx, n_groups), tc::THREADS, C::SMEM, st>>>(tmA, tmBh, tmBl, (uint32_t)a.n_rows, a.dpad / tc::KC, n_tiles,
                                                                       a.k, a.doc_ids, a.scratch, a.thr_init);
    if (a.ev1) cudaEventRecord(a.ev1, st);
    SSB_CUDA_TRY(cudaGetLastError());
    // scratch layout [group][list][q in NQ][32] -> generic merge with qt = NQ
    merge_lists_generic(a.scratch, gx * 4, NQ, a.nq_pad, a.keys_out, st);
    SSB_CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

static int32_t launch_scan_tc_impl(const ScanArgs& a, uint32_t nq_tile, cudaStream_t st);

int32_t launch_scan_tc(const ScanArgs& a, uint32_t nq_tile, cudaStream_t st) {
    // threshold pre-sampling (see vec_scan.cu): scan the first rows, seed the thresholds, then the full scan
    if (a.thr_init || !a.thr_buf || vec_presample_rows(a.n_rows) == 0) return launch_scan_tc_impl(a, nq_tile, st);
    ScanArgs pre = a;
    pre.n_rows = vec_presample_rows(a.n_rows); pre.ev0 = nullptr; pre.ev1 = nullptr; pre.thr_buf = nullptr;
    SSB_TRY(launch_scan_tc_impl(pre, nq_tile, st));
    launch_kth_threshold(a.keys_out, a.nq_pad, a.k, a.thr_buf, st);
    ScanArgs full = a;
    full.thr_init = a.thr_buf;
    return launch_scan_tc_impl(full, nq_tile, st);
}

static int32_t launch_scan_tc_impl(const ScanArgs& a, uint32_t nq_tile, cudaStream_t st) {
    if (a.n_rows == 0 || a.nq_pad == 0) return SSB_OK;
    if (a.similarity == SSB_SIM_EUCLIDEAN) { set_error("tcgen05 scan supports Dot/Cosine only"); return SSB_E_UNSUPPORTED; }
    if ((nq_tile != 64 && nq_tile != 128) || a.nq_pad % nq_tile != 0) { set_error("tcgen05 scan: query count must be padded to the 64/128 query tile"); return SSB_E_INVALID; }
    return nq_tile == 64 ? launch_tc_n<64>(a, st) : launch_tc_n<128>(a, st);
}

size_t scan_tc_scratch_bytes(int n_sms, uint32_t nq_pad) { return (size_t)nq_pad * (size_t)n_sms * 4 * LIST * 8; }

}  // namespace vec
}  // namespace ssb
