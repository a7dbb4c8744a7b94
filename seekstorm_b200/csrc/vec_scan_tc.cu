// vec_scan_tc.cu — tensor-core variant of the brute-force vector scan (sm_100a, tcgen05 + TMEM + TMA).
//
// Same contract as scan_ffma (vec_scan.cu): corpus [n_rows, Dpad] f32 x a block of NQ queries -> per-warp
// top-32 lists, but the query x corpus contraction runs on the 5th-gen tensor cores:
//   D[128 corpus rows, NQ queries] (f32, TMEM) += A[128 x K] . B[NQ x K]^T          tcgen05.mma.cta_group::1
// f32-level accuracy (north-star tolerance 1e-4 on cosine scores) comes from a 3-product split
//   a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo                (the dropped a_lo.b_lo term is second order)
// in one of two operand precisions, or exactly on int8 codes (template PREC):
//   PREC_TF32 : kind::tf32, x_hi = x & 0xFFFFE000 (exactly representable in tf32, so the result does not depend on
//               whether the tensor core truncates or rounds), x_lo = x - x_hi (exact); error ~2^-22 per product.
//   PREC_BF16 : kind::f16 (bf16 operands, f32 accumulate), x_hi = bf16_rn(x), x_lo = bf16_rn(x - x_hi);
//               error ~3*2^-17 per product (random sign) -> ~1e-5 relative on a 768-d score.  Twice the MACs per MMA
//               instruction and half the operand bytes per MAC of tf32: the faster split (measured), now within a few
//               percent of the tensor-pipe floor of 12 MMAs per 256-row stage.
//   PREC_I8   : kind::i8 over an int8 corpus (Cosine + ScalarQuantizationI8, quantised at load time): ONE exact product, s32
//               accumulators; the TMA'd tile is the MMA operand (no splitters), the 128-query block stays resident in smem
//               (template BRES) and 12 warps run the epilogue.  See DESIGN.md §3.2b.
// The query parts are prepared once per batch in global memory.  PREC_BF16: the corpus parts are prepared ONCE AT LOAD TIME as
// two bf16 planes (hi, lo) — together 4 bytes per element, so the algorithmic bytes of a pass stay n_rows*dims*4 — and TMA
// delivers them straight into the MMA-ready SWIZZLE_64B tiles: no splitter warps, no generic-proxy pass over the tile, half the
// shared-memory traffic per stage (round 1 split the f32 tile in shared memory per stage: ncu showed LSU wavefronts at 47 % of
// peak with 49.5 M bank conflicts competing with the tensor core's operand reads for the 128 B/clk of the SM's shared memory,
// and the pass ran at 1.35x the HBM floor).  48 KB per stage, 4 stages.  PREC_TF32 keeps the f32 corpus and the 8 splitter
// warps (out of place, 2 stages).  Each smem stage holds MT=2 corpus tiles (256 rows) against ONE query chunk.
//
// Warp roles (512 threads, 1 CTA/SM): w0 TMA producer | w1 TMEM alloc + MMA issuer | w4-7 epilogue (TMEM lane quadrant =
// warp%4) | w8-15 splitters (tf32) / additional epilogue warps (int8) / idle (bf16).  Pipelines: full/split/empty per smem stage,
// tfull/tempty per TMEM accumulator buffer (double-buffered: the epilogue of tile t overlaps the MMAs of tile t+1).
// Epilogue: tcgen05.ld 8 (int8: 16) query columns at a time; the whole chunk is tested against the per-query thresholds
// branch-free with ONE vote, and only chunks with a candidate take the per-column path (ballot, per-warp sorted lists in the
// CTA's slice of the output scratch, no CTA barriers).
// Threshold seeding: the same kernel runs first in sample mode over one tile per SM and writes per-(32-row group, query)
// score maxima; kth_from_groupmax turns them into valid lower bounds of the k-th best score.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "vec_scan.h"

namespace ssb {
namespace vec {

namespace tc {
constexpr int KC = 32;                 // floats per k-chunk = one 128-byte swizzle row of the f32 corpus tile
constexpr int TM = 128;                // UMMA M
constexpr int MT = 2;                  // M-tiles per stage: the query chunk is fetched once per MT*128 corpus rows
constexpr int TROWS = TM * MT;         // corpus rows per stage
constexpr int A1_BYTES = TM * KC * 4;  // one 128-row f32 tile, 16 KB
constexpr int A_BYTES = MT * A1_BYTES; // 32 KB
constexpr int THREADS = 512;            // 16 warps: TMA, MMA, 2 idle, 4 epilogue, 8 splitters
constexpr int SPLIT_THREADS = 256;
constexpr int CHUNK = 8;               // query columns per epilogue step
enum { PREC_TF32 = 0, PREC_BF16 = 1, PREC_I8 = 2, PREC_F16F = 3 };
// PREC_F16F — the FILTER scan: ONE product h(a).h(b) over an fp16 plane of the corpus (2 bytes per element, a third of
// the tensor work of the 3-product split; fp16 keeps 11 significand bits, bf16 8 — the margin below is 8x tighter than with the bf16
// hi plane; rows and queries are pre-scaled by powers of two into fp16's range, everything below lives in that scaled space and is
// never returned).  The approximate score s^ differs from the (scaled) f32 score s by at most eps_q =
// max_r|a_r - h(a_r)| * |b| + max_r|h(a_r)| * |b - h(b)| + accumulation slack (Cauchy-Schwarz; both row maxima are computed at load time),
// so every row of the exact top-k satisfies s^ >= (k-th best s^) - 2 eps_q.  The epilogue keeps exactly that candidate set (keys carry
// the ROW index, thresholds are lowered by the per-query margin 2 eps_q), refine_candidates (vec_refine.cu) re-scores the <= 32 candidates
// in f32 from the f32 rows and flags the queries whose candidate set may not have fitted the 32-entry list for the exact fallback scan.

constexpr int MAX_STAGES = 6;
template <int NQ, int PREC, bool BRES = false> struct Cfg {
    // NQ = 256 (bf16 only): ONE 128-row corpus tile per stage against 256 queries — the corpus is streamed once per 256 queries (half
    // the HBM bytes per query) and an MMA reads 4 KB of A per 8 KB of B (6 KB per 128x128x16 instead of 8); TMEM: 2 x 1 x 256 columns
    static constexpr int MT = NQ == 256 ? 1 : 2;
    static constexpr int TROWS = TM * MT;
    static constexpr int A_BYTES = MT * A1_BYTES;
    // TF32: [A (-> A_hi in place) | A_lo | B_hi | B_lo], all f32 SWIZZLE_128B tiles
    // BF16: [A1 bf16 (hi plane) | A2 bf16 (lo plane) | B1 bf16 | B2 bf16], 64-byte rows, SWIZZLE_64B, all four delivered by TMA.
    //       48 KB per stage -> 4 stages.
    // I8  : [A i8 | B i8]: the int8 corpus (quantised at load time) is the MMA operand as TMA delivers it — no splitter
    //       pass, 128 dims per 128-byte swizzle row, 4 smem stages
    static constexpr int STAGES = PREC == PREC_TF32 ? 2 : 4;
    static constexpr int KCE = PREC == PREC_I8 ? 128 : (PREC == PREC_F16F ? 64 : KC);   // elements per k-chunk (128 bytes; bf16 3-product: 64-byte rows)
    static constexpr int B_BYTES = PREC == PREC_BF16 ? NQ * KC * 2 : NQ * KC * 4;
    static constexpr int ALO_OFF = PREC == PREC_BF16 ? 0 : A_BYTES;            // TF32: A_lo   | BF16: A1 (in place)
    static constexpr int A2_OFF = A_BYTES / 2;                                 // BF16: A2 (in place)
    static constexpr int B_OFF = PREC == PREC_TF32 ? 2 * A_BYTES : A_BYTES;
    // BRES (int8 only): the whole quantised query block [n_kchunks][NQ x 128 B] stays resident in smem behind the A stages,
    // so the per-stage L2->SM traffic is the corpus tile alone (stage count chosen at launch from what is left of 227 KB)
    static constexpr int STAGE_BYTES = BRES ? A_BYTES : ((PREC == PREC_I8 || PREC == PREC_F16F) ? A_BYTES + B_BYTES : (PREC == PREC_BF16 ? A_BYTES + 2 * B_BYTES : 2 * A_BYTES + 2 * B_BYTES));
    static constexpr int TX_BYTES = BRES ? A_BYTES : ((PREC == PREC_I8 || PREC == PREC_F16F) ? A_BYTES + B_BYTES : A_BYTES + 2 * B_BYTES);
    static constexpr int SMEM = STAGES * STAGE_BYTES + NQ * 12 + 256;   // thresholds + (scaled int8) per-query scale / norm
    static constexpr int TMEM_COLS = 2 * MT * NQ;                              // double-buffered MT accumulators
};

// K-major SWIZZLE_128B canonical layout ((8,n),2):((8,SBO),1) in 16-byte units: LBO = 1, SBO = 1024 B
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major SWIZZLE_64B: 64-byte rows, 8-row atom = 512 B -> SBO = 512 B, layout type 4
__device__ __forceinline__ uint64_t umma_desc_k64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
template <int PREC>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (PREC == PREC_TF32)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    else if (PREC == PREC_I8)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint32_t bf16x2_hi(float x, float y, float& rx, float& ry) {
    // hi = bf16_rn(x); the residual x - hi is exact in f32
    __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    rx = x - __low2float(h); ry = y - __high2float(h);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t bf16x2(float x, float y) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
}

// int8 path: thresholds are kept as int32 dot products (scores are integer valued, so ord_f32((float)d) is monotone in d)
__device__ __forceinline__ int ord_to_int(uint32_t o) {
    return o == 0u ? INT_MIN : (o == 0xFFFFFFFFu ? INT_MAX : (int)unord_f32(o));
}

// SCALED (int8 only): 0 = Cosine + ScalarQuantizationI8 (score = the int32 dot product); 1 = Dot + ScalarQuantizationI8 (per-vector
// scale, score = dot_i32 as f32 * query_scale * row_scale, dot_i8_quantized vector_similarity.rs:1754-1758); 2 = Euclidean +
// ScalarQuantizationI8, non-affine (score = -max(0, query_norm + row_norm - 2*dot), euclidean_i8_quantized :1721-1734); 3 = the AFFINE
// variant (integer-valued 0..255 data, euclidean_i8_quantized_affine :1770-1795): the int32 dot product is first corrected for the two zero
// points, dot - zp_row*sum_q(query) - zp_query*sum_q(row) + n*zp_query*zp_row, regrouped as dot - zp_row*sum_q(query) + zp_query*(n*zp_row - sum_q(row))
template <int NQ, int PREC, bool BRES, int SCALED>
__global__ void __launch_bounds__(THREADS, 1)
scan_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2 /*bf16: lo plane*/,
        const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, uint32_t n_rows, uint32_t n_kchunks, uint32_t n_tiles, uint32_t k,
        const uint32_t* __restrict__ doc_ids, uint64_t* __restrict__ scratch /*[gridDim.y][gridDim.x*4][NQ][32]*/,
        const uint32_t* __restrict__ thr_init /*[gridDim.y*NQ] or null*/, uint32_t nq_valid,
        const uint64_t* __restrict__ ceil_keys /*[gridDim.y*NQ] or null*/, uint32_t nst_rt,
        const uint32_t* __restrict__ del_slot, const uint64_t* __restrict__ del_words /*delete set or null*/,
        const float* __restrict__ row_scale, const float* __restrict__ row_norm /*SCALED: [n_rows]*/,
        const float* __restrict__ q_scale, const float* __restrict__ q_norm /*SCALED: [gridDim.y*NQ]*/,
        const int2* __restrict__ row_aff /*SCALED 3: [n_rows] (zero_point, dims*zero_point - sum_q)*/, const int2* __restrict__ q_aff /*SCALED 3: [gridDim.y*NQ] (zero_point, sum_q)*/,
        const uint32_t* __restrict__ ivf_sel, uint32_t ivf_words, const uint32_t* __restrict__ row_cluster /*IVF selection mask (f32 epilogue) or null*/,
        uint32_t sample_mode /*int8 only: write per-(32-row group, query) score maxima instead of lists*/) {
    using C = Cfg<NQ, PREC, BRES>;
    constexpr int MT = C::MT, TROWS = C::TROWS, A_BYTES = C::A_BYTES;   // (shadow the namespace-level 2-tile defaults)
    const uint32_t STAGES = BRES ? nst_rt : (uint32_t)C::STAGES;
    // no static shared memory: the dynamic segment starts at offset 0 of the CTA window (1024-aligned for the swizzled
    // tiles) and pointers derived from it stay in the shared address space (LDS/STS instead of generic LD/ST)
    extern __shared__ __align__(1024) uint8_t base[];
    uint8_t* stage0 = base;
    // per-query sorted lists live directly in this CTA's slice of the output scratch (global, L2-resident): they are
    // touched only on the rare candidate insert, and each (warp, query) list is always owned by the same warp
    uint64_t* lists = scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 * NQ * LIST;   // [4 epilogue warps][NQ][32]
    uint8_t* bres = base + STAGES * C::STAGE_BYTES;                           // BRES: resident query block
    uint32_t* thr_u = (uint32_t*)(bres + (BRES ? n_kchunks * C::B_BYTES : 0)); // [NQ] ordered-uint score thresholds
    float* qs_sm = (float*)(thr_u + NQ);          // [NQ] SCALED: query scale
    float* qn_sm = qs_sm + NQ;                    // [NQ] SCALED: query norm
    uint64_t* bars = (uint64_t*)(thr_u + 3 * NQ);
    uint64_t* full = bars;                     // [STAGES]
    uint64_t* split = full + MAX_STAGES;       // [STAGES]
    uint64_t* empty = split + MAX_STAGES;      // [STAGES]
    uint64_t* tfull = empty + MAX_STAGES;      // [2]
    uint64_t* tempty = tfull + 2;              // [2]
    uint64_t* bfull = tempty + 2;              // BRES: query block landed
    uint32_t* tmem_slot = (uint32_t*)(bfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t group = blockIdx.y;

    if (threadIdx.x == 0) {
        mbar_init(bfull, 1);
        for (uint32_t s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&split[s], SPLIT_THREADS / 32); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], PREC == PREC_TF32 ? 4 : 12); }
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < NQ; i += THREADS) {  // seeded by the pre-sample pass when present
        uint32_t t = blockIdx.y * NQ + i >= nq_valid ? 0xFFFFFFFFu   // zero-padded query slot: unreachable threshold
                                                     : (thr_init ? __ldg(&thr_init[blockIdx.y * NQ + i]) : 0u);
        if (PREC == PREC_I8 && !SCALED) t = (uint32_t)ord_to_int(t);   // unscaled int8 path compares the raw int32 dot products
        thr_u[i] = t;
        if (PREC == PREC_F16F) qs_sm[i] = __ldg(&q_scale[blockIdx.y * NQ + i]);   // filter scan: per-query margin 2 eps_q
        if (SCALED) { qs_sm[i] = __ldg(&q_scale[blockIdx.y * NQ + i]); qn_sm[i] = SCALED >= 2 ? __ldg(&q_norm[blockIdx.y * NQ + i]) : 0.f; }
    }
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
            tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
            uint32_t it = 0;
            if (BRES) {
                mbar_arrive_expect_tx(bfull, n_kchunks * C::B_BYTES);
                for (uint32_t kc = 0; kc < n_kchunks; ++kc)
                    tma_load_2d(bres + kc * C::B_BYTES, &tmBh, (int)(kc * C::KCE), (int)(group * NQ), bfull);
            }
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                    uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    uint8_t* st = stage0 + s * C::STAGE_BYTES;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full[s], C::TX_BYTES);
                    tma_load_2d(st, &tmA, (int)(kc * C::KCE), (int)(tile * TROWS), &full[s]);   // 256-row box = MT swizzled tiles
                    if (PREC == PREC_BF16) tma_load_2d(st + C::A2_OFF, &tmA2, (int)(kc * C::KCE), (int)(tile * TROWS), &full[s]);   // lo plane
                    if (!BRES) tma_load_2d(st + C::B_OFF, &tmBh, (int)(kc * C::KCE), (int)(group * NQ), &full[s]);
                    if (PREC == PREC_TF32 || PREC == PREC_BF16) tma_load_2d(st + C::B_OFF + C::B_BYTES, &tmBl, (int)(kc * C::KCE), (int)(group * NQ), &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            // instruction descriptor: D=f32 (bit 4), A/B format at bits 7/10 (tf32 = 2, bf16 = 1, f16 = 0), both K-major,
            // N>>3 at bit 17, M>>4 at bit 24
            // (kind::i8: D = s32 (2 at bit 4), A/B format 1 = signed int8)
            constexpr uint32_t fmt = PREC == PREC_TF32 ? 2u : (PREC == PREC_F16F ? 0u /*kind::f16: F16*/ : 1u);
            constexpr uint32_t cfmt = PREC == PREC_I8 ? 2u : 1u;
            const uint32_t idesc = (cfmt << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NQ >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            uint32_t it = 0, ti = 0;
            if (BRES) mbar_wait(bfull, 0);
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
                const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
                mbar_wait(&tempty[buf], tph ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * (MT * NQ);
                for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                    uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    if (PREC == PREC_TF32) mbar_wait(&split[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage0 + s * C::STAGE_BYTES);
                    if (PREC == PREC_I8 || PREC == PREC_F16F) {   // one product; 4 x (K = 32 bytes) inside the 128-byte swizzle row
                        const uint64_t a = umma_desc_k128(sa), b = umma_desc_k128(BRES ? smem_u32(bres + kc * C::B_BYTES) : sa + C::B_OFF);
#pragma unroll
                        for (uint32_t m = 0; m < MT; m++) {
                            const uint64_t am = (uint64_t)((m * A1_BYTES) >> 4);
#pragma unroll
                            for (uint32_t kk = 0; kk < 4; kk++)          // 4 x K=32 int8 (32 bytes) inside the 128-byte swizzle row
                                umma<PREC>(d + m * NQ, a + am + (uint64_t)(kk * 2), b + (uint64_t)(kk * 2), idesc, (kc | kk) != 0);
                        }
                    } else if (PREC == PREC_TF32) {
                        const uint64_t a_hi = umma_desc_k128(sa), a_lo = umma_desc_k128(sa + C::ALO_OFF);
                        const uint64_t b_hi = umma_desc_k128(sa + C::B_OFF), b_lo = umma_desc_k128(sa + C::B_OFF + C::B_BYTES);
#pragma unroll
                        for (uint32_t m = 0; m < MT; m++) {
                            const uint64_t am = (uint64_t)((m * A1_BYTES) >> 4);   // next 128-row tile of the stage
#pragma unroll
                            for (uint32_t kk = 0; kk < 4; kk++) {        // 4 x K=8 (32 bytes) inside the 128-byte swizzle row
                                const uint64_t o = (uint64_t)(kk * 2);  // +32 bytes in 16-byte units
                                umma<PREC>(d + m * NQ, a_hi + am + o, b_hi + o, idesc, (kc | kk) != 0);
                                umma<PREC>(d + m * NQ, a_lo + am + o, b_hi + o, idesc, 1);
                                umma<PREC>(d + m * NQ, a_hi + am + o, b_lo + o, idesc, 1);
                            }
                        }
                    } else {
                        const uint64_t a1 = umma_desc_k64(sa + C::ALO_OFF), a2 = umma_desc_k64(sa + C::A2_OFF);
                        const uint64_t b1 = umma_desc_k64(sa + C::B_OFF), b2 = umma_desc_k64(sa + C::B_OFF + C::B_BYTES);
#pragma unroll
                        for (uint32_t m = 0; m < MT; m++) {
                            const uint64_t am = (uint64_t)((m * (A1_BYTES / 2)) >> 4);   // bf16 tile = 8 KB
#pragma unroll
                            for (uint32_t kk = 0; kk < 2; kk++) {        // 2 x K=16 bf16 (32 bytes) inside the 64-byte swizzle row
                                const uint64_t o = (uint64_t)(kk * 2);
                                umma<PREC>(d + m * NQ, a1 + am + o, b1 + o, idesc, (kc | kk) != 0);
                                umma<PREC>(d + m * NQ, a2 + am + o, b1 + o, idesc, 1);
                                umma<PREC>(d + m * NQ, a1 + am + o, b2 + o, idesc, 1);
                            }
                        }
                    }
                    umma_commit(&empty[s]);                      // stage reusable once these MMAs retire
                    if (kc + 1 == n_kchunks) umma_commit(&tfull[buf]);
                }
            }
        }
    } else if (warp >= 8 && PREC == PREC_TF32) {
        // ===================== splitters (tf32 only): f32 corpus tile -> (hi, lo) operand tiles =====================
        const int t = threadIdx.x - 256;   // 0..SPLIT_THREADS-1
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                uint8_t* st = stage0 + s * C::STAGE_BYTES;
                if (PREC == PREC_TF32) {
                    uint4* A = (uint4*)st;
                    uint4* Al = (uint4*)(st + C::ALO_OFF);
#pragma unroll
                    for (int j = 0; j < (A_BYTES / 16) / SPLIT_THREADS; j++) {
                        uint4 x = A[t + SPLIT_THREADS * j], h, l;
                        h.x = x.x & 0xFFFFE000u; h.y = x.y & 0xFFFFE000u; h.z = x.z & 0xFFFFE000u; h.w = x.w & 0xFFFFE000u;
                        l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x));
                        l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y));
                        l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z));
                        l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w));
                        A[t + SPLIT_THREADS * j] = h;
                        Al[t + SPLIT_THREADS * j] = l;
                    }
                }
                fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
            }
        }
    } else if (warp >= 4 && PREC == PREC_I8 && !SCALED) {
        // ===================== int8 epilogue: 12 warps = 4 TMEM lane quadrants x 3 query-column groups =====================
        // With the int8 operands the MMA + smem side of a 128-query pass costs a fraction of the HBM time, so the epilogue
        // (one compare + ballot per (row, query) pair) is what has to keep up: the 8 warps that are splitters in the f32
        // variants join in.  Warp (quadrant ew, group g) owns the 16-query column chunks c = g, g+3, ... for both M tiles,
        // and with them the lists (ew, q) of those queries — the scratch layout is the same as in the 4-warp epilogue.
        const int ew = warp & 3, g = (warp - 4) >> 2;
        int* thr_i = (int*)thr_u;
        uint64_t* mylists = lists + (size_t)ew * NQ * LIST;
        if (!sample_mode)
            for (int c = g; c < NQ / 16; c += 3)
                for (int i = lane; i < 16 * LIST; i += 32) mylists[c * 16 * LIST + i] = 0;
        __syncwarp();
        // sample mode (threshold seeding): no lists.  Every (tile, m, quadrant) is a group of 32 rows; per query the group's
        // best score goes to gmax[query][group].  The k-th largest group maximum is a valid lower bound of the k-th best score
        // (k groups each hold a row at least that good) and, for k << #groups, nearly as tight as an exact k-th of the sample.
        int* gmax = (int*)scratch;                       // [gridDim.y * NQ][n_tiles * 8]
        const uint32_t n_rg = n_tiles * (MT * 4);
        uint32_t ti = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
            const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
            mbar_wait(&tfull[buf], tph);
            tc_fence_after();
            for (int m = 0; m < MT; m++) {
                const uint32_t row = tile * TROWS + (uint32_t)(m * TM) + (uint32_t)(ew * 32 + lane);
                const bool valid = row < n_rows;
                for (int c = g; c < NQ / 16; c += 3) {
                    uint32_t v[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * (MT * NQ) + m * NQ + c * 16;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (sample_mode) {
                        int keep = INT_MIN;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            int d = valid ? (int)v[j] : INT_MIN;
                            if (ceil_keys && valid) {   // paging: rows already returned by an earlier page do not count
                                const uint64_t key = ((uint64_t)ord_f32((float)d) << 32) | (uint64_t)(0xFFFFFFFFu - (doc_ids ? __ldg(&doc_ids[row]) : row));
                                if (key >= __ldg(&ceil_keys[blockIdx.y * NQ + c * 16 + j])) d = INT_MIN;
                            }
                            const int mx = __reduce_max_sync(FULL, d);
                            if (lane == j) keep = mx;
                        }
                        if (lane < 16)
                            gmax[(size_t)(blockIdx.y * NQ + c * 16 + lane) * n_rg + (tile * (MT * 4) + m * 4 + ew)] = keep;
                        continue;
                    }
                    {   // common case, branch-free: does ANY of the 32 x 16 scores reach its query's threshold?  (ncu: with a
                        // ballot + branch per column this loop, not HBM, set the tile period)
                        const int4* t4 = reinterpret_cast<const int4*>(thr_i + c * 16);
                        const int4 t0 = t4[0], t1 = t4[1], t2 = t4[2], t3 = t4[3];
                        bool any = ((int)v[0] >= t0.x) | ((int)v[1] >= t0.y) | ((int)v[2] >= t0.z) | ((int)v[3] >= t0.w) |
                                   ((int)v[4] >= t1.x) | ((int)v[5] >= t1.y) | ((int)v[6] >= t1.z) | ((int)v[7] >= t1.w) |
                                   ((int)v[8] >= t2.x) | ((int)v[9] >= t2.y) | ((int)v[10] >= t2.z) | ((int)v[11] >= t2.w) |
                                   ((int)v[12] >= t3.x) | ((int)v[13] >= t3.y) | ((int)v[14] >= t3.z) | ((int)v[15] >= t3.w);
                        if (!__any_sync(FULL, any && valid)) continue;
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int q = c * 16 + j;
                        const int d = (int)v[j];
                        const bool pass = valid && d >= thr_i[q];
                        unsigned pm = __ballot_sync(FULL, pass);
                        if (pm) {                                           // rare after warm-up
                            uint64_t key = 0;
                            if (pass) {
                                const uint32_t doc = doc_ids ? __ldg(&doc_ids[row]) : row;
                                key = ((uint64_t)ord_f32((float)d) << 32) | (uint64_t)(0xFFFFFFFFu - doc);   // dot_i8 as f32: exact below 2^24
                                if (doc_deleted(del_slot, del_words, doc)) key = 0;
                            }
                            if (ceil_keys || del_slot) {
                                if (ceil_keys) { const uint64_t ceil = __ldg(&ceil_keys[blockIdx.y * NQ + q]); if (key >= ceil) key = 0; }
                                pm = __ballot_sync(FULL, key != 0);
                                if (!pm) continue;
                            }
                            uint64_t L = mylists[q * LIST + lane];
                            if (__popc(pm) > 3) {
                                L = wl_merge(L, wl_sort_desc(key, lane), lane);
                            } else {
                                while (pm) { const int src = __ffs(pm) - 1; pm &= pm - 1; wl_insert(L, shfl64(key, src), lane); }
                            }
                            mylists[q * LIST + lane] = L;
                            const uint32_t kth = (uint32_t)(shfl64(L, (int)k - 1) >> 32);
                            if (lane == 0 && kth) atomicMax(&thr_i[q], ord_to_int(kth));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    } else if (warp >= 4 && (warp < 8 || PREC != PREC_TF32)) {
        // ===================== epilogue: TMEM -> filter -> per-warp per-query top-k (no CTA-level barriers) =====================
        // Each epilogue warp owns the 32 TMEM lanes (= corpus rows) of its quadrant and keeps its own sorted list per
        // query.  The per-query threshold (ordered-uint score of the best k-th entry any warp of the CTA has seen, seeded
        // by the pre-sample pass) is shared through smem with atomicMax: monotone, so stale reads only cost an extra
        // insert.  (An earlier version pushed candidates into per-query buckets with two CTA-wide named barriers per
        // 8-query chunk and inserted them one by one: the dependent-shuffle insert chain made the epilogue, not the
        // tensor pipe, the limiter.)
        // Except in the tf32 variant (warps 8-15 are its splitters) all 12 warps 4-15 run the epilogue: quadrant ew = warp % 4, column group
        // g = (warp - 4) / 4 owns the 8-query chunks c = g, g + 3, ... of both M tiles and with them the lists (ew, q) of those queries —
        // same scratch layout as with 4 warps.  A candidate insert is a dependent global (L2) load + shuffle insert + store, ~1 us of pure
        // latency for the warp: with one warp per quadrant those inserts, not the tensor pipe or HBM, set the tile period of the filter scan.
        constexpr int EG = PREC == PREC_TF32 ? 1 : 3;
        constexpr int NCH3 = (NQ / CHUNK + EG - 1) / EG * EG;   // chunk slots per M tile, rounded up to a multiple of EG so that mc % EG == c % EG
        const int ew = warp & 3, g = (warp - 4) >> 2;
        uint64_t* mylists = lists + (size_t)ew * NQ * LIST;
        if (!sample_mode)
            for (int c = g; c < NQ / CHUNK; c += EG)
                for (int i = lane; i < CHUNK * LIST; i += 32) mylists[c * CHUNK * LIST + i] = 0;
        __syncwarp();
        uint32_t* gmaxu = (uint32_t*)scratch;            // sample mode (see the int8 epilogue): ordered-uint group maxima
        const uint32_t n_rg = n_tiles * (MT * 4);
        uint32_t ti = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
            const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
            mbar_wait(&tfull[buf], tph);
            tc_fence_after();
            for (int mc = g; mc < MT * NCH3; mc += EG) {             // chunk c of either M tile belongs to warp group c % EG (list ownership)
                const int m = mc / NCH3, c = mc % NCH3;
                if (c >= NQ / CHUNK) continue;                        // padding slot of the rounded-up chunk count
                const uint32_t row = tile * TROWS + (uint32_t)(m * TM) + (uint32_t)(ew * 32 + lane);
                const bool valid = row < n_rows;
                uint32_t v[CHUNK];
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * (MT * NQ) + m * NQ + c * CHUNK;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (SCALED) {
                    // scaled int8: the accumulators are exact int32 dot products; the score is rebuilt with the reference's
                    // operation order so that it is bit-identical to the CPU path, then treated like an f32 score below
                    const float rs = valid ? __ldg(&row_scale[row]) : 0.f;
                    const float rn = (SCALED >= 2 && valid) ? __ldg(&row_norm[row]) : 0.f;
                    const int2 ra = (SCALED == 3 && valid) ? __ldg(&row_aff[row]) : make_int2(0, 0);
#pragma unroll
                    for (int j = 0; j < CHUNK; j++) {
                        int di = (int)v[j];
                        if (SCALED == 3) { const int2 qa = __ldg(&q_aff[blockIdx.y * NQ + c * CHUNK + j]); di = di - ra.x * qa.y + qa.x * ra.y; }
                        const float dotf = __fmul_rn(__fmul_rn((float)di, qs_sm[c * CHUNK + j]), rs);
                        const float sc = SCALED >= 2 ? -fmaxf(__fsub_rn(__fadd_rn(qn_sm[c * CHUNK + j], rn), __fmul_rn(2.0f, dotf)), 0.0f) : dotf;
                        v[j] = __float_as_uint(sc);
                    }
                }
                if (sample_mode) {
                    uint32_t keep = 0;
#pragma unroll
                    for (int j = 0; j < CHUNK; j++) {
                        const float sc = __uint_as_float(v[j]);
                        uint32_t so = (valid && sc == sc) ? ord_f32(sc) : 0u;
                        if (ceil_keys && so) {
                            const uint64_t key = ((uint64_t)so << 32) | (uint64_t)(0xFFFFFFFFu - (doc_ids ? __ldg(&doc_ids[row]) : row));
                            if (key >= __ldg(&ceil_keys[blockIdx.y * NQ + c * CHUNK + j])) so = 0u;
                        }
                        const uint32_t mx = __reduce_max_sync(FULL, so);
                        if (lane == j) keep = mx;
                    }
                    if (lane < CHUNK)
                        gmaxu[(size_t)(blockIdx.y * NQ + c * CHUNK + lane) * n_rg + (tile * (MT * 4) + m * 4 + ew)] = keep;
                    continue;
                }
                {   // common case, branch-free: one vote per 8-query chunk instead of one per column (a NaN score maps above every
                    // threshold here and is rejected by the per-column test below)
                    const uint4* t4 = reinterpret_cast<const uint4*>(thr_u + c * CHUNK);
                    const uint4 t0 = t4[0], t1 = t4[1];
                    const bool any = (ord_f32(__uint_as_float(v[0])) >= t0.x) | (ord_f32(__uint_as_float(v[1])) >= t0.y) |
                                     (ord_f32(__uint_as_float(v[2])) >= t0.z) | (ord_f32(__uint_as_float(v[3])) >= t0.w) |
                                     (ord_f32(__uint_as_float(v[4])) >= t1.x) | (ord_f32(__uint_as_float(v[5])) >= t1.y) |
                                     (ord_f32(__uint_as_float(v[6])) >= t1.z) | (ord_f32(__uint_as_float(v[7])) >= t1.w);
                    if (!__any_sync(FULL, any && valid)) continue;
                }
#pragma unroll
                for (int j = 0; j < CHUNK; j++) {
                    const int q = c * CHUNK + j;
                    const float sc = __uint_as_float(v[j]);
                    const uint32_t so = ord_f32(sc);
                    const bool pass = valid && sc == sc && so >= thr_u[q];
                    unsigned pm = __ballot_sync(FULL, pass);
                    if (pm) {                                           // rare after warm-up
                        uint64_t key = 0;
                        if (pass) {
                            const uint32_t doc = doc_ids ? __ldg(&doc_ids[row]) : row;
                            key = ((uint64_t)so << 32) | (uint64_t)(0xFFFFFFFFu - (PREC == PREC_F16F ? row : doc));   // filter scan: candidates are named by row
                            if (doc_deleted(del_slot, del_words, doc) || ivf_skipped(ivf_sel, ivf_words, blockIdx.y * NQ + q, row_cluster, row)) key = 0;
                        }
                        if (ceil_keys || del_slot || ivf_sel) {   // paging: keys >= ceil were returned by an earlier page (0 = exhausted)
                            if (ceil_keys) { const uint64_t ceil = __ldg(&ceil_keys[blockIdx.y * NQ + q]); if (key >= ceil) key = 0; }
                            pm = __ballot_sync(FULL, key != 0);
                            if (!pm) continue;
                        }
                        uint64_t L = mylists[q * LIST + lane];
                        if (__popc(pm) > 3) {
                            // bulk (warm-up tiles: many rows pass): sort the 32 keys, bitonic-merge into the list
                            L = wl_merge(L, wl_sort_desc(key, lane), lane);
                        } else {
                            while (pm) { const int src = __ffs(pm) - 1; pm &= pm - 1; wl_insert(L, shfl64(key, src), lane); }
                        }
                        mylists[q * LIST + lane] = L;
                        uint32_t kth = (uint32_t)(shfl64(L, (int)k - 1) >> 32);
                        if (PREC == PREC_F16F && kth) kth = ord_f32(__fsub_rd(unord_f32(kth), qs_sm[q]));   // candidates: s^ >= k-th best s^ - 2 eps_q
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

// ------------------------------------------------------------------------------------------------------------------------------------
// scan_tc2 — the 256-query filter scan on CTA PAIRS (tcgen05 cta_group::2, thread-block cluster of 2).
// The one-CTA 256-query pass is bounded by what an SM can take in from L2: per 64-dim stage 16 KB of corpus + 32 KB of the query block
// at ~64 B/clk (ncu: 374 us against an HBM time of 234 us).  A pair of SMs shares ONE copy of the query block: each CTA loads its own
// 128 corpus rows (A, 16 KB) and HALF of the 256 queries (B, 16 KB) per stage, the leader's single thread issues M = 256 x N = 256 MMAs
// that read both halves (the peer's through the pair's shared-memory path), and every SM takes in 32 KB per stage instead of 48.
// Pipelines as in scan_tc; what changes: both CTAs' TMA loads complete on the LEADER's full barrier (cta_group::2 loads, peer bit of the
// barrier address cleared), tcgen05.commit is multicast to both CTAs' empty / tfull barriers, the epilogue warps of BOTH CTAs arrive on
// the leader's tempty barrier, TMEM is allocated with cta_group::2.  Filter precision only (PREC_F16F, NQ = 256); the threshold-seeding
// sample pass stays on scan_tc<256, F16F>.
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;        // shared::cluster address -> the same offset in the pair's even (leader) CTA
struct Cfg2 {
    static constexpr int NQ = 256, NQH = 128;
    static constexpr int A_BYTES = TM * 128;           // 128 rows x 64 halves
    static constexpr int B_BYTES = NQH * 128;          // this CTA's half of the query block
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
    static constexpr int STAGES = 6;
    static constexpr int SMEM = STAGES * STAGE_BYTES + NQ * 12 + 256;
    static constexpr int TMEM_COLS = 2 * NQ;            // double-buffered [128 lanes x 256 columns] per CTA
};
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar_in_leader) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar_in_leader) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs once the MMAs issued so far retire
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
scan_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, uint32_t n_rows, uint32_t n_kchunks, uint32_t n_pairs, uint32_t k,
         const uint32_t* __restrict__ doc_ids, uint64_t* __restrict__ scratch /*[gridDim.y][gridDim.x*4][256][32]*/,
         const uint32_t* __restrict__ thr_init, uint32_t nq_valid, const uint32_t* __restrict__ del_slot, const uint64_t* __restrict__ del_words,
         const float* __restrict__ q_margin, const uint32_t* __restrict__ ivf_sel, uint32_t ivf_words, const uint32_t* __restrict__ row_cluster) {
    using C = Cfg2;
    constexpr int NQ = C::NQ;
    extern __shared__ __align__(1024) uint8_t base[];
    uint8_t* stage0 = base;
    uint64_t* lists = scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 * NQ * LIST;
    uint32_t* thr_u = (uint32_t*)(base + C::STAGES * C::STAGE_BYTES);
    float* qs_sm = (float*)(thr_u + NQ);
    uint64_t* bars = (uint64_t*)(thr_u + 3 * NQ);
    uint64_t* full = bars;                     // [STAGES]  (used in the leader only)
    uint64_t* empty = full + MAX_STAGES;       // [STAGES]
    uint64_t* tfull = empty + MAX_STAGES;      // [2]
    uint64_t* tempty = tfull + 2;              // [2]       (used in the leader only: 24 arrivals = 12 epilogue warps x 2 CTAs)
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair0 = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 24); }
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < NQ; i += THREADS) {
        thr_u[i] = blockIdx.y * NQ + i >= nq_valid ? 0xFFFFFFFFu : (thr_init ? __ldg(&thr_init[blockIdx.y * NQ + i]) : 0u);
        qs_sm[i] = __ldg(&q_margin[blockIdx.y * NQ + i]);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *(volatile uint32_t*)tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): own corpus rows + own half of the query block =====================
        if (lane == 0) {
            tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
            uint32_t it = 0;
            for (uint32_t pr = pair0; pr < n_pairs; pr += n_clusters) {
                for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                    const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                    uint8_t* st = stage0 + s * C::STAGE_BYTES;
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::STAGE_BYTES);       // the bytes of both CTAs land on the leader's barrier
                    tma_load_2d_pair(st, &tmA, (int)(kc * 64), (int)((pr * 2 + rank) * TM), &full[s]);
                    tma_load_2d_pair(st + C::A_BYTES, &tmB, (int)(kc * 64), (int)(blockIdx.y * NQ + rank * C::NQH), &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread of the LEADER CTA, M = 256 (128 rows per CTA) x N = 256 =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(NQ >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
            uint32_t it = 0, ti = 0;
            for (uint32_t pr = pair0; pr < n_pairs; pr += n_clusters, ++ti) {
                const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
                mbar_wait(&tempty[buf], tph ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * NQ;
                for (uint32_t kc = 0; kc < n_kchunks; ++kc, ++it) {
                    const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage0 + s * C::STAGE_BYTES);
                    const uint64_t a = umma_desc_k128(sa), b = umma_desc_k128(sa + C::A_BYTES);
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; kk++)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(d), "l"(a + (uint64_t)(kk * 2)), "l"(b + (uint64_t)(kk * 2)), "r"(idesc), "r"((uint32_t)((kc | kk) != 0)) : "memory");
                    umma_commit_pair(&empty[s]);
                    if (kc + 1 == n_kchunks) umma_commit_pair(&tfull[buf]);
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, 12 warps each): the f32 epilogue of scan_tc for MT = 1, filter margins =====================
        constexpr int EG = 3;
        constexpr int NCH3 = (NQ / CHUNK + EG - 1) / EG * EG;
        const int ew = warp & 3, g = (warp - 4) >> 2;
        uint64_t* mylists = lists + (size_t)ew * NQ * LIST;
        for (int c = g; c < NQ / CHUNK; c += EG)
            for (int i = lane; i < CHUNK * LIST; i += 32) mylists[c * CHUNK * LIST + i] = 0;
        __syncwarp();
        uint32_t ti = 0;
        for (uint32_t pr = pair0; pr < n_pairs; pr += n_clusters, ++ti) {
            const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
            mbar_wait(&tfull[buf], tph);
            tc_fence_after();
            const uint32_t row = (pr * 2 + rank) * TM + (uint32_t)(ew * 32 + lane);
            const bool valid = row < n_rows;
            for (int c = g; c < NCH3; c += EG) {
                if (c >= NQ / CHUNK) continue;
                uint32_t v[CHUNK];
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * NQ + c * CHUNK;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                {
                    const uint4* t4 = reinterpret_cast<const uint4*>(thr_u + c * CHUNK);
                    const uint4 t0 = t4[0], t1 = t4[1];
                    const bool any = (ord_f32(__uint_as_float(v[0])) >= t0.x) | (ord_f32(__uint_as_float(v[1])) >= t0.y) |
                                     (ord_f32(__uint_as_float(v[2])) >= t0.z) | (ord_f32(__uint_as_float(v[3])) >= t0.w) |
                                     (ord_f32(__uint_as_float(v[4])) >= t1.x) | (ord_f32(__uint_as_float(v[5])) >= t1.y) |
                                     (ord_f32(__uint_as_float(v[6])) >= t1.z) | (ord_f32(__uint_as_float(v[7])) >= t1.w);
                    if (!__any_sync(FULL, any && valid)) continue;
                }
#pragma unroll
                for (int j = 0; j < CHUNK; j++) {
                    const int q = c * CHUNK + j;
                    const float sc = __uint_as_float(v[j]);
                    const uint32_t so = ord_f32(sc);
                    const bool pass = valid && sc == sc && so >= thr_u[q];
                    unsigned pm = __ballot_sync(FULL, pass);
                    if (pm) {
                        uint64_t key = 0;
                        if (pass) {
                            const uint32_t doc = doc_ids ? __ldg(&doc_ids[row]) : row;
                            key = ((uint64_t)so << 32) | (uint64_t)(0xFFFFFFFFu - row);          // filter scan: candidates are named by row
                            if (doc_deleted(del_slot, del_words, doc) || ivf_skipped(ivf_sel, ivf_words, blockIdx.y * NQ + q, row_cluster, row)) key = 0;
                        }
                        if (del_slot || ivf_sel) { pm = __ballot_sync(FULL, key != 0); if (!pm) continue; }
                        uint64_t L = mylists[q * LIST + lane];
                        if (__popc(pm) > 3) L = wl_merge(L, wl_sort_desc(key, lane), lane);
                        else while (pm) { const int src = __ffs(pm) - 1; pm &= pm - 1; wl_insert(L, shfl64(key, src), lane); }
                        mylists[q * LIST + lane] = L;
                        uint32_t kth = (uint32_t)(shfl64(L, (int)k - 1) >> 32);
                        if (kth) kth = ord_f32(__fsub_rd(unord_f32(kth), qs_sm[q]));
                        if (lane == 0 && kth > thr_u[q]) atomicMax(&thr_u[q], kth);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the peer may still be reading this CTA's shared memory / arriving on its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// [nq_pad][dpad] f32 -> hi / lo parts: tf32 (f32 containers) or bf16
__global__ void split_queries_tf32(const float* __restrict__ q, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = __float_as_uint(q[i]), h = x & 0xFFFFE000u;
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(x) - __uint_as_float(h);
}
// queries [nq][dims] f32 -> padded, (Cosine:) L2-normalised exactly like prep_queries (vec_scan.cu), split into bf16 hi / lo parts:
// one launch instead of prep_queries + split (the per-batch launch chain is what limits small shards, SCALE_r01)
__global__ void prep_split_queries_bf16(const float* __restrict__ q, uint32_t nq, uint32_t dims, uint64_t qstride,
                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, uint32_t nq_pad, uint32_t dpad, int normalize,
                                        float* __restrict__ f32_out /*filter scan: padded f32 queries for the refine step, else null*/,
                                        float* __restrict__ margin_out /*filter scan: [nq_pad] 2 eps_q*/, const uint32_t* __restrict__ row_err /*{max|a-h(a)|, max|h(a)|} bits*/) {
    const int lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nq_pad) return;
    const bool filter = margin_out != nullptr;             // filter scan: `hi` receives the scaled fp16 query, `lo` is not written
    __nv_bfloat16* oh = hi + (size_t)row * dpad;
    __nv_bfloat16* ol = lo + (size_t)row * dpad;
    __half* o16 = reinterpret_cast<__half*>(oh);
    float* of = f32_out ? f32_out + (size_t)row * dpad : nullptr;
    if (row >= nq) {
        for (uint32_t i = lane; i < dpad; i += 32) { oh[i] = __float2bfloat16_rn(0.f); if (!filter) ol[i] = __float2bfloat16_rn(0.f); if (of) of[i] = 0.f; }
        if (filter && lane == 0) margin_out[row] = 0.f;
        return;
    }
    const float* src = q + (size_t)row * qstride;
    float f = 1.f;
    if (normalize) {
        float s = 0.f;
        for (uint32_t i = lane; i < dims; i += 32) { float v = src[i]; s = fmaf(v, v, s); }
        for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(FULL, s, m);
        f = 1.0f / sqrtf(s);
    }
    if (!filter) {
        for (uint32_t i = lane; i < dpad; i += 32) {
            const float x = i < dims ? src[i] * f : 0.f;
            const __nv_bfloat16 h = __float2bfloat16_rn(x);
            oh[i] = h;
            ol[i] = __float2bfloat16_rn(x - __bfloat162float(h));
            if (of) of[i] = x;
        }
        return;
    }
    // filter scan: scale the query by a power of two so that its largest element lands in [128, 256) (exact; keeps the small elements
    // out of fp16's subnormal range), round to fp16, and bound the error of s^ for this query
    float mx = 0.f;
    for (uint32_t i = lane; i < dims; i += 32) mx = fmaxf(mx, fabsf(src[i] * f));
    for (int m = 16; m; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, m));
    const float sb = (mx > 0.f && mx < 3.0e38f) ? exp2f((float)(7 - ilogbf(mx))) : 1.f;
    float nb2 = 0.f, eb2 = 0.f;   // |b|^2 and |b - h(b)|^2 of the scaled query
    for (uint32_t i = lane; i < dpad; i += 32) {
        const float x = i < dims ? src[i] * f : 0.f;
        const float xs = x * sb;
        const __half h = __float2half_rn(xs);
        const float r = xs - __half2float(h);
        o16[i] = h;
        if (of) of[i] = x;
        nb2 = fmaf(xs, xs, nb2); eb2 = fmaf(r, r, eb2);
    }
    // eps_q >= |s - s^| for every row: |a.b - h(a).h(b)| = |(a - h(a)).b + h(a).(b - h(b))| <= E_a |b| + H_a |b - h(b)| (Cauchy-Schwarz, E_a / H_a =
    // the row maxima computed at load time), plus the f32 accumulation slack of the tensor core (truncating adds: <= 2^-23 of the absolute
    // sum per element) and of the refine dot product (<= 2^-24): dpad * 2^-22 * H_a |b| covers both with room.  1.001 absorbs the rounding
    // of this arithmetic itself.
    for (int m = 16; m; m >>= 1) { nb2 += __shfl_xor_sync(FULL, nb2, m); eb2 += __shfl_xor_sync(FULL, eb2, m); }
    if (lane == 0) {
        const float Ea = __uint_as_float(row_err[0]), Ha = __uint_as_float(row_err[1]);
        const float nb = sqrtf(nb2) * 1.00001f, eb = sqrtf(eb2) * 1.00001f;
        const float eps = (Ea * nb + Ha * eb + (float)dpad * 2.38418579e-7f * Ha * nb) * 1.001f;
        margin_out[row] = 2.0f * eps;
    }
}
// load time (filter scan): the fp16 plane h = half_rn(x * scale) (scale = a power of two chosen per index) and the index-wide error bounds
// E_a = max over rows of |x*scale - h|_2, H_a = max over rows of |h|_2, rounded up; one warp per row, maxima kept as f32 bit patterns
// (non-negative floats order like their bits)
__global__ void rows_f16_err(const float* __restrict__ x, __half* __restrict__ h16, uint64_t n, uint32_t dpad, float scale, uint32_t* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    float e2 = 0.f, h2 = 0.f;
    for (uint32_t i = lane; i < dpad; i += 32) {
        const float xs = x[row * dpad + i] * scale;
        const __half hh = __float2half_rn(xs);
        const float h = __half2float(hh), r = xs - h;
        h16[row * dpad + i] = hh;
        e2 = fmaf(r, r, e2); h2 = fmaf(h, h, h2);
    }
    for (int m = 16; m; m >>= 1) { e2 += __shfl_xor_sync(FULL, e2, m); h2 += __shfl_xor_sync(FULL, h2, m); }
    if (lane == 0) {
        const float e = sqrtf(e2) * 1.00001f, h = sqrtf(h2) * 1.00001f;
        if (!(e < 3.0e38f) || !(h < 3.0e38f)) { atomicMax(&err[0], 0x7F800000u); atomicMax(&err[1], 0x7F800000u); return; }   // NaN / overflowing row: infinite margin -> every query falls back
        atomicMax(&err[0], __float_as_uint(e)); atomicMax(&err[1], __float_as_uint(h));
    }
}
// max |x| over a block of rows (chooses the fp16 scale of a Dot index from its first level); err[0] as f32 bits
__global__ void max_abs_f32(const float* __restrict__ x, size_t n, uint32_t* __restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const float v = fabsf(x[i]); if (v < 3.0e38f) m = fmaxf(m, v); }
    for (int s = 16; s; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// load time: (normalised) f32 corpus rows -> the two bf16 planes the tensor-core scan streams (hi = bf16_rn(x), lo = bf16_rn(x - hi))
__global__ void split_rows_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// thr[q] = ordered-uint of the k-th largest of gmax[q][0..n_rg) as an f32 score (0 = no threshold when fewer than k groups
// hold an eligible row).  One warp per query, lane-distributed sorted list, chunks that cannot enter the list are skipped.
__global__ void __launch_bounds__(256)
kth_from_groupmax(const int* __restrict__ gmax, uint32_t n_rg, uint32_t nq, uint32_t k, uint32_t* __restrict__ thr,
                  int is_int /*1: int32 dot products (INT_MIN = none), 0: ordered-uint f32 scores (0 = none)*/,
                  const float* __restrict__ margin /*filter scan: thr = k-th - margin[q]; else null*/) {
    // one CTA per query: 8 warps each reduce a slice of the group maxima to a sorted top-32, warp 0 merges the 8 lists (one warp per
    // query walked n_rg / 32 dependent iterations: 57 us under ncu for 2368 groups x 256 queries, as long as the sample scan itself)
    __shared__ uint64_t sm[8][LIST];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x;
    const int* g = gmax + (size_t)q * n_rg;
    uint64_t L = 0;
    for (uint32_t base = warp * 32; base < n_rg; base += 256) {
        const uint32_t i = base + lane;
        // key: order-preserving unsigned value in the high word (> 0 for every eligible row), group index below it keeps keys distinct
        const uint32_t u = i < n_rg ? ((uint32_t)g[i] ^ (is_int ? 0x80000000u : 0u)) : 0u;
        const uint64_t key = u == 0u ? 0ull : ((uint64_t)u << 32) | (uint64_t)(0xFFFFFFFFu - i);
        const uint64_t kth = shfl64(L, (int)k - 1);
        if (__any_sync(FULL, key > kth)) L = wl_merge(L, wl_sort_desc(key, lane), lane);
    }
    sm[warp][lane] = L;
    __syncthreads();
    if (warp != 0) return;
    for (int w = 1; w < 8; w++) L = wl_merge(L, sm[w][lane], lane);
    const uint64_t kth = shfl64(L, (int)k - 1);
    if (lane == 0) {
        uint32_t t = !kth ? 0u : (is_int ? ord_f32((float)(int)((uint32_t)(kth >> 32) ^ 0x80000000u)) : (uint32_t)(kth >> 32));
        if (t && margin) t = ord_f32(__fsub_rd(unord_f32(t), margin[q]));
        thr[q] = t;
    }
}

}  // namespace tc

template <int NQ, int PREC, bool BRES = false, int SCALED = 0>
static int32_t launch_tc_n(const ScanArgs& a, cudaStream_t st) {
    using C = tc::Cfg<NQ, PREC, BRES>;
    CUtensorMap tmA, tmA2, tmBh, tmBl;
    uint32_t n_tiles = (uint32_t)((a.n_rows + C::TROWS - 1) / C::TROWS);
    uint32_t n_groups = a.nq_pad / NQ;
    size_t nel = (size_t)a.nq_pad * a.dpad;
    uint32_t n_kchunks = a.dpad / tc::KC;
    if constexpr (PREC == tc::PREC_I8) {
        // int8 corpus / queries (quantised by the caller): 128 dims per 128-byte swizzle row
        SSB_TRY(encode_tmap_2d(&tmA, a.rows_i8, 1, a.dpad8, a.n_rows, a.dpad8, 128, C::TROWS, 128));
        SSB_TRY(encode_tmap_2d(&tmBh, a.queries_i8, 1, a.dpad8, a.nq_pad, a.dpad8, 128, NQ, 128));
        tmBl = tmBh; tmA2 = tmA;
        n_kchunks = a.dpad8 / 128;
    } else if constexpr (PREC == tc::PREC_TF32) {
        SSB_TRY(encode_tmap_2d_f32(&tmA, a.rows, a.dpad, a.n_rows, (uint64_t)a.dpad * 4, tc::KC, C::TROWS, 1));
        SSB_TRY(encode_tmap_2d_f32(&tmBh, a.q_hi, a.dpad, a.nq_pad, (uint64_t)a.dpad * 4, tc::KC, NQ, 1));
        SSB_TRY(encode_tmap_2d_f32(&tmBl, a.q_lo, a.dpad, a.nq_pad, (uint64_t)a.dpad * 4, tc::KC, NQ, 1));
        tmA2 = tmA;
        if (!a.thr_init) tc::split_queries_tf32<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(a.queries_padded, a.q_hi, a.q_lo, nel);
    } else if constexpr (PREC == tc::PREC_F16F) {
        // filter scan: fp16 plane and fp16 queries only, 64 halves (128 bytes) per swizzle row; the box may run past dpad (zero fill)
        if (!a.rows_h16 || !a.q_scale) { set_error("tcgen05 filter scan: the index holds no fp16 plane / no margins"); return SSB_E_STATE; }
        SSB_TRY(encode_tmap_2d(&tmA, a.rows_h16, 2 /*fp16: 2-byte elements*/, a.dpad, a.n_rows, (uint64_t)a.dpad * 2, 64, C::TROWS, 128));
        SSB_TRY(encode_tmap_2d(&tmBh, a.q_hi, 2, a.dpad, a.nq_pad, (uint64_t)a.dpad * 2, 64, NQ, 128));
        tmBl = tmBh; tmA2 = tmA;
        n_kchunks = (a.dpad + 63) / 64;
    } else {
        // corpus: the two bf16 planes written at load time; queries: the two bf16 parts live in the q_hi / q_lo buffers (half of
        // each is used) and were written by prep_split_queries_bf16 (launch_prep_split_queries_bf16)
        if (!a.rows_hi || !a.rows_lo) { set_error("tcgen05 bf16 scan: the index holds no bf16 planes"); return SSB_E_STATE; }
        SSB_TRY(encode_tmap_2d(&tmA, a.rows_hi, 2 /*bf16*/, a.dpad, a.n_rows, (uint64_t)a.dpad * 2, tc::KC, C::TROWS, 64));
        SSB_TRY(encode_tmap_2d(&tmA2, a.rows_lo, 2 /*bf16*/, a.dpad, a.n_rows, (uint64_t)a.dpad * 2, tc::KC, C::TROWS, 64));
        SSB_TRY(encode_tmap_2d(&tmBh, a.q_hi, 2 /*bf16*/, a.dpad, a.nq_pad, (uint64_t)a.dpad * 2, tc::KC, NQ, 64));
        SSB_TRY(encode_tmap_2d(&tmBl, a.q_lo, 2 /*bf16*/, a.dpad, a.nq_pad, (uint64_t)a.dpad * 2, tc::KC, NQ, 64));
    }
    uint32_t gx = n_tiles < (uint32_t)a.n_sms ? n_tiles : (uint32_t)a.n_sms;
    if ((size_t)n_groups * gx * 4 * NQ * LIST * 8 > a.scratch_bytes) { set_error("vector scan scratch too small"); return SSB_E_STATE; }
    constexpr int SMEM_MAX = 232448;   // 227 KB opt-in limit per CTA
    uint32_t nst = C::STAGES;
    int smem = C::SMEM;
    if (BRES) {   // resident query block + as many corpus stages as fit (launch_scan_tc_impl guarantees >= 3)
        const int fixed = (int)n_kchunks * C::B_BYTES + NQ * 12 + 256;
        nst = (uint32_t)((SMEM_MAX - fixed) / C::STAGE_BYTES);
        if (nst > (uint32_t)tc::MAX_STAGES) nst = tc::MAX_STAGES;
        smem = fixed + (int)nst * C::STAGE_BYTES;
    }
    // per launch, not once per process: the opt-in applies to the CURRENT device's context only (ssb_config.device allows
    // several indexes on different GPUs in one process); the call is a cheap host-side attribute write
    SSB_CUDA_TRY(cudaFuncSetAttribute(tc::scan_tc<NQ, PREC, BRES, SCALED>, cudaFuncAttributeMaxDynamicSharedMemorySize, BRES ? SMEM_MAX : C::SMEM));
    if (a.ev0) cudaEventRecord(a.ev0, st);
    tc::scan_tc<NQ, PREC, BRES, SCALED><<<dim3(gx, n_groups), tc::THREADS, smem, st>>>(tmA, tmA2, tmBh, tmBl, (uint32_t)a.n_rows, n_kchunks, n_tiles,
                                                                             a.k, a.doc_ids, a.scratch, a.thr_init, a.nq_valid ? a.nq_valid : a.nq_pad, a.ceil_keys, nst,
                                                                             a.del_slot, a.del_words, a.row_scale, a.row_norm, a.q_scale, a.q_norm,
                                                                             (const int2*)a.row_aff, (const int2*)a.q_aff,
                                                                             PREC == tc::PREC_I8 ? nullptr : a.ivf_sel, a.ivf_words, a.row_cluster,
                                                                             a.sample_groupmax ? 1u : 0u);
    if (a.ev1) cudaEventRecord(a.ev1, st);
    SSB_CUDA_TRY(cudaGetLastError());
    if (a.sample_groupmax) {   // threshold seeding pass: scratch holds gmax[nq_pad][n_tiles * 8]
        tc::kth_from_groupmax<<<a.nq_pad, 256, 0, st>>>((const int*)a.scratch, n_tiles * (C::MT * 4), a.nq_pad, a.k, a.thr_buf,
                                                                      (PREC == tc::PREC_I8 && !SCALED) ? 1 : 0, PREC == tc::PREC_F16F ? a.q_scale : nullptr);
        SSB_CUDA_TRY(cudaGetLastError());
        if (a.launches) *a.launches += PREC == tc::PREC_TF32 ? 3 : 2;   // (tf32 query split +) scan + kth
        return SSB_OK;
    }
    // scratch layout [group][list][q in NQ][32] -> generic merge with qt = NQ
    merge_lists_generic(a.scratch, gx * 4, NQ, a.nq_pad, a.keys_out, st);
    SSB_CUDA_TRY(cudaGetLastError());
    if (a.launches) *a.launches += 2;   // scan + merge (the tf32 query split is counted with the sample pass)
    return SSB_OK;
}

// 256-query filter scan on CTA pairs (scan_tc2): thread-block clusters of 2, one pair per 256 corpus rows and k-chunk
static int32_t launch_tc2(const ScanArgs& a, cudaStream_t st) {
    using C = tc::Cfg2;
    if (!a.rows_h16 || !a.q_scale) { set_error("tcgen05 filter scan: the index holds no fp16 plane / no margins"); return SSB_E_STATE; }
    if (a.nq_pad % C::NQ != 0) { set_error("tcgen05 pair scan: query count must be padded to 256"); return SSB_E_INVALID; }
    CUtensorMap tmA, tmB;
    SSB_TRY(encode_tmap_2d(&tmA, a.rows_h16, 2, a.dpad, a.n_rows, (uint64_t)a.dpad * 2, 64, tc::TM, 128));
    SSB_TRY(encode_tmap_2d(&tmB, a.q_hi, 2, a.dpad, a.nq_pad, (uint64_t)a.dpad * 2, 64, C::NQH, 128));
    const uint32_t n_kchunks = (a.dpad + 63) / 64, n_groups = a.nq_pad / C::NQ;
    const uint32_t n_pairs = (uint32_t)((a.n_rows + 2 * tc::TM - 1) / (2 * tc::TM));
    uint32_t n_clusters = (uint32_t)a.n_sms / 2;
    if (n_clusters > n_pairs) n_clusters = n_pairs;
    const uint32_t gx = 2 * n_clusters;
    if ((size_t)n_groups * gx * 4 * C::NQ * LIST * 8 > a.scratch_bytes) { set_error("vector scan scratch too small"); return SSB_E_STATE; }
    SSB_CUDA_TRY(cudaFuncSetAttribute(tc::scan_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(gx, n_groups); cfg.blockDim = dim3(tc::THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (a.ev0) cudaEventRecord(a.ev0, st);
    SSB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::scan_tc2, tmA, tmB, (uint32_t)a.n_rows, n_kchunks, n_pairs, a.k, a.doc_ids, a.scratch, a.thr_init,
                                    a.nq_valid ? a.nq_valid : a.nq_pad, a.del_slot, a.del_words, a.q_scale, a.ivf_sel, a.ivf_words, a.row_cluster));
    if (a.ev1) cudaEventRecord(a.ev1, st);
    merge_lists_generic(a.scratch, gx * 4, C::NQ, a.nq_pad, a.keys_out, st);
    SSB_CUDA_TRY(cudaGetLastError());
    if (a.launches) *a.launches += 2;
    return SSB_OK;
}

static int32_t launch_scan_tc_impl(const ScanArgs& a, uint32_t nq_tile, int prec /*0 tf32, 1 bf16, 2 int8, 3 fp16 filter, 4 fp16 filter on CTA pairs*/, cudaStream_t st) {
    if (a.n_rows == 0 || a.nq_pad == 0) return SSB_OK;
    if (prec == 4) return a.sample_groupmax ? launch_tc_n<256, tc::PREC_F16F>(a, st) : launch_tc2(a, st);   // the seeding pass stays on one CTA per SM
    if (prec == 2) {
        if (nq_tile != 128 || a.nq_pad % 128 != 0 || !a.rows_i8 || !a.queries_i8 || a.dpad8 % 128) { set_error("int8 scan: bad arguments"); return SSB_E_INVALID; }
        // query block resident in smem when it leaves room for >= 3 corpus stages (dims <= 1024), else streamed per stage
        if (a.i8_scaled) {
            if (!a.row_scale || !a.q_scale || (a.i8_scaled >= 2 && (!a.row_norm || !a.q_norm)) || (a.i8_scaled == 3 && (!a.row_aff || !a.q_aff))) { set_error("scaled int8 scan: missing scale / norm arrays"); return SSB_E_INVALID; }
            if (a.i8_scaled == 3) return a.dpad8 <= 1024 ? launch_tc_n<128, tc::PREC_I8, true, 3>(a, st) : launch_tc_n<128, tc::PREC_I8, false, 3>(a, st);
            if (a.i8_scaled == 1) return a.dpad8 <= 1024 ? launch_tc_n<128, tc::PREC_I8, true, 1>(a, st) : launch_tc_n<128, tc::PREC_I8, false, 1>(a, st);
            return a.dpad8 <= 1024 ? launch_tc_n<128, tc::PREC_I8, true, 2>(a, st) : launch_tc_n<128, tc::PREC_I8, false, 2>(a, st);
        }
        return a.dpad8 <= 1024 ? launch_tc_n<128, tc::PREC_I8, true>(a, st) : launch_tc_n<128, tc::PREC_I8, false>(a, st);
    }
    if (a.similarity == SSB_SIM_EUCLIDEAN) { set_error("tcgen05 scan supports Dot/Cosine only"); return SSB_E_UNSUPPORTED; }
    if (prec == 3) {
        if ((nq_tile != 128 && nq_tile != 256) || a.nq_pad % nq_tile != 0) { set_error("tcgen05 filter scan: query count must be padded to the 128/256 query tile"); return SSB_E_INVALID; }
        return nq_tile == 256 ? launch_tc_n<256, tc::PREC_F16F>(a, st) : launch_tc_n<128, tc::PREC_F16F>(a, st);
    }
    if ((nq_tile != 64 && nq_tile != 128 && !(nq_tile == 256 && prec == 1)) || a.nq_pad % nq_tile != 0) { set_error("tcgen05 scan: query count must be padded to the 64/128(/256 bf16) query tile"); return SSB_E_INVALID; }
    if (prec == 1) return nq_tile == 64 ? launch_tc_n<64, tc::PREC_BF16>(a, st) : (nq_tile == 256 ? launch_tc_n<256, tc::PREC_BF16>(a, st) : launch_tc_n<128, tc::PREC_BF16>(a, st));
    return nq_tile == 64 ? launch_tc_n<64, tc::PREC_TF32>(a, st) : launch_tc_n<128, tc::PREC_TF32>(a, st);
}

int32_t launch_scan_tc(const ScanArgs& a, uint32_t nq_tile, int prec, cudaStream_t st) {
    // threshold pre-sampling (see vec_scan.cu): scan the first rows, seed the thresholds, then the full scan
    // (with a delete set the sample pass is skipped: a deleted row must never seed a threshold)
    if (a.thr_init || !a.thr_buf || a.del_slot || a.ivf_sel || vec_presample_rows(a.n_rows, true) == 0) return launch_scan_tc_impl(a, nq_tile, prec, st);
    ScanArgs pre = a;
    // The sample pass writes per-(32-row group, query) score maxima instead of lists (no insert storm) and costs the same for one
    // 256-row tile per CTA as for a handful of tiles: sample one tile per SM.  (An earlier version ran the normal list epilogue
    // over N/128 rows: ~90 us per pass, and ncu showed the full scan's epilogue warps waiting on list loads for candidates that
    // a better seed rejects.)
    const uint64_t trows = nq_tile == 256 ? 128 : tc::TROWS;      // rows per stage of the variant that will run
    // sample tiles per SM: the seed is the k-th best of S sampled rows, the full scan then sees ~k*N/S candidates per query, each a
    // ~1 us latency-bound list insert for an epilogue warp.  The filter scan streams a pass in half the time of the 3-product scan, so the
    // same insert load weighs twice as much: it samples more (SSB_TC_SAMPLE_TILES overrides; measured in DESIGN.md §3.2c)
    static const int env_tiles = [] { const char* e = getenv("SSB_TC_SAMPLE_TILES"); return e ? atoi(e) : 0; }();
    const int sample_tiles = env_tiles > 0 ? (env_tiles > 16 ? 16 : env_tiles) : ((prec >= 3 && nq_tile == 256) ? 2 : 1);
    uint64_t s = (uint64_t)a.n_sms * trows * sample_tiles;
    if (s > a.n_rows / 4) s = a.n_rows / 4 / trows * trows;
    pre.n_rows = s; pre.ev0 = nullptr; pre.ev1 = nullptr;
    pre.sample_groupmax = true;                          // writes the thresholds straight into thr_buf
    SSB_TRY(launch_scan_tc_impl(pre, nq_tile, prec, st));
    ScanArgs full = a;
    full.thr_init = a.thr_buf;
    return launch_scan_tc_impl(full, nq_tile, prec, st);
}

void launch_kth_from_groupmax(const void* gmax, uint32_t n_groups, uint32_t nq, uint32_t k, uint32_t* thr, int is_int, cudaStream_t st) {
    if (nq) tc::kth_from_groupmax<<<nq, 256, 0, st>>>((const int*)gmax, n_groups, nq, k, thr, is_int, nullptr);
}

int32_t launch_prep_split_queries_bf16(const float* q, uint32_t nq, uint32_t dims, uint64_t qstride, void* hi, void* lo, uint32_t nq_pad,
                                       uint32_t dpad, int normalize, cudaStream_t st, float* f32_out, float* margin_out, const uint32_t* row_err) {
    if (nq_pad == 0) return SSB_OK;
    tc::prep_split_queries_bf16<<<(nq_pad + 7) / 8, 256, 0, st>>>(q, nq, dims, qstride, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, nq_pad, dpad, normalize,
                                                                  f32_out, margin_out, row_err);
    SSB_CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

int32_t launch_split_rows_bf16(const float* rows, void* hi, void* lo, size_t n_elems, cudaStream_t st) {
    if (n_elems == 0) return SSB_OK;
    tc::split_rows_bf16<<<(unsigned)((n_elems + 255) / 256), 256, 0, st>>>(rows, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n_elems);
    SSB_CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

int32_t launch_rows_f16_err(const float* rows, void* h16, uint64_t n, uint32_t dpad, float scale, uint32_t* err, cudaStream_t st) {
    if (n == 0) return SSB_OK;
    tc::rows_f16_err<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(rows, (__half*)h16, n, dpad, scale, err);
    SSB_CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}
int32_t launch_max_abs_f32(const float* x, size_t n, uint32_t* out_bits, cudaStream_t st) {
    if (n == 0) return SSB_OK;
    tc::max_abs_f32<<<296, 256, 0, st>>>(x, n, out_bits);
    SSB_CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

size_t scan_tc_scratch_bytes(int n_sms, uint32_t nq_pad) { return (size_t)nq_pad * (size_t)n_sms * 4 * LIST * 8; }

}  // namespace vec
}  // namespace ssb
