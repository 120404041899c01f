// Residual projection fused with the Norm that follows it (Model/layers.py:26-29, 62-72: `x = x + dropout(sublayer(...))` and
// then `x2 = self.norm_k(x)`; Norm = Model/modules.py:80-95), for d_model = 512:
//     x    = dropout(A W^T + bias) + res          [M, 512] fp32   (the residual stream; saved for the Norm backward)
//     y    = alpha (x - mean) / (std_unbiased + eps) + beta        -> bf16 (operand of the next GEMM) [+ fp32 copy]
// One CTA owns whole rows: the accumulator of a 128 x 512 tile fills all 512 TMEM columns (single-buffered), so the row
// statistics never leave the SM: the epilogue makes three passes over TMEM (build x and write it back in place + row sum;
// centred sum of squares; normalise), exactly the two-pass mean / unbiased-std arithmetic of norm_fwd_kernel.  This removes the
// separate norm_fwd launch and its fp32 re-read of x, and a 128 x 512 tile has 1.6x the arithmetic intensity per operand byte
// of the 128 x 128 tiles these short-K projections otherwise use.
// [B200] measured (scripts/one_rownorm.py, profiles/r02_rownorm_fusion_ab.txt): correct, but NOT faster than the residual GEMM +
// norm_fwd pair.  With the residual read per lane (RES_TMA = false): M = 41472, K = 512: 96.8 us; with the residual fetched as
// TMA boxes into the staging tiles (RES_TMA = true, the default): 79 us -- against 84.6 us for the pair at that time and 71 us
// (47.4 + 23.5) once the same boxed fetch went into the persistent GEMM's own epilogue.  With all 512 TMEM columns holding one
// accumulator the epilogue (three TMEM passes, 96 staged TMA stores per tile) cannot overlap the next tile's main loop, while the
// unfused kernel double-buffers 128-column accumulators; writing x / y straight from registers instead of through the staging
// tiles was slower still (134 us).  Kept as a tested option (gct_set_rownorm_fusion), off by default.
// Warp roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 epilogue (TMEM lane quarter = warp % 4,
// column slice = (warp - 2) / 4 of 128 columns).  2-stage ring of (A 128 x 64, B 512 x 64) bf16 tiles, SWIZZLE_128B.
#pragma once
#include "gemm_tc.cuh"

extern int g_gct_rownorm_res_tma;

namespace tc {

struct RowNormParams {
    const float* bias;        // [512] or null
    const float* res32;       // [M, 512] fp32 or null
    float* out32;             // [M, 512] fp32 x (may alias res32)
    const float* alpha;       // Norm parameters [512]
    const float* beta;
    float* norm32;            // optional [M, 512] fp32 Norm(x)
    DropCtx drop;
    float eps;
    int M, K;
};

constexpr int RN_N = 512, RN_STAGES = 2, RN_EW = 4, RN_THREADS = (2 + 4 * RN_EW) * 32;
struct RnSmem {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = RN_N * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_OFF = RN_STAGES * STAGE_BYTES;              // 16 epilogue warps x 4 KB
    static constexpr int RED_OFF = STAGING_OFF + 4 * RN_EW * 4096;           // [4 slices][128 rows] floats
    static constexpr int BAR_OFF = RED_OFF + RN_EW * 128 * 4;
    static constexpr int TOTAL = BAR_OFF + 64 + 16 * 8;                      // pipeline barriers + one residual barrier per epilogue warp
    static constexpr int REQUEST = 232448;                                   // the whole 227 KB: up to 960 B of alignment slack
};
static_assert(RnSmem::TOTAL <= RnSmem::REQUEST, "row-norm GEMM does not fit in shared memory");

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// named barrier over the 16 epilogue warps
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(4 * RN_EW * 32) : "memory"); }

// RES_TMA: the residual arrives as 32-row x 32-column boxes in the warp's staging tile (coalesced, no per-row loads) instead of
// through per-lane loads of the lane's own row.
template <bool RES_TMA>
__global__ void __launch_bounds__(RN_THREADS, 1)
gemm_rownorm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                    const __grid_constant__ CUtensorMap tmR, RowNormParams p) {
    using L = RnSmem;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (base - smem_u32(smem_raw) + L::TOTAL > L::REQUEST) __trap();          // (never: dynamic shared memory starts 1 KB aligned)
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + L::BAR_OFF;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (RN_STAGES + s); };
    const uint32_t tfull_bar = bars + 8u * (2 * RN_STAGES), tempty_bar = tfull_bar + 8u;
    const uint32_t tmem_slot = tempty_bar + 8u;
    uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(sm + L::BAR_OFF + 8 * (2 * RN_STAGES + 2));
    float* red = reinterpret_cast<float*>(sm + L::RED_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;
    const int m_tiles = (p.M + BM - 1) / BM;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < RN_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int w = 0; w < 4 * RN_EW; ++w) mbar_init(bars + 64u + 8u * w, 1);
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4 * RN_EW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
                const int m0 = tile * BM;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(s), phase ^ 1u);
                    mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
                    const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
                    tma_load_2d(sa, &tmA, kb * BK, m0, full_bar(s));
                    tma_load_2d(sb, &tmB, kb * BK, 0, full_bar(s));              // W rows [0, 256)
                    tma_load_2d(sb + 32768, &tmB, kb * BK, 256, full_bar(s));    // W rows [256, 512)
                    if (++s == RN_STAGES) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, 256, false, false);
            int s = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar, acc_phase ^ 1u);            // the epilogue has drained the accumulator
                tcgen05_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(s), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = make_smem_desc(sa + k * 32, 16, 1024);
                        umma_bf16(tmem_base, ad, make_smem_desc(sb + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_bf16(tmem_base + 256, ad, make_smem_desc(sb + 32768 + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));
                    if (++s == RN_STAGES) { s = 0; phase ^= 1u; }
                }
                umma_commit(tfull_bar);
                acc_phase ^= 1u;
            }
        }
    } else {
        const int q = warp & 3, cs = (warp - 2) >> 2;
        const int c0 = cs * 128;
        uint8_t* stg = sm + L::STAGING_OFF + (warp - 2) * 4096;
        const uint32_t stg_s = smem_u32(stg);
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        uint32_t acc_phase = 0, res_phase = 0;
        for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
            const int m0 = tile * BM;
            const int row = m0 + q * 32 + lane;
            const bool rok = row < p.M;
            const size_t roff = (size_t)row * RN_N + c0;
            // ---- pass 1: x = dropout(acc + bias) + res, written back to TMEM in place and out through the staging tile
            float sum = 0.f;
            if constexpr (RES_TMA) {
                const uint32_t rbar = bars + 64u + 8u * (uint32_t)(warp - 2);
                const bool has_res = p.res32 != nullptr;
                if (has_res && lane == 0) {                 // first residual box: in flight before the accumulator is ready
                    mbar_expect_tx(rbar, 4096);
                    tma_load_2d(stg_s, &tmR, c0, m0 + q * 32, rbar);
                }
                mbar_wait(tfull_bar, acc_phase);
                tcgen05_fence_after();
                epi_bar_sync();                    // every warp is past the previous tile's reads of `red`
#pragma unroll 1
                for (int sg = 0; sg < 4; ++sg) {
                    float v[2][16];
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int ch = sg * 2 + c2;
                        tmem_ld16(trow + ch * 16, v[c2]);
                        if (p.bias) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + ch * 16) + i);
                                v[c2][4 * i] += b.x; v[c2][4 * i + 1] += b.y; v[c2][4 * i + 2] += b.z; v[c2][4 * i + 3] += b.w;
                            }
                        }
                        if (p.drop.thresh) {
                            const uint32_t pair0 = (uint32_t)((roff + ch * 16) >> 1);
#pragma unroll
                            for (int i = 0; i < 8; ++i) drop_pair(p.drop, pair0 + i, v[c2][2 * i], v[c2][2 * i + 1]);
                        }
                    }
                    if (has_res) mbar_wait(rbar, res_phase);           // the segment's residual box has landed
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float4* slot = reinterpret_cast<float4*>(stg + stage_off(lane, c2 * 4 + i));
                            float4 x4 = make_float4(v[c2][4 * i], v[c2][4 * i + 1], v[c2][4 * i + 2], v[c2][4 * i + 3]);
                            if (has_res) { const float4 r4 = *slot; x4.x += r4.x; x4.y += r4.y; x4.z += r4.z; x4.w += r4.w; }
                            *slot = x4;
                            v[c2][4 * i] = x4.x; v[c2][4 * i + 1] = x4.y; v[c2][4 * i + 2] = x4.z; v[c2][4 * i + 3] = x4.w;
                            sum += (x4.x + x4.y) + (x4.z + x4.w);
                        }
                        tmem_st16(trow + (sg * 2 + c2) * 16, v[c2]);
                    }
                    if (has_res) res_phase ^= 1u;
                    if (p.out32) stage_tma_store(&tmX, stg_s, c0 + sg * 32, m0 + q * 32, lane);      // also frees the tile for the next box
                    else __syncwarp();
                    if (has_res && sg + 1 < 4 && lane == 0) {
                        mbar_expect_tx(rbar, 4096);
                        tma_load_2d(stg_s, &tmR, c0 + (sg + 1) * 32, m0 + q * 32, rbar);
                    }
                }
            } else {
            float4 rnext[4];
            if (p.res32 && rok) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rnext[i] = *(reinterpret_cast<const float4*>(p.res32 + roff) + i);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) rnext[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(tfull_bar, acc_phase);
            tcgen05_fence_after();
            epi_bar_sync();                    // every warp is past the previous tile's reads of `red`
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                float4 rc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) rc[i] = rnext[i];
                if (ch + 1 < 8 && p.res32 && rok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) rnext[i] = *(reinterpret_cast<const float4*>(p.res32 + roff + (ch + 1) * 16) + i);
                }
                float v[16];
                tmem_ld16(trow + ch * 16, v);
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + ch * 16) + i);
                        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                    }
                }
                if (p.drop.thresh) {
                    const uint32_t pair0 = (uint32_t)((roff + ch * 16) >> 1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) drop_pair(p.drop, pair0 + i, v[2 * i], v[2 * i + 1]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[4 * i] += rc[i].x; v[4 * i + 1] += rc[i].y; v[4 * i + 2] += rc[i].z; v[4 * i + 3] += rc[i].w;
                    sum += (v[4 * i] + v[4 * i + 1]) + (v[4 * i + 2] + v[4 * i + 3]);
                }
                tmem_st16(trow + ch * 16, v);
                if (p.out32) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(stg + stage_off(lane, (ch & 1) * 4 + i)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    if (ch & 1) stage_tma_store(&tmX, stg_s, c0 + (ch - 1) * 16, m0 + q * 32, lane);      // 32 fp32 columns x 32 rows
                }
            }
            }
            tmem_st_wait();
            red[cs * 128 + q * 32 + lane] = sum;
            epi_bar_sync();
            const int rr = q * 32 + lane;
            const float mean = ((red[rr] + red[128 + rr]) + (red[256 + rr] + red[384 + rr])) * (1.f / RN_N);
            epi_bar_sync();                    // everyone has read the sums before the slots are reused
            // ---- pass 2: centred sum of squares
            float sq = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                float v[16];
                tmem_ld16(trow + ch * 16, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) { const float dlt = v[i] - mean; sq = fmaf(dlt, dlt, sq); }
            }
            red[cs * 128 + rr] = sq;
            epi_bar_sync();
            const float var = ((red[rr] + red[128 + rr]) + (red[256 + rr] + red[384 + rr])) * (1.f / (RN_N - 1));
            const float rinv = 1.f / (sqrtf(var) + p.eps);
            // ---- pass 3: normalise -> bf16 through the staging tile (64 columns per store) [+ fp32 copy, direct]
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                float v[16];
                tmem_ld16(trow + ch * 16, v);
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(p.alpha + c0 + ch * 16) + i);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + ch * 16) + i);
                    v[4 * i] = a.x * (v[4 * i] - mean) * rinv + b.x;
                    v[4 * i + 1] = a.y * (v[4 * i + 1] - mean) * rinv + b.y;
                    v[4 * i + 2] = a.z * (v[4 * i + 2] - mean) * rinv + b.z;
                    v[4 * i + 3] = a.w * (v[4 * i + 3] - mean) * rinv + b.w;
                    pk[2 * i] = pack_bf16x2(v[4 * i], v[4 * i + 1]);
                    pk[2 * i + 1] = pack_bf16x2(v[4 * i + 2], v[4 * i + 3]);
                }
                if (p.norm32 && rok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *(reinterpret_cast<float4*>(p.norm32 + roff + ch * 16) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                *reinterpret_cast<uint4*>(stg + stage_off(lane, (ch & 3) * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(stg + stage_off(lane, (ch & 3) * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                if ((ch & 3) == 3) stage_tma_store(&tmY, stg_s, c0 + (ch - 3) * 16, m0 + q * 32, lane);   // 64 bf16 columns x 32 rows
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar) : "memory");
            acc_phase ^= 1u;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// A [M, K] bf16 K-major (row pitch lda), W [512, K] bf16 K-major (row pitch ldw); K % 8 == 0.
static int launch_gemm_rownorm(const bf16* A, long long lda, const bf16* W, long long ldw, bf16* normT, const RowNormParams& p,
                               cudaStream_t st) {
    GCT_REQUIRE(p.M >= 1 && p.K >= 8 && (p.K % 8) == 0 && p.alpha && p.beta && normT, "row-norm GEMM: bad arguments");
    CUtensorMap ta, tb, tx, ty;
    GCT_TRY(get_tensor_map(A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, BK, BM, &ta));
    GCT_TRY(get_tensor_map(W, (uint64_t)p.K, (uint64_t)RN_N, (uint64_t)ldw * 2, BK, 256, &tb));
    tx = ta;
    if (p.out32) GCT_TRY(get_tensor_map(p.out32, (uint64_t)RN_N, (uint64_t)p.M, (uint64_t)RN_N * 4, 32, 32, &tx, 4));
    GCT_TRY(get_tensor_map(normT, (uint64_t)RN_N, (uint64_t)p.M, (uint64_t)RN_N * 2, 64, 32, &ty, 2));
    CUtensorMap tr = tx;
    if (p.res32 && p.res32 != p.out32) GCT_TRY(get_tensor_map(p.res32, (uint64_t)RN_N, (uint64_t)p.M, (uint64_t)RN_N * 4, 32, 32, &tr, 4));
    const int m_tiles = (p.M + BM - 1) / BM;
    const int grid = m_tiles < sm_count() ? m_tiles : sm_count();
    // the boxed residual needs the fp32 output's staging geometry (always present in the model's uses)
    if (g_gct_rownorm_res_tma && p.out32) {
        GCT_SMEM_LIMIT(gemm_rownorm_kernel<true>, RnSmem::REQUEST);
        GCT_CUDA(launch_k(gemm_rownorm_kernel<true>, dim3(grid), dim3(RN_THREADS), (size_t)RnSmem::REQUEST, st, true, ta, tb, tx, ty, tr, p));
    } else {
        GCT_SMEM_LIMIT(gemm_rownorm_kernel<false>, RnSmem::REQUEST);
        GCT_CUDA(launch_k(gemm_rownorm_kernel<false>, dim3(grid), dim3(RN_THREADS), (size_t)RnSmem::REQUEST, st, true, ta, tb, tx, ty, tr, p));
    }
    return GCT_OK;
}

}  // namespace tc
