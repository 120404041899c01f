// tcgen05 GEMM for sm_100a:  C[M,N] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulation in TMEM.
//
//   * operands are staged in shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a
//     STAGES-deep mbarrier ring; one elected thread issues tcgen05.mma (UMMA 128 x BN x 16), and
//     tcgen05.commit releases the ring slot / publishes the accumulator;
//   * the 128 x BN fp32 accumulator lives in TMEM (BN columns); four epilogue warps read it back
//     with tcgen05.ld (32 lanes x 32 columns per instruction) and apply the fused epilogue
//     (bias / GELU / dropout / residual / dGELU / split-K accumulate) straight from registers;
//   * either operand may be K-major (row = M/N index, K contiguous) or MN-major (row = K index),
//     so fprop (K,K), dgrad (K,MN) and wgrad (MN,MN) all run without transposed copies.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue
// (warp 2 also owns the TMEM allocation).  Shared-memory use is kept <= ~100 KB so two CTAs are
// co-resident per SM: one CTA's epilogue overlaps the other's main loop.
#pragma once
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include "epilogue.cuh"

extern int g_gct_persist;
extern int g_gct_tma_store;
extern int g_gct_sm_budget;
extern int g_gct_res_box;
extern int g_gct_ew4;
extern int g_gct_pair;

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster each hold their own 128 rows of A and half of B's rows;
// the leader (cluster rank 0) issues one M = 256 MMA per k-step, completion is multicast to the barriers of both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose transaction bytes are credited to a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D=f32, A=B=bf16, majorness bits, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 8 consecutive columns of one row through the fused epilogue, vectorised when aligned
template <typename T>
__device__ __forceinline__ void epilogue_vec8(const Epilogue& e, int row, int col, const float* acc, int N, bool atomic) {
    const bool fast = (col + 8 <= N) && ((e.ldc & 7) == 0) && !(e.flags & EPI_BIAS_ROW);
    if (!fast) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (col + j < N) epilogue_apply<T>(e, row, col + j, acc[j], atomic);
        return;
    }
    const size_t idx = (size_t)row * e.ldc + col;
    f8 v;
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = acc[j] * e.alpha;
    if (e.bias) {
        f8 b = ld8(e.bias + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] += b.v[j];
    }
    if (e.flags & EPI_GELU) {
        if (e.flags & EPI_GELU_GRAD) {
            f8 dg;
#pragma unroll
            for (int j = 0; j < 8; ++j) dg.v[j] = drop_mult(e.drop, idx + j) * gelu_fast_grad(v.v[j]);
            st8(reinterpret_cast<T*>(e.aux_out) + idx, dg);
        } else if (e.aux_out) st8(reinterpret_cast<T*>(e.aux_out) + idx, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] = gelu_fast(v.v[j]);
    }
    if (e.drop.thresh) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] = drop_apply(e.drop, idx + j, v.v[j]);
    }
    if (e.flags & (EPI_DGELU | EPI_MUL_AUX)) {
        f8 h = ld8(reinterpret_cast<const T*>(e.aux_in) + idx);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] *= (e.flags & EPI_DGELU) ? gelu_fast_grad(h.v[j]) : h.v[j];
    }
    if (e.res32) {
        f8 r = ld8(e.res32 + idx);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] += r.v[j];
    }
    if (e.flags & EPI_ACCUM) {
        if (atomic) {
            atomicAdd(reinterpret_cast<float4*>(e.out32 + idx), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
            atomicAdd(reinterpret_cast<float4*>(e.out32 + idx + 4), make_float4(v.v[4], v.v[5], v.v[6], v.v[7]));
        } else {
            f8 o = ld8(e.out32 + idx);
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] += v.v[j];
            st8(e.out32 + idx, o);
        }
        return;
    }
    if (e.out32) st8(e.out32 + idx, v);
    if (e.outT) st8(reinterpret_cast<T*>(e.outT) + idx, v);
}

// 32 consecutive columns of one row: the global operands of the epilogue (bias, residual, dGELU pre-activation)
// are requested first, as independent 16/32-byte loads, and only then is the accumulator pulled out of TMEM, so
// their latency overlaps the tcgen05.ld instead of serialising behind it.
template <int NJ> struct EpiOperandsT { f8 b[NJ], r[NJ], h[NJ]; };
using EpiOperands = EpiOperandsT<4>;

template <typename T, int NJ = 4>
__device__ __forceinline__ bool epi_fast(const Epilogue& e, int col, int N) {
    return (col + 8 * NJ <= N) && ((e.ldc & 7) == 0) && !(e.flags & EPI_BIAS_ROW);
}
template <typename T, int NJ = 4>
__device__ __forceinline__ void epi_prefetch(const Epilogue& e, int row, int col, EpiOperandsT<NJ>& o) {
    const size_t idx = (size_t)row * e.ldc + col;
    if (e.bias) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) o.b[j] = ld8(e.bias + col + 8 * j);
    }
    if (e.res32) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) o.r[j] = ld8(e.res32 + idx + 8 * j);
    }
    if (e.flags & (EPI_DGELU | EPI_MUL_AUX)) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) o.h[j] = ld8(reinterpret_cast<const T*>(e.aux_in) + idx + 8 * j);
    }
}
template <typename T, int NJ = 4>
__device__ __forceinline__ void epi_finish(const Epilogue& e, int row, int col, const float* acc, const EpiOperandsT<NJ>& o,
                                           bool atomic) {
    const size_t idx = (size_t)row * e.ldc + col;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        f8 v;
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] = acc[8 * j + k] * e.alpha;
        if (e.bias) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] += o.b[j].v[k];
        }
        if (e.flags & EPI_GELU) {
            if (e.flags & EPI_GELU_GRAD) {
                f8 dg;
#pragma unroll
                for (int k = 0; k < 8; ++k) dg.v[k] = drop_mult(e.drop, idx + 8 * j + k) * gelu_fast_grad(v.v[k]);
                st8(reinterpret_cast<T*>(e.aux_out) + idx + 8 * j, dg);
            } else if (e.aux_out) st8(reinterpret_cast<T*>(e.aux_out) + idx + 8 * j, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] = gelu_fast(v.v[k]);
        }
        if (e.drop.thresh) {
            const uint32_t pair0 = (uint32_t)((idx + 8 * j) >> 1);      // idx is even; tensors stay below 2^33 elements
#pragma unroll
            for (int k = 0; k < 4; ++k) drop_pair(e.drop, pair0 + k, v.v[2 * k], v.v[2 * k + 1]);
        }
        if (e.flags & EPI_DGELU) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] *= gelu_fast_grad(o.h[j].v[k]);
        }
        if (e.flags & EPI_MUL_AUX) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] *= o.h[j].v[k];
        }
        if (e.res32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] += o.r[j].v[k];
        }
        if (e.flags & EPI_ACCUM) {
            if (atomic) {
                atomicAdd(reinterpret_cast<float4*>(e.out32 + idx + 8 * j), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
                atomicAdd(reinterpret_cast<float4*>(e.out32 + idx + 8 * j + 4), make_float4(v.v[4], v.v[5], v.v[6], v.v[7]));
            } else {
                f8 c = ld8(e.out32 + idx + 8 * j);
#pragma unroll
                for (int k = 0; k < 8; ++k) c.v[k] += v.v[k];
                st8(e.out32 + idx + 8 * j, c);
            }
            continue;
        }
        if (e.flags & EPI_NOSTORE) { if (v.v[0] == 1.2345e30f) st8(e.out32 + idx, v); continue; }
        if (e.out32) st8(e.out32 + idx + 8 * j, v);
        if (e.outT) st8(reinterpret_cast<T*>(e.outT) + idx + 8 * j, v);
    }
}

__device__ __forceinline__ uint32_t stage_off(int row, int chunk16);
__device__ __forceinline__ void stage_wait_free(int lane);

// Compile-time specialised epilogue for one 16-column chunk of one row (the persistent kernel's hot path):
// no per-element branches, no 64-bit index arithmetic (the caller passes element offsets), operands requested
// before the TMEM load.  ACT: 0 none, 1 GELU (pre-activation saved to aux_out), 2 dGELU (pre-activation from aux_in).
// With stg != nullptr the results are not stored to global memory by this thread: they are parked in the warp's
// swizzled staging tile (32 rows x 128 B) at 16-byte chunk `sc` (and `sc_aux` for the saved GELU pre-activation) and
// leave later as whole row segments (stage_flush), which cuts the number of L2 write requests by 4-8x.
// wait_stg: the staging tile may still be read by the TMA store of the previous segment (stage_tma_store_async): wait for that
// read just before the first write, i.e. AFTER this chunk's TMEM load and math, which is what hides the store's latency.
template <bool HAS_BIAS, int ACT, bool HAS_RES, bool OUT_F32, int WHICH = 0>
__device__ __forceinline__ void epi_chunk16(const Epilogue& e, const float* __restrict__ bias, uint32_t taddr, size_t off, int col,
                                            uint8_t* stg = nullptr, int lane = 0, int sc = 0, int sc_aux = 0, bool wait_stg = false) {
    float4 b4[4], r4[4];
    uint4 h2[2];
    if constexpr (HAS_BIAS) {
#pragma unroll
        for (int i = 0; i < 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
    }
    if constexpr (HAS_RES) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r4[i] = *(reinterpret_cast<const float4*>(e.res32 + off) + i);
    }
    if constexpr (ACT == 2 || ACT == 4) {
#pragma unroll
        for (int i = 0; i < 2; ++i) h2[i] = *(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.aux_in) + off) + i);
    }
    float v[16];
    tmem_ld16(taddr, v);
    if constexpr (HAS_BIAS) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w; }
    }
    if constexpr (ACT == 1) {
        if (WHICH != 2 && e.aux_out) {
            if (wait_stg) stage_wait_free(lane);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint4 u;
                __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int i = 0; i < 4; ++i) hh[i] = __floats2bfloat162_rn(v[8 * g + 2 * i], v[8 * g + 2 * i + 1]);
                if (stg) *reinterpret_cast<uint4*>(stg + stage_off(lane, sc_aux + g)) = u;
                else *(reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.aux_out) + off) + g) = u;
            }
        }
        if constexpr (WHICH == 1) return;        // aux-only pass: the pre-activation is all that was wanted
        const uint32_t pair0 = (uint32_t)(off >> 1);
        float2 x[8], m[8], g[8];                 // packed fp32x2 GELU with the dropout multiplier folded in
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = make_float2(v[2 * i], v[2 * i + 1]); m[i] = drop_mult_pair(e.drop, pair0 + i); }
        gelu_pairs<4, false>(x, m, g, nullptr);
        gelu_pairs<4, false>(x + 4, m + 4, g + 4, nullptr);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = g[i].x; v[2 * i + 1] = g[i].y; }
    } else if (e.drop.thresh) {
        const uint32_t pair0 = (uint32_t)(off >> 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) drop_pair(e.drop, pair0 + i, v[2 * i], v[2 * i + 1]);
    }
    if constexpr (ACT == 4) {
        const uint32_t* hw = reinterpret_cast<const uint32_t*>(h2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 hf = make_float2(__uint_as_float(hw[i] << 16), __uint_as_float(hw[i] & 0xffff0000u));
            const float2 r = __fmul2_rn(make_float2(v[2 * i], v[2 * i + 1]), hf);
            v[2 * i] = r.x; v[2 * i + 1] = r.y;
        }
    }
    if constexpr (ACT == 2) {       // v *= gelu'(pre) (the dropout multiplier was applied above), packed evaluation
        const uint32_t* hw = reinterpret_cast<const uint32_t*>(h2);
        float2 x[8], one[8], g[8], dg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = make_float2(__uint_as_float(hw[i] << 16), __uint_as_float(hw[i] & 0xffff0000u));
            one[i] = make_float2(1.f, 1.f);
        }
        gelu_pairs<4, true>(x, one, g, dg);
        gelu_pairs<4, true>(x + 4, one + 4, g + 4, dg + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 r = __fmul2_rn(make_float2(v[2 * i], v[2 * i + 1]), dg[i]);
            v[2 * i] = r.x; v[2 * i + 1] = r.y;
        }
    }
    if constexpr (HAS_RES) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[4 * i] += r4[i].x; v[4 * i + 1] += r4[i].y; v[4 * i + 2] += r4[i].z; v[4 * i + 3] += r4[i].w; }
    }
    if (e.flags & 64) {      // measurement hook: keep the math, drop (almost all) stores
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += v[i];
        if (sum != 1.2345e30f) return;
    }
    if (WHICH != 1 && wait_stg) stage_wait_free(lane);
    if constexpr (OUT_F32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 f = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            if (stg) *reinterpret_cast<float4*>(stg + stage_off(lane, sc + i)) = f;
            else *(reinterpret_cast<float4*>(e.out32 + off) + i) = f;
        }
    } else {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            uint4 u;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) hh[i] = __floats2bfloat162_rn(v[8 * g + 2 * i], v[8 * g + 2 * i + 1]);
            if (stg) *reinterpret_cast<uint4*>(stg + stage_off(lane, sc + g)) = u;
            else *(reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.outT) + off) + g) = u;
        }
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// GELU forward that also saves the gradient factor, 16 columns of one row in ONE pass over the accumulator:
//   gp[8] = bf16x2 pairs of keep*gelu(acc + bias),  dp[8] = bf16x2 pairs of keep*gelu'(acc + bias)
__device__ __forceinline__ void epi_gelu16(const Epilogue& e, const float* __restrict__ bias, uint32_t taddr, size_t off, int col,
                                           uint32_t* gp, uint32_t* dp) {
    float4 b4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
    float v[16];
    tmem_ld16(taddr, v);
    const uint32_t pair0 = (uint32_t)(off >> 1);
    float2 x[8], m[8], g[8], dg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 bb = (i & 1) ? make_float2(b4[i >> 1].z, b4[i >> 1].w) : make_float2(b4[i >> 1].x, b4[i >> 1].y);
        x[i] = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), bb);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = drop_mult_pair<false>(e.drop, pair0 + i);
    gelu_pairs<4, true>(x, m, g, dg);
    gelu_pairs<4, true>(x + 4, m + 4, g + 4, dg + 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        gp[i] = pack_bf16x2(g[i].x, g[i].y);
        dp[i] = pack_bf16x2(dg[i].x, dg[i].y);
    }
}

// One elected lane hands the warp's staged 32-row x 128-byte tile to the TMA engine (tensor store, SWIZZLE_128B box);
// rows / columns outside the tensor are clipped by the hardware.  Returns once the tile may be overwritten.
__device__ __forceinline__ void stage_tma_store(const CUtensorMap* map, uint32_t stg_smem, int col, int row, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                         reinterpret_cast<uint64_t>(map)),
                     "r"(col), "r"(row), "r"(stg_smem)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
}

// The same store without the wait: the caller calls stage_wait_free before it writes the tile again (and before the kernel ends).
__device__ __forceinline__ void stage_tma_store_async(const CUtensorMap* map, uint32_t stg_smem, int col, int row, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                         reinterpret_cast<uint64_t>(map)),
                     "r"(col), "r"(row), "r"(stg_smem)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
__device__ __forceinline__ void stage_wait_free(int lane) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
}

// Warp-cooperative write-out of `nchunk` 16-byte chunks per row (starting at staging chunk `sc0`) of the warp's
// 32 staged rows: consecutive lanes cover one row segment, so each request is a whole 64- or 128-byte piece of a row.
__device__ __forceinline__ void stage_flush(const uint8_t* stg, int lane, int sc0, int nchunk, uint8_t* gbase /* row 0 */,
                                            size_t row_pitch_bytes, int rows_valid) {
    const int rows_per_it = 32 / nchunk;
    const int piece = lane % nchunk, rsub = lane / nchunk;
    for (int r0 = 0; r0 < 32; r0 += rows_per_it) {
        const int r = r0 + rsub;
        if (r < rows_valid) {
            const uint4 u = *reinterpret_cast<const uint4*>(stg + stage_off(r, sc0 + piece));
            __stcs(reinterpret_cast<uint4*>(gbase + (size_t)r * row_pitch_bytes + (size_t)piece * 16), u);
        }
    }
}

// epilogue "mode": which specialisation serves this launch (0 = generic runtime-flag path)
__host__ __device__ inline int epi_mode(const Epilogue& e) {
    if ((e.flags & (EPI_ACCUM | EPI_BIAS_ROW | EPI_NOSTORE)) || e.alpha != 1.f || (e.ldc & 7)) return 0;
    const bool b = e.bias != nullptr, r = e.res32 != nullptr, f = e.out32 != nullptr, t = e.outT != nullptr;
    if (f == t) return 0;                                    // exactly one output
    if (e.flags & EPI_GELU) {
        if (e.flags & EPI_GELU_GRAD) return (b && !r && t && e.aux_out) ? 9 : 0;
        return (b && !r && t) ? 3 : 0;
    }
    if (e.flags & EPI_DGELU) return (!b && !r && t) ? 6 : 0;
    if (e.flags & EPI_MUL_AUX) return (!b && !r && t) ? 10 : 0;
    if (t) return r ? 0 : (b ? 1 : 5);
    return b ? (r ? 2 : 7) : (r ? 8 : 4);
}

// Warp-private staging tile: 32 rows x 128 bytes, 16-byte chunks XOR-swizzled by (row & 7).
__device__ __forceinline__ uint32_t stage_off(int row, int chunk16) { return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4)); }

template <int BN, bool A_MN, bool B_MN, int STAGES>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024 /* alignment slack */;
};

template <int BN, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(192)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
               int kb_per_split, Epilogue epi) {
    using L = SmemLayout<BN, A_MN, B_MN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + L::BAR_OFF;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);
    const uint32_t tmem_slot = tmem_full_bar + 8u;
    uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int num_kb = (K + BK - 1) / BK;
    const int kb_begin = blockIdx.z * kb_per_split;
    const int kb_end = min(num_kb, kb_begin + kb_per_split);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)(BN < 32 ? 32 : BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    // everything above touched only kernel parameters and on-chip state: under programmatic dependent launch it
    // overlaps the previous kernel's tail.  Operands / residuals written by that kernel are read only after this wait.
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(empty_bar(s), phase ^ 1u);
                mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
                const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
                if constexpr (!A_MN) {
                    tma_load_2d(sa, &tmA, kb * BK, m0, full_bar(s));
                } else {
#pragma unroll
                    for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m0 + c * 64, kb * BK, full_bar(s));
                }
                if constexpr (!B_MN) {
                    tma_load_2d(sb, &tmB, kb * BK, n0, full_bar(s));
                } else {
#pragma unroll
                    for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n0 + c * 64, kb * BK, full_bar(s));
                }
                if (++s == STAGES) { s = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
            int s = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(full_bar(s), phase);
                tcgen05_fence_after();
                const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
                    const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
                    umma_bf16(tmem_base, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar(s));       // frees the ring slot once these MMAs have read it
                if (++s == STAGES) { s = 0; phase ^= 1u; }
            }
            umma_commit(tmem_full_bar);          // accumulator complete
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int q = warp & 3;
        if (kb_end > kb_begin) {
            mbar_wait(tmem_full_bar, 0);
            tcgen05_fence_after();
            const int row = m0 + q * 32 + lane;
            const bool atomic = gridDim.z > 1;
            if (blockIdx.z != 0) epi.bias = nullptr;      // split-K: the bias is added once
            if constexpr (BN == 16) {
                float acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16), acc);
                if (row < M) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = n0 + j * 8;
                        if (col < N) epilogue_vec8<bf16>(epi, row, col, acc + j * 8, N, atomic);
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int col0 = n0 + c * 32;
                    const bool fast = row < M && epi_fast<bf16>(epi, col0, N);
                    EpiOperands ops;
                    if (fast) epi_prefetch<bf16>(epi, row, col0, ops);
                    float acc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), acc);
                    if (fast) {
                        epi_finish<bf16>(epi, row, col0, acc, ops, atomic);
                    } else if (row < M) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int col = col0 + j * 8;
                            if (col < N) epilogue_vec8<bf16>(epi, row, col, acc + j * 8, N, atomic);
                        }
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)(BN < 32 ? 32 : BN))
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// Persistent variant for problems with more tiles than SMs: one CTA per SM walks the tile list.
// The TMA producer runs ahead across tile boundaries (no pipeline refill bubble per tile), the
// accumulator is double-buffered in TMEM (2 x BN columns) so the eight epilogue warps drain tile i
// while the MMA warp already accumulates tile i+1.
// Warp roles (576 threads): 0 = TMA producer, 1 = MMA issuer, 2..17 = epilogue (TMEM lane quarter =
// warp % 4, column slice = (warp - 2) / 4); warp 2 owns the TMEM allocation.
// ------------------------------------------------------------------------------------------
template <int BN, int STAGES, int EW = 2, int CG = 1>
struct PersistSmem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (BN / CG) * BK * 2;            // CTA pair: each CTA stages half of B's rows
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;            // 4*EW epilogue warps x 4 KB
    static constexpr int BAR_OFF = STAGING_OFF + 4 * EW * 4096;
    static constexpr int RBAR_OFF = BAR_OFF + (2 * STAGES + 4) * 8 + 16;    // one residual-box barrier per epilogue warp
    static constexpr int TOTAL = RBAR_OFF + 4 * EW * 8 + 1024;
    static constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));
};

// CG = 2: launched as clusters of two CTAs (a CTA pair on one TPC).  One work item is a 256 x BN output tile: each CTA
// loads its own 128 rows of A and BN/2 rows of B (32 KB per k-block instead of 48 KB: these short-K GEMMs are bound by
// the L2 -> shared-memory operand fill), the leader issues tcgen05.mma.cta_group::2 (M = 256), each CTA drains its own
// 128 x BN accumulator.
template <int BN, bool A_MN, bool B_MN, int STAGES, int EW = 2, int CG = 1>
__global__ void __launch_bounds__((2 + 4 * EW) * 32, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                       const __grid_constant__ CUtensorMap tmR, int M, int N, int K,
                       int kb_per_split, int num_splits, Epilogue epi, int use_tma_store) {
    using L = PersistSmem<BN, STAGES, EW, CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + L::BAR_OFF;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + BK - 1) / BK;
    constexpr int BMT = BM * CG;                                  // rows of one work item (pair: 256)
    const int m_tiles = (M + BMT - 1) / BMT, n_tiles = (N + BN - 1) / BN;
    const int total = m_tiles * n_tiles * num_splits;
    const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;    // 0 = leader
    const int tile0 = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;
    const int m_rank_off = (int)crank * BM;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        // pair: the leader's full barrier collects one arrive.expect_tx per CTA, its accumulator-empty barrier the epilogue
        // warps of both CTAs
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), CG); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * EW * CG); }
        for (int w = 0; w < 4 * EW; ++w) mbar_init(base + L::RBAR_OFF + 8u * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)L::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"((uint32_t)L::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();          // the peer's barriers are initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < total; tile += tile_step) {
                const int n0 = (tile % n_tiles) * BN, m0 = ((tile / n_tiles) % m_tiles) * BMT + m_rank_off, z = tile / (n_tiles * m_tiles);
                const int kb_begin = z * kb_per_split, kb_end = min(num_kb, kb_begin + kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(empty_bar(s), phase ^ 1u);
                    const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
                    if constexpr (CG == 2) {
                        // this CTA's rows of A and its half of B's rows; bytes credited to the leader's full barrier
                        const uint32_t lead_full = mapa_u32(full_bar(s), 0);
                        mbar_expect_tx_cluster(lead_full, L::STAGE_BYTES);
                        if constexpr (!A_MN) {
                            tma_load_2d_pair(sa, &tmA, kb * BK, m0, lead_full);
                        } else {
#pragma unroll
                            for (int c = 0; c < BM / 64; ++c) tma_load_2d_pair(sa + c * 8192, &tmA, m0 + c * 64, kb * BK, lead_full);
                        }
                        const int nh = n0 + (int)crank * (BN / 2);
                        if constexpr (!B_MN) {
                            tma_load_2d_pair(sb, &tmB, kb * BK, nh, lead_full);
                        } else {
#pragma unroll
                            for (int c = 0; c < BN / 2 / 64; ++c) tma_load_2d_pair(sb + c * 8192, &tmB, nh + c * 64, kb * BK, lead_full);
                        }
                    } else {
                        mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
                        if constexpr (!A_MN) {
                            tma_load_2d(sa, &tmA, kb * BK, m0, full_bar(s));
                        } else {
#pragma unroll
                            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m0 + c * 64, kb * BK, full_bar(s));
                        }
                        if constexpr (!B_MN) {
                            tma_load_2d(sb, &tmB, kb * BK, n0, full_bar(s));
                        } else {
#pragma unroll
                            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n0 + c * 64, kb * BK, full_bar(s));
                        }
                    }
                    if (++s == STAGES) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {                    // pair: only the leader issues MMAs
            constexpr uint32_t idesc = make_idesc(BMT, BN, A_MN, B_MN);
            int s = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = tile0; tile < total; tile += tile_step) {
                const int z = tile / (n_tiles * m_tiles);
                const int kb_begin = z * kb_per_split, kb_end = min(num_kb, kb_begin + kb_per_split);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);       // epilogue (of both CTAs) has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(full_bar(s), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
                        if constexpr (CG == 2) umma_bf16_pair(tacc, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                        else umma_bf16(tacc, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                    }
                    if constexpr (CG == 2) umma_commit_pair(empty_bar(s)); else umma_commit(empty_bar(s));
                    if (++s == STAGES) { s = 0; phase ^= 1u; }
                }
                if constexpr (CG == 2) umma_commit_pair(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // 4*EW epilogue warps: TMEM lane quarter q = warp % 4, column slice cs = (warp - 2) / 4 owns BN/EW columns,
        // walked in 16-column chunks; results leave through the warp's 4 KB staging tile as whole row segments
        const int q = warp & 3, cs = (warp - 2) >> 2;
        constexpr int SLICE = BN / EW;
        constexpr int NCH = (SLICE + 15) / 16;
        int acc = 0;
        uint32_t acc_phase = 0, res_phase = 0;
        const bool atomic = num_splits > 1;
        const int mode = epi_mode(epi);
        for (int tile = tile0; tile < total; tile += tile_step) {
            const int n0 = (tile % n_tiles) * BN, m0 = ((tile / n_tiles) % m_tiles) * BMT + m_rank_off, z = tile / (n_tiles * m_tiles);
            Epilogue e = epi;
            if (z != 0) e.bias = nullptr;
            const int row = ((e.flags & 32) ? 0 : m0) + q * 32 + lane;      // flag 32: measurement hook, every tile writes rows [0,128)
            bool waited = false;
            const uint32_t tslice = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cs * SLICE);
            if (mode != 0 && n0 + BN <= N && m0 + BM <= M) {
                // hot path: full tile (keeps tcgen05.ld warp-uniform), compile-time specialised epilogue
                // fp32 residual (modes 2 / 8) as TMA boxes: the 32-row x 32-column box of a segment lands in the warp's staging tile
                // (coalesced, no L1 tag traffic), the lane reads its own row from shared memory, adds, writes the result over it and
                // the same tile leaves as the output box.  A lane reading its own row from global memory makes every 16-byte request
                // of the warp hit a different 128-byte line ([B200] gemm_rownorm.cuh: 97 -> 79 us from this change alone).
                const bool res_box = (use_tma_store & 2) && (mode == 2 || mode == 8) && !(e.flags & (32 | 64)) && (SLICE % 32) == 0;
                const uint32_t rbar = base + L::RBAR_OFF + 8u * (uint32_t)(warp - 2);
                uint8_t* stg_r = smem_raw + (base - smem_u32(smem_raw)) + L::STAGING_OFF + (warp - 2) * 4096;
                const bool aux_box = (use_tma_store & 4) && mode == 10 && !(e.flags & (32 | 64)) && (SLICE % 64) == 0;
                if ((res_box || aux_box) && lane == 0) {        // first box: in flight before the accumulator is ready
                    mbar_expect_tx(rbar, 4096);
                    tma_load_2d(smem_u32(stg_r), &tmR, n0 + cs * SLICE, m0 + q * 32, rbar);
                }
                mbar_wait(tfull_bar(acc), acc_phase);
                tcgen05_fence_after();
                waited = true;
                if (res_box) {
                    const int colb = n0 + cs * SLICE;
                    const size_t off0 = (size_t)row * e.ldc + colb;
                    const float* bias = (z != 0) ? nullptr : e.bias;
                    const uint32_t stg_s = smem_u32(stg_r);
                    constexpr int NSEG = SLICE / 32;
#pragma unroll 1
                    for (int sg = 0; sg < NSEG; ++sg) {
                        float v[2][16];
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2) {
                            const int col = colb + sg * 32 + c2 * 16;
                            tmem_ld16(tslice + sg * 32 + c2 * 16, v[c2]);
                            if (bias) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
                                    v[c2][4 * i] += b.x; v[c2][4 * i + 1] += b.y; v[c2][4 * i + 2] += b.z; v[c2][4 * i + 3] += b.w;
                                }
                            }
                            if (e.drop.thresh) {
                                const uint32_t pair0 = (uint32_t)((off0 + sg * 32 + c2 * 16) >> 1);
#pragma unroll
                                for (int i = 0; i < 8; ++i) drop_pair(e.drop, pair0 + i, v[c2][2 * i], v[c2][2 * i + 1]);
                            }
                        }
                        mbar_wait(rbar, res_phase);               // the segment's residual box has landed
                        res_phase ^= 1u;
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float4* slot = reinterpret_cast<float4*>(stg_r + stage_off(lane, c2 * 4 + i));
                                const float4 r4 = *slot;
                                *slot = make_float4(v[c2][4 * i] + r4.x, v[c2][4 * i + 1] + r4.y, v[c2][4 * i + 2] + r4.z, v[c2][4 * i + 3] + r4.w);
                            }
                        }
                        stage_tma_store(&tmC, stg_s, colb + sg * 32, m0 + q * 32, lane);       // also frees the tile for the next box
                        if (sg + 1 < NSEG && lane == 0) {
                            mbar_expect_tx(rbar, 4096);
                            tma_load_2d(stg_s, &tmR, colb + (sg + 1) * 32, m0 + q * 32, rbar);
                        }
                    }
                } else if (aux_box) {
                    // multiply-by-aux (backward of the FFN's first projection): the saved keep * gelu'(pre) factor arrives as 32-row x
                    // 64-column bf16 boxes in the staging tile, the product leaves from the same tile
                    const int colb = n0 + cs * SLICE;
                    const size_t off0 = (size_t)row * e.ldc + colb;
                    const uint32_t stg_s = smem_u32(stg_r);
                    constexpr int NSEG = SLICE / 64;
#pragma unroll 1
                    for (int sg = 0; sg < NSEG; ++sg) {
                        mbar_wait(rbar, res_phase);
                        res_phase ^= 1u;
#pragma unroll 1
                        for (int ci = 0; ci < 4; ++ci) {
                            float v[16];
                            tmem_ld16(tslice + sg * 64 + ci * 16, v);
                            if (e.drop.thresh) {
                                const uint32_t pair0 = (uint32_t)((off0 + sg * 64 + ci * 16) >> 1);
#pragma unroll
                                for (int i = 0; i < 8; ++i) drop_pair(e.drop, pair0 + i, v[2 * i], v[2 * i + 1]);
                            }
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                uint4* slot = reinterpret_cast<uint4*>(stg_r + stage_off(lane, ci * 2 + g));
                                const uint4 h = *slot;
                                const uint32_t* hw = reinterpret_cast<const uint32_t*>(&h);
                                uint4 o;
                                uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 hf = make_float2(__uint_as_float(hw[i] << 16), __uint_as_float(hw[i] & 0xffff0000u));
                                    ow[i] = pack_bf16x2(v[8 * g + 2 * i] * hf.x, v[8 * g + 2 * i + 1] * hf.y);
                                }
                                *slot = o;
                            }
                        }
                        stage_tma_store(&tmC, stg_s, colb + sg * 64, m0 + q * 32, lane);
                        if (sg + 1 < NSEG && lane == 0) {
                            mbar_expect_tx(rbar, 4096);
                            tma_load_2d(stg_s, &tmR, colb + (sg + 1) * 64, m0 + q * 32, rbar);
                        }
                    }
                } else {
                    const int colb = n0 + cs * SLICE;
                    const size_t off0 = (size_t)row * e.ldc + colb;
                    const float* bias = (z != 0) ? nullptr : e.bias;
                    uint8_t* stg = smem_raw + (base - smem_u32(smem_raw)) + L::STAGING_OFF + (warp - 2) * 4096;
                    const int rows_valid = 32;                      // full tile
                    const size_t row0_off = (size_t)(m0 + q * 32) * e.ldc + colb;     // element offset of staged row 0
                    // use_tma_store bit 3: TMA stores of the staging tile are not waited for where they are issued but just
                    // before the tile's next write (after the next segment's TMEM load and math) and at the end of the kernel
                    const bool defer = (use_tma_store & 8) != 0;
#define GCT_EPI_RUN(HB, ACT_, HR, F32)                                                                                  \
    {                                                                                                                   \
        constexpr int ESZ = F32 ? 4 : 2;                                                                                \
        constexpr int SEGC = F32 ? 32 : 64;                           /* columns per 128-byte row segment */             \
        constexpr int SEG_COLS = SEGC < SLICE ? SEGC : SLICE;                                                           \
        constexpr int CPS = SEG_COLS / 16;                            /* chunks per segment */                          \
        constexpr int C16 = 16 * ESZ / 16;                            /* 16-byte pieces per 16-column chunk */            \
        const uint32_t stg_s = smem_u32(stg);                                                                           \
        _Pragma("unroll 1") for (int c0 = 0; c0 < NCH; c0 += CPS) {                                                     \
            const int seg0 = c0 * 16;                                 /* first column of the segment (slice-relative) */ \
            if (ACT_ == 1 && e.aux_out) {                             /* pass A: saved pre-activation */                 \
                _Pragma("unroll 1") for (int ci = 0; ci < CPS; ++ci)                                                    \
                    epi_chunk16<HB, ACT_, HR, F32, 1>(e, bias, tslice + (c0 + ci) * 16, off0 + (c0 + ci) * 16,           \
                                                      colb + (c0 + ci) * 16, stg, lane, 0, ci * 2, defer && ci == 0);     \
                if (use_tma_store && SEG_COLS * 2 == 128) {                                                             \
                    if (defer) stage_tma_store_async(&tmAux, stg_s, colb + seg0, m0 + q * 32, lane);                    \
                    else stage_tma_store(&tmAux, stg_s, colb + seg0, m0 + q * 32, lane);                                \
                } else {                                                                                                \
                    __syncwarp();                                                                                       \
                    stage_flush(stg, lane, 0, CPS * 2, reinterpret_cast<uint8_t*>(e.aux_out) + (row0_off + seg0) * 2,   \
                                (size_t)e.ldc * 2, rows_valid);                                                         \
                    __syncwarp();                                                                                       \
                }                                                                                                       \
            }                                                                                                           \
            _Pragma("unroll 1") for (int ci = 0; ci < CPS; ++ci)                                                        \
                epi_chunk16<HB, ACT_, HR, F32, 2>(e, bias, tslice + (c0 + ci) * 16, off0 + (c0 + ci) * 16,               \
                                                  colb + (c0 + ci) * 16, stg, lane, ci * C16, 0, defer && ci == 0);      \
            if (use_tma_store && SEG_COLS * ESZ == 128) {                                                               \
                if (defer) stage_tma_store_async(&tmC, stg_s, colb + seg0, m0 + q * 32, lane);                          \
                else stage_tma_store(&tmC, stg_s, colb + seg0, m0 + q * 32, lane);                                      \
            } else {                                                                                                    \
                __syncwarp();                                                                                           \
                uint8_t* gout = F32 ? reinterpret_cast<uint8_t*>(e.out32) : reinterpret_cast<uint8_t*>(e.outT);         \
                stage_flush(stg, lane, 0, CPS * C16, gout + (row0_off + seg0) * ESZ, (size_t)e.ldc * ESZ, rows_valid);  \
                __syncwarp();                                                                                           \
            }                                                                                                           \
        }                                                                                                               \
    }
                    switch (mode) {
                        case 1: GCT_EPI_RUN(true, 0, false, false) break;
                        case 2: GCT_EPI_RUN(true, 0, true, true) break;
                        case 3: GCT_EPI_RUN(true, 1, false, false) break;
                        case 4: GCT_EPI_RUN(false, 0, false, true) break;
                        case 5: GCT_EPI_RUN(false, 0, false, false) break;
                        case 6: GCT_EPI_RUN(false, 2, false, false) break;
                        case 7: GCT_EPI_RUN(true, 0, false, true) break;
                        case 8: GCT_EPI_RUN(false, 0, true, true) break;
                        case 10: GCT_EPI_RUN(false, 4, false, false) break;
                        default: {       // 9: GELU + saved gradient factor, one pass over TMEM, two bf16 outputs
                            constexpr int SEG_COLS = 64 < SLICE ? 64 : SLICE;
                            constexpr int CPS = SEG_COLS / 16;
                            const uint32_t stg_s = smem_u32(stg);
                            const bool tma = use_tma_store && SEG_COLS * 2 == 128;
#pragma unroll 1
                            for (int c0 = 0; c0 < NCH; c0 += CPS) {
                                const int seg0 = c0 * 16;
                                uint32_t dgp[CPS][8];
#pragma unroll
                                for (int ci = 0; ci < CPS; ++ci) {
                                    uint32_t gp[8];
                                    epi_gelu16(e, bias, tslice + (c0 + ci) * 16, off0 + (c0 + ci) * 16, colb + (c0 + ci) * 16, gp, dgp[ci]);
                                    if (defer && ci == 0) stage_wait_free(lane);      // the previous segment's gradient-factor store
                                    *reinterpret_cast<uint4*>(stg + stage_off(lane, ci * 2)) = make_uint4(gp[0], gp[1], gp[2], gp[3]);
                                    *reinterpret_cast<uint4*>(stg + stage_off(lane, ci * 2 + 1)) = make_uint4(gp[4], gp[5], gp[6], gp[7]);
                                }
                                if (tma) {
                                    stage_tma_store(&tmC, stg_s, colb + seg0, m0 + q * 32, lane);
                                } else {
                                    __syncwarp();
                                    stage_flush(stg, lane, 0, CPS * 2, reinterpret_cast<uint8_t*>(e.outT) + (row0_off + seg0) * 2,
                                                (size_t)e.ldc * 2, rows_valid);
                                    __syncwarp();
                                }
#pragma unroll
                                for (int ci = 0; ci < CPS; ++ci) {
                                    *reinterpret_cast<uint4*>(stg + stage_off(lane, ci * 2)) = make_uint4(dgp[ci][0], dgp[ci][1], dgp[ci][2], dgp[ci][3]);
                                    *reinterpret_cast<uint4*>(stg + stage_off(lane, ci * 2 + 1)) = make_uint4(dgp[ci][4], dgp[ci][5], dgp[ci][6], dgp[ci][7]);
                                }
                                if (tma) {
                                    if (defer) stage_tma_store_async(&tmAux, stg_s, colb + seg0, m0 + q * 32, lane);
                                    else stage_tma_store(&tmAux, stg_s, colb + seg0, m0 + q * 32, lane);
                                } else {
                                    __syncwarp();
                                    stage_flush(stg, lane, 0, CPS * 2, reinterpret_cast<uint8_t*>(e.aux_out) + (row0_off + seg0) * 2,
                                                (size_t)e.ldc * 2, rows_valid);
                                    __syncwarp();
                                }
                            }
                        } break;
                    }
#undef GCT_EPI_RUN
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int coff = cs * SLICE + c * 16;
                    const int col0 = n0 + coff;
                    const bool fast = row < M && epi_fast<bf16, 2>(e, col0, N);
                    EpiOperandsT<2> ops;
                    if (fast) epi_prefetch<bf16, 2>(e, row, col0, ops);
                    if (!waited) { mbar_wait(tfull_bar(acc), acc_phase); tcgen05_fence_after(); waited = true; }
                    float v[16];
                    tmem_ld16(tslice + c * 16, v);
                    if (fast) {
                        epi_finish<bf16, 2>(e, row, col0, v, ops, atomic);
                    } else if (row < M) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int col = col0 + j * 8;
                            if (col < N) epilogue_vec8<bf16>(e, row, col, v + j * 8, N, atomic);
                        }
                    }
                }
            }
            if (!waited) { mbar_wait(tfull_bar(acc), acc_phase); tcgen05_fence_after(); }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0));      // the leader's MMA warp waits on it
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(acc)) : "memory");
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (use_tma_store & 8) stage_wait_free(lane);      // deferred TMA stores: the staging tile outlives their reads
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();            // no CTA leaves (or frees TMEM) while its peer may still signal it
    if (warp == 2) {
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::TMEM_COLS)
                         : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapKey {
    const void* p; uint64_t d0, d1, stride; uint32_t b0, b1; int esz;
    bool operator==(const MapKey& o) const {
        return p == o.p && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1 && esz == o.esz;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = (size_t)k.p;
        h = h * 1000003u ^ k.d0; h = h * 1000003u ^ k.d1; h = h * 1000003u ^ k.stride;
        h = h * 1000003u ^ k.b0; h = h * 1000003u ^ k.b1;
        return h;
    }
};

static int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride_bytes, uint32_t box_inner,
                          uint32_t box_outer, CUtensorMap* out, int esz = 2) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    static PFN_encodeTiled encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GCT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) GCT_FAIL(GCT_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    MapKey key{ptr, inner, outer, stride_bytes, box_inner, box_outer, esz};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return GCT_OK; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (stride_bytes & 15))
        GCT_FAIL(GCT_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (ptr=%p pitch=%llu)",
                 ptr, (unsigned long long)stride_bytes);
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = encode(&m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) GCT_FAIL(GCT_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d (dims %llu x %llu pitch %llu box %u x %u)",
                                    (int)r, (unsigned long long)inner, (unsigned long long)outer,
                                    (unsigned long long)stride_bytes, box_inner, box_outer);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return GCT_OK;
}

// Rank-3 map over a [batch][rows][inner] bf16 tensor (row pitch `row_pitch_bytes`, batch pitch rows * row pitch): boxes of
// box_inner x box_rows x 1, SWIZZLE_128B.  Rows past `rows` of a batch element are out of bounds for the map, so a store of a
// padded tile never touches the next element (the rank-2 operand maps treat [batch * rows] as one dimension).
static int get_tensor_map3(const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t row_pitch_bytes, uint32_t box_inner,
                           uint32_t box_rows, CUtensorMap* out) {
    struct Key3 {
        const void* p; uint64_t inner, rows, batch, pitch; uint32_t b0, b1;
        bool operator==(const Key3& o) const { return p == o.p && inner == o.inner && rows == o.rows && batch == o.batch && pitch == o.pitch && b0 == o.b0 && b1 == o.b1; }
    };
    struct Hash3 { size_t operator()(const Key3& k) const { size_t h = (size_t)k.p; h = h * 1000003u ^ k.inner; h = h * 1000003u ^ k.rows; h = h * 1000003u ^ k.batch; h = h * 1000003u ^ k.pitch; h = h * 1000003u ^ k.b1; return h; } };
    static std::mutex mu;
    static std::unordered_map<Key3, CUtensorMap, Hash3> cache;
    static PFN_encodeTiled encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GCT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) GCT_FAIL(GCT_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    Key3 key{ptr, inner, rows, batch, row_pitch_bytes, box_inner, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return GCT_OK; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_pitch_bytes & 15))
        GCT_FAIL(GCT_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (ptr=%p pitch=%llu)", ptr,
                 (unsigned long long)row_pitch_bytes);
    cuuint64_t dims[3] = {inner, rows, batch};
    cuuint64_t strides[2] = {row_pitch_bytes, rows * row_pitch_bytes};
    cuuint32_t box[3] = {box_inner, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) GCT_FAIL(GCT_ERR_CUDA, "cuTensorMapEncodeTiled (rank 3) failed: %d", (int)r);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return GCT_OK;
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(src_smem)
                 : "memory");
}

template <int BN, bool A_MN, bool B_MN, int STAGES>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int split_k, const Epilogue& epi,
                      cudaStream_t st) {
    using L = SmemLayout<BN, A_MN, B_MN, STAGES>;
    auto kern = gemm_tc_kernel<BN, A_MN, B_MN, STAGES>;
    GCT_SMEM_LIMIT(kern, L::TOTAL);
    const int num_kb = (K + BK - 1) / BK;
    int kps = (num_kb + split_k - 1) / split_k;
    split_k = (num_kb + kps - 1) / kps;
    dim3 grid(cdiv(N, BN), cdiv(M, BM), split_k);
    GCT_CUDA(launch_k(kern, grid, dim3(192), (size_t)L::TOTAL, st, true, ta, tb, M, N, K, kps, epi));
    return GCT_OK;
}

// Tensor map of the fp32 residual for the boxed-residual epilogue (32 columns x 32 rows, the output's staging geometry); sets
// bit 1 of use_tma when the launch qualifies (fp32 output through TMA stores, 16-byte aligned residual with the output's pitch).
static int residual_box_map(const Epilogue& epi, int M, int N, int& use_tma, CUtensorMap* out) {
    const int mode = epi_mode(epi);
    if (!g_gct_res_box || !(use_tma & 1)) return GCT_OK;
    if ((mode == 2 || mode == 8) && epi.res32 && epi.out32) {
        if ((reinterpret_cast<uintptr_t>(epi.res32) & 15) || ((size_t)epi.ldc * 4) % 16) return GCT_OK;
        GCT_TRY(get_tensor_map(epi.res32, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc * 4, 32, 32, out, 4));
        use_tma |= 2;
    } else if (mode == 10 && epi.aux_in && epi.outT) {      // multiply-by-aux: the bf16 factor as 64-column boxes
        if ((reinterpret_cast<uintptr_t>(epi.aux_in) & 15) || ((size_t)epi.ldc * 2) % 16) return GCT_OK;
        GCT_TRY(get_tensor_map(epi.aux_in, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc * 2, 64, 32, out, 2));
        use_tma |= 4;
    }
    return GCT_OK;
}

static int sm_count() {
    static int per_dev[GCT_MAX_DEVICES] = {};
    const int dev = gct_cur_device();
    int& n = per_dev[dev];
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    // SM budget of the persistent kernels: leaving a few SMs free lets a concurrently running NCCL kernel (overlapped gradient
    // exchange) get its CTAs without pushing some of a persistent GEMM's CTAs into a second wave
    return (g_gct_sm_budget > 0 && g_gct_sm_budget < n) ? g_gct_sm_budget : n;
}

template <int BN, bool A_MN, bool B_MN, int STAGES, int EW = 2>
static int launch_persist(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int split_k, const Epilogue& epi,
                          cudaStream_t st) {
    using L = PersistSmem<BN, STAGES, EW>;
    auto kern = gemm_tc_persist_kernel<BN, A_MN, B_MN, STAGES, EW>;
    GCT_SMEM_LIMIT(kern, L::TOTAL);
    const int num_kb = (K + BK - 1) / BK;
    int kps = (num_kb + split_k - 1) / split_k;
    split_k = (num_kb + kps - 1) / kps;
    const long long total = (long long)cdiv(M, BM) * cdiv(N, BN) * split_k;
    dim3 grid((unsigned)(total < sm_count() ? total : sm_count()));
    // output tensor maps for the TMA-store epilogue (32 rows x 128 bytes per store)
    CUtensorMap tc_ = ta, taux = ta;
    int use_tma = 0;
    if (g_gct_tma_store && epi_mode(epi) != 0) {
        const bool f32 = epi.out32 != nullptr;
        const void* cptr = f32 ? (const void*)epi.out32 : (const void*)epi.outT;
        const int esz = f32 ? 4 : 2;
        if ((reinterpret_cast<uintptr_t>(cptr) & 15) == 0 && ((size_t)epi.ldc * esz) % 16 == 0) {
            GCT_TRY(get_tensor_map(cptr, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc * esz, 128 / esz, 32, &tc_, esz));
            use_tma = 1 | (g_gct_tma_store == 1 ? 8 : 0);
            if ((epi.flags & EPI_GELU) && epi.aux_out)
                GCT_TRY(get_tensor_map(epi.aux_out, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc * 2, 64, 32, &taux, 2));
        }
    }
    CUtensorMap tr_ = tc_;
    GCT_TRY(residual_box_map(epi, M, N, use_tma, &tr_));
    GCT_CUDA(launch_k(kern, grid, dim3((2 + 4 * EW) * 32), (size_t)L::TOTAL, st, true, ta, tb, tc_, taux, tr_, M, N, K, kps, split_k, epi, use_tma));
    return GCT_OK;
}

// CTA-pair launch: clusters of two CTAs; a K-major B needs its own tensor map with a BN/2-row box (MN-major operands are
// fetched in 64-column boxes either way).
template <int BN, bool A_MN, bool B_MN, int STAGES, int EW>
static int launch_persist_pair(const CUtensorMap& ta, const CUtensorMap& tb_mn, const bf16* B, long long ldb, int M, int N, int K,
                               int split_k, const Epilogue& epi, cudaStream_t st) {
    using L = PersistSmem<BN, STAGES, EW, 2>;
    auto kern = gemm_tc_persist_kernel<BN, A_MN, B_MN, STAGES, EW, 2>;
    GCT_SMEM_LIMIT(kern, L::TOTAL);
    CUtensorMap tb = tb_mn;
    if (!B_MN) GCT_TRY(get_tensor_map(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, (uint32_t)(BN / 2), &tb));
    const int num_kb = (K + BK - 1) / BK;
    int kps = (num_kb + split_k - 1) / split_k;
    split_k = (num_kb + kps - 1) / kps;
    const long long total = (long long)cdiv(M, 2 * BM) * cdiv(N, BN) * split_k;
    long long ctas = 2 * total < (long long)(sm_count() & ~1) ? 2 * total : (long long)(sm_count() & ~1);
    CUtensorMap tc_ = ta, taux = ta;
    int use_tma = 0;
    if (g_gct_tma_store && epi_mode(epi) != 0) {
        const bool f32 = epi.out32 != nullptr;
        const void* cptr = f32 ? (const void*)epi.out32 : (const void*)epi.outT;
        const int esz = f32 ? 4 : 2;
        if ((reinterpret_cast<uintptr_t>(cptr) & 15) == 0 && ((size_t)epi.ldc * esz) % 16 == 0) {
            GCT_TRY(get_tensor_map(cptr, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc * esz, 128 / esz, 32, &tc_, esz));
            use_tma = 1 | (g_gct_tma_store == 1 ? 8 : 0);
            if ((epi.flags & EPI_GELU) && epi.aux_out)
                GCT_TRY(get_tensor_map(epi.aux_out, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc * 2, 64, 32, &taux, 2));
        }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3((2 + 4 * EW) * 32); cfg.dynamicSmemBytes = (size_t)L::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_gct_pdl ? 2 : 1;
    CUtensorMap tr_ = tc_;
    GCT_TRY(residual_box_map(epi, M, N, use_tma, &tr_));
    GCT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc_, taux, tr_, M, N, K, kps, split_k, epi, use_tma));
    return GCT_OK;
}

// A: K-major -> storage [M rows, K cols] with pitch lda; MN-major -> storage [K rows, M cols] with pitch lda.
// B: K-major -> storage [N rows, K cols] with pitch ldb; MN-major -> storage [K rows, N cols] with pitch ldb.
static int launch_gemm_tc(const bf16* A, bool a_mn, long long lda, const bf16* B, bool b_mn, long long ldb, int M, int N,
                          int K, int split_k, int bn_hint, const Epilogue& epi, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return GCT_OK;
    if (split_k < 1) split_k = 1;
    if (split_k > 1 && !(epi.flags & EPI_ACCUM)) GCT_FAIL(GCT_ERR_ARG, "split-K needs an accumulating epilogue");
    // bn_hint = BN + 1000*STAGES (either part may be 0 = choose here)
    int BN = bn_hint % 1000, ST = bn_hint / 1000;
    if (BN == 0) {
        // largest tile that still fills the SMs (wide tiles have the best smem-read : MMA ratio);
        // small-M (decode) problems fall through to narrow tiles
        const long long mt = (M + BM - 1) / BM, sp = split_k, sms = sm_count();
        if (N <= 32) BN = 32;
        else if (N <= 64) BN = 64;
        else if (N % 256 == 0 && mt * (N / 256) * sp >= sms) {
            BN = 256;
            // short-K projections with an fp32 residual + fp32 output are bound by their epilogue traffic, not by the MMA:
            // narrower tiles quantise better over the SMs ([B200] M=41472, N=K=512: 63 -> 59 us)
            const long long t256 = mt * (N / 256) * sp;
            const double eff256 = (double)t256 / (double)(((t256 + sms - 1) / sms) * sms);
            if (K <= 1024 && epi.res32 && epi.out32 && eff256 < 0.9) BN = 128;
        }
        else if (mt * ((N + 127) / 128) * sp >= sms) BN = 128;
        else if (mt * ((N + 63) / 64) * sp >= sms) BN = 64;
        else if (mt * ((N + 31) / 32) * sp >= 100 || a_mn || b_mn) BN = 32;
        else BN = 16;
    }
    if (b_mn && BN < 64) BN = 64;
    if (ST == 0) ST = (BN == 128) ? 3 : (BN == 16 ? 8 : 4);
    CUtensorMap ta, tb;
    if (!a_mn) GCT_TRY(get_tensor_map(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM, &ta));
    else GCT_TRY(get_tensor_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, BK, &ta));
    if (!b_mn) GCT_TRY(get_tensor_map(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, (uint32_t)BN, &tb));
    else GCT_TRY(get_tensor_map(B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, BK, &tb));

    if (g_gct_persist && (bn_hint / 1000) == 0 && (BN == 128 || BN == 256 || BN == 64) &&
        (long long)cdiv(M, BM) * cdiv(N, BN) * split_k > sm_count()) {
        // 16 epilogue warps (3 pipeline stages make room for their staging tiles) for the activation epilogues (GELU,
        // GELU + saved gradient, multiply-by-aux): twice the warps hide their MUFU / dependency / load latencies behind
        // each other ([B200] 129 -> 99 us, 136 -> 126 us, 161 -> 154 us); the plain modes keep 8 warps and 4 stages,
        // where the deeper pipeline is worth more than the extra warps (85 vs 95 us)
        // CTA pairs for the K-major (forward / decode) GEMMs with enough 256-row work items
        // (the plain-GELU epilogue of the decode FFN keeps the single-CTA kernel with 16 epilogue warps: 99 vs 129 us at M = 30000)
        if (g_gct_pair && BN == 256 && (long long)cdiv(M, 2 * BM) * cdiv(N, 256) * split_k * 2 > sm_count() &&
            !(g_gct_ew4 && epi_mode(epi) == 3)) {
            if (!a_mn && !b_mn) return launch_persist_pair<256, false, false, 6, 2>(ta, tb, B, ldb, M, N, K, split_k, epi, st);
            if (g_gct_pair > 1 && !a_mn && b_mn) return launch_persist_pair<256, false, true, 6, 2>(ta, tb, B, ldb, M, N, K, split_k, epi, st);
            if (g_gct_pair > 1 && a_mn && b_mn) return launch_persist_pair<256, true, true, 6, 2>(ta, tb, B, ldb, M, N, K, split_k, epi, st);
        }
        const int emode = epi_mode(epi);
        if (g_gct_ew4 && BN == 256 && (emode == 3 || emode == 6 || emode == 9 || emode == 10)) {
            if (!a_mn && !b_mn) return launch_persist<256, false, false, 3, 4>(ta, tb, M, N, K, split_k, epi, st);
            if (!a_mn && b_mn) return launch_persist<256, false, true, 3, 4>(ta, tb, M, N, K, split_k, epi, st);
        }
#define GCT_TCP_CASE(bn, amn, bmn, st_) \
        if (BN == bn && a_mn == amn && b_mn == bmn) return launch_persist<bn, amn, bmn, st_>(ta, tb, M, N, K, split_k, epi, st);
        GCT_TCP_CASE(64, false, false, 8)
        GCT_TCP_CASE(128, false, false, 6)
        GCT_TCP_CASE(256, false, false, 4)
        GCT_TCP_CASE(64, false, true, 8)
        GCT_TCP_CASE(128, false, true, 6)
        GCT_TCP_CASE(256, false, true, 4)
        GCT_TCP_CASE(64, true, true, 8)
        GCT_TCP_CASE(128, true, true, 6)
        GCT_TCP_CASE(256, true, true, 4)
#undef GCT_TCP_CASE
    }
#define GCT_TC_CASE(bn, amn, bmn, st_)                                                                   \
    if (BN == bn && a_mn == amn && b_mn == bmn && ST == st_) return launch_cfg<bn, amn, bmn, st_>(ta, tb, M, N, K, split_k, epi, st);
    GCT_TC_CASE(16, false, false, 4)
    GCT_TC_CASE(16, false, false, 8)
    GCT_TC_CASE(32, false, false, 4)
    GCT_TC_CASE(32, false, false, 8)
    GCT_TC_CASE(64, false, false, 4)
    GCT_TC_CASE(64, false, false, 8)
    GCT_TC_CASE(128, false, false, 3)
    GCT_TC_CASE(128, false, false, 6)
    GCT_TC_CASE(256, false, false, 4)
    GCT_TC_CASE(64, false, true, 4)
    GCT_TC_CASE(128, false, true, 3)
    GCT_TC_CASE(128, false, true, 6)
    GCT_TC_CASE(256, false, true, 4)
    GCT_TC_CASE(64, true, true, 4)
    GCT_TC_CASE(128, true, true, 3)
    GCT_TC_CASE(128, true, true, 6)
    GCT_TC_CASE(256, true, true, 4)
    GCT_TC_CASE(32, true, false, 4)
    GCT_TC_CASE(64, true, false, 4)
    GCT_TC_CASE(128, true, false, 3)
    GCT_TC_CASE(256, true, false, 4)
#undef GCT_TC_CASE
    GCT_FAIL(GCT_ERR_UNSUPPORTED, "no tcgen05 GEMM instantiation for BN=%d stages=%d a_mn=%d b_mn=%d", BN, ST, (int)a_mn, (int)b_mn);
}

}  // namespace tc
