// tcgen05 attention for the training / teacher-forced path: L <= 128, d_k = 64, bf16 operands.
// One CTA per (batch, head); thread t owns query row t (= TMEM lane t), so the softmax needs no
// shuffles: each thread reads its whole score row out of TMEM.
//
//   forward :  S = Q K^T          (UMMA 128 x Lk16 x 64, A/B K-major)          -> TMEM cols [0,128)
//              P = dropout(exp(scale*S - max)) as bf16 into swizzled smem       (row max / sum in registers)
//              O = P V            (UMMA 128 x 64 x Lk, A = P K-major, B = V MN-major) -> TMEM cols [128,192)
//              out = O / sum, lse = max + log(sum)
//   backward:  S = Q K^T, dP = dO V^T  -> TMEM;  P = exp(scale*S - lse), D = sum_j P dP',
//              dS = P (dP' - D), Pd = dropout(P) to smem;  then three UMMAs that reuse the SAME smem tiles under
//              different descriptors:  dV = Pd^T dO (A = Pd MN-major, B = dO MN-major),
//              dQ = dS K (A = dS K-major, B = K MN-major),  dK = dS^T Q (A = dS MN-major, B = Q MN-major).
// Q/K/V/dO tiles arrive by TMA (box 64 x 128 rows, SWIZZLE_128B); rows past this batch element's length
// belong to the next element (finite) or are zero-filled and are neutralised by zeroing the matching P / dS entries.
#pragma once
#include "attention.cuh"
#include "gemm_tc.cuh"
#include <math_constants.h>

namespace atc {
using namespace tc;

constexpr int TILE = 16384;          // 128 rows x 128 B

// copies n mask bytes global -> shared with 16-byte loads; returns the shared pointer that mirrors `mg`
// (offset so that both sides share the same 16-byte misalignment)
__device__ __forceinline__ const uint8_t* load_mask(const uint8_t* __restrict__ mg, int n, uint8_t* mask_s, int t, int nt) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(mg) & 15);
    uint8_t* ms = mask_s + mis;
    const int head = min(n, (16 - mis) & 15);
    if (t < head) ms[t] = mg[t];
    const int nvec = (n - head) >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(mg + head);
    uint4* dst = reinterpret_cast<uint4*>(ms + head);
    for (int i = t; i < nvec; i += nt) dst[i] = src[i];
    for (int i = head + nvec * 16 + t; i < n; i += nt) ms[i] = mg[i];
    return ms;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of element (row, col) inside a [2][128][64] bf16 tile pair with the 128B swizzle (col in [0,128))
__device__ __forceinline__ uint32_t swz_off(int row, int col) {
    const int blk = col >> 6, c = col & 63;
    return (uint32_t)(blk * TILE + row * 128 + ((((c >> 3) ^ (row & 7)) << 4) | ((c & 7) << 1)));
}
// stores 8 consecutive columns [col, col+8) of row `row` (col % 8 == 0) as bf16
__device__ __forceinline__ void st_row8(uint8_t* tile, int row, int col, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(tile + swz_off(row, col)) = u;
}

struct FwdSmem {
    static constexpr int Q = 0, K = TILE, V = 2 * TILE, P = 3 * TILE, MASK = 5 * TILE, RED = MASK + 16384 + 64,
                         BARS = RED + 2 * 4 * 128 * 4;
    static constexpr int TOTAL = BARS + 64 + 1024;
};

// Forward: 256 threads.  Thread (q = warp % 4, half = warp / 4, lane) owns query row q*32+lane (its TMEM lane) and
// the score columns [64*half, 64*half+64); row max / sum are combined across the two halves through shared memory.
__global__ void __launch_bounds__(256, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + FwdSmem::BARS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + FwdSmem::BARS + 32);
    uint8_t* mask_s = sm + FwdSmem::MASK;
    float* red = reinterpret_cast<float*>(sm + FwdSmem::RED);      // [2 kinds][2 halves][128 rows]
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    const int Lq = p.Lq, Lk = p.Lk;
    if (t == 32) {
        for (int i = 0; i < 3; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (t == 0) {
        mbar_expect_tx(bars, 3 * TILE);
        tma_load_2d(base + FwdSmem::Q, &tmQ, h * 64, b * Lq, bars);
        tma_load_2d(base + FwdSmem::K, &tmK, h * 64, b * Lk, bars);
        tma_load_2d(base + FwdSmem::V, &tmV, h * 64, b * Lk, bars);
    }
    const bool dense = p.mask_rstride != 0;
    const uint8_t* mask_l = mask_s;
    if (p.mask) mask_l = load_mask(p.mask + (size_t)b * p.mask_bstride, dense ? Lq * Lk : Lk, mask_s, t, 256);
    __syncthreads();
    mbar_wait(bars, 0);
    const int NS = (Lk + 15) & ~15;
    if (t == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc(128, NS, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, make_smem_desc(base + FwdSmem::Q + k * 32, 16, 1024), make_smem_desc(base + FwdSmem::K + k * 32, 16, 1024),
                      idesc, k > 0 ? 1u : 0u);
        umma_commit(bars + 8);
    }
    mbar_wait(bars + 8, 0);
    __syncwarp();            // tcgen05.ld is .sync.aligned: reconverge after the elected-thread branch / spin loop
    tcgen05_fence_after();
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int nch = (Lk + 31) >> 5;                  // 32-column chunks that hold keys
    const int c0 = half * 2;                         // this thread's chunks: c0, c0+1
    const uint8_t* mrow = p.mask ? (mask_l + (dense ? min(row, Lq - 1) * Lk : 0)) : nullptr;
    float v[64];
    float mx = -INFINITY;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c = c0 + cc;
        if (c < nch) {                               // warp-uniform
            tmem_ld32(trow + c * 32, v + cc * 32);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
                const int j = c * 32 + jj;
                float s = -INFINITY;
                if (j < Lk) {
                    s = v[cc * 32 + jj] * p.scale;
                    if (mrow && mrow[j] == 0) s = -1e9f;
                }
                v[cc * 32 + jj] = s;
                mx = fmaxf(mx, s);
            }
        }
    }
    red[half * 128 + row] = mx;
    __syncthreads();
    mx = fmaxf(red[row], red[128 + row]);
    float sum = 0.f;
    const size_t prow = (((size_t)b * p.H + h) * Lq + row) * Lk;
    const uint32_t drow32 = (uint32_t)((((size_t)b * p.H + h) * Lq + row) * ((Lk + 1) & ~1));    // dropout index base
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c = c0 + cc;
        if (c < nch) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
                const int j = c * 32 + jj;
                float pj = 0.f;
                if (j < Lk) {
                    pj = __expf(v[cc * 32 + jj] - mx);
                    sum += pj;
                }
                v[cc * 32 + jj] = pj;
            }
        }
    }
    red[256 + half * 128 + row] = sum;
    __syncthreads();
    sum = red[256 + row] + red[256 + 128 + row];
    const float inv = 1.f / sum;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c = c0 + cc;
        if (c < nch) {
            if (p.probs && row < Lq) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                    if (c * 32 + jj < Lk) p.probs[prow + c * 32 + jj] = v[cc * 32 + jj] * inv;
            }
            if (p.drop.thresh) {
                const uint32_t pair0 = (drow32 + (uint32_t)(c * 32)) >> 1;      // drow32 is even
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) drop_pair(p.drop, pair0 + jj, v[cc * 32 + 2 * jj], v[cc * 32 + 2 * jj + 1]);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) st_row8(sm + FwdSmem::P, row, c * 32 + g * 8, v + cc * 32 + g * 8);
        }
    }
    if (half == 0 && row < Lq && p.lse) p.lse[((size_t)b * p.H + h) * Lq + row] = mx + __logf(sum);
    fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (t == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc(128, 64, false, true);
        const int nk = (Lk + 15) >> 4;
        for (int k = 0; k < nk; ++k)
            umma_bf16(tmem + 128, make_smem_desc(base + FwdSmem::P + (k >> 2) * TILE + (k & 3) * 32, 16, 1024),
                      make_smem_desc(base + FwdSmem::V + k * 2048, 8192, 1024), idesc, k > 0 ? 1u : 0u);
        umma_commit(bars + 16);
    }
    mbar_wait(bars + 16, 0);
    __syncwarp();
    tcgen05_fence_after();
    {
        // O is 64 columns: each half stores 32 of them
        float o32[32];
        bf16* og = reinterpret_cast<bf16*>(p.O) + ((size_t)b * Lq + row) * p.ldo + h * 64 + half * 32;
        tmem_ld32(trow + 128 + half * 32, o32);
        if (row < Lq) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                f8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] = o32[g * 8 + e] * inv;
                st8(og + g * 8, o);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

struct BwdSmem {
    static constexpr int Q = 0, K = TILE, V = 2 * TILE, DO = 3 * TILE, PD = 4 * TILE, DS = 6 * TILE, MASK = 8 * TILE,
                         RED = MASK + 16384 + 64, BARS = RED + 4 * 128 * 4;
    static constexpr int TOTAL = BARS + 64 + 1024;
};

// Backward: 512 threads.  Thread (q = warp % 4, c = warp / 4, lane) owns query row q*32+lane and the 32 score /
// dP columns of chunk c, which it reads from TMEM exactly once; D is combined across the four chunks through
// shared memory.  dV / dK / dQ reuse the TMEM columns of S / dP once those are dead.
__global__ void __launch_bounds__(512, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, AttnBwdParams bp) {
    const AttnParams& p = bp.f;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + BwdSmem::BARS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + BwdSmem::BARS + 32);
    uint8_t* mask_s = sm + BwdSmem::MASK;
    float* red = reinterpret_cast<float*>(sm + BwdSmem::RED);      // [4 chunks][128 rows]
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, c = warp >> 2, row = q * 32 + lane;
    const int Lq = p.Lq, Lk = p.Lk;
    if (t == 32) {
        for (int i = 0; i < 3; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns: S [0,128)  dP [128,256); after the softmax pass: dV [0,64)  dK [64,128)  dQ [128,192)
    if (t == 0) {
        mbar_expect_tx(bars, 4 * TILE);
        tma_load_2d(base + BwdSmem::Q, &tmQ, h * 64, b * Lq, bars);
        tma_load_2d(base + BwdSmem::K, &tmK, h * 64, b * Lk, bars);
        tma_load_2d(base + BwdSmem::V, &tmV, h * 64, b * Lk, bars);
        tma_load_2d(base + BwdSmem::DO, &tmDO, h * 64, b * Lq, bars);
    }
    const bool dense = p.mask_rstride != 0;
    const uint8_t* mask_l = mask_s;
    if (p.mask) mask_l = load_mask(p.mask + (size_t)b * p.mask_bstride, dense ? Lq * Lk : Lk, mask_s, t, 512);
    __syncthreads();
    mbar_wait(bars, 0);
    const int NS = (Lk + 15) & ~15;
    if (t == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc(128, NS, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, make_smem_desc(base + BwdSmem::Q + k * 32, 16, 1024), make_smem_desc(base + BwdSmem::K + k * 32, 16, 1024),
                      idesc, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + 128, make_smem_desc(base + BwdSmem::DO + k * 32, 16, 1024),
                      make_smem_desc(base + BwdSmem::V + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
        umma_commit(bars + 8);
    }
    mbar_wait(bars + 8, 0);
    __syncwarp();
    tcgen05_fence_after();
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int nch = (Lk + 31) >> 5;
    const bool qok = row < Lq;
    const bool live = c < nch;                        // warp-uniform: this chunk holds keys
    const uint8_t* mrow = p.mask ? (mask_l + (dense ? min(row, Lq - 1) * Lk : 0)) : nullptr;
    const float lse = qok ? p.lse[((size_t)b * p.H + h) * Lq + row] : 0.f;
    const size_t prow = (((size_t)b * p.H + h) * Lq + row) * Lk;
    float s[32], g[32], pd_keep[32];
    const uint32_t drow32 = (uint32_t)((((size_t)b * p.H + h) * Lq + row) * ((Lk + 1) & ~1));    // dropout index base
    float D = 0.f;
    if (live) {
        tmem_ld32(trow + c * 32, s);
        tmem_ld32(trow + 128 + c * 32, g);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const int j = c * 32 + jj;
            float pr = 0.f, gd = 0.f;
            if (qok && j < Lk) {
                float sc = s[jj] * p.scale;
                const bool masked = mrow && mrow[j] == 0;
                if (masked) sc = -1e9f;
                pr = __expf(sc - lse);
                gd = masked ? CUDART_INF_F : g[jj];   // INF marks "no gradient through a masked score" (pr is 0 there anyway)
            }
            s[jj] = pr;
            g[jj] = gd;
        }
        // dropout: one hash per element pair, the same mask for the probability (Pd) and for its gradient
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) pd_keep[jj] = s[jj];
        if (p.drop.thresh) {
            const uint32_t pair0 = (drow32 + (uint32_t)(c * 32)) >> 1;
            const uint32_t t16 = p.drop.thresh >> 16;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const uint32_t hsh = mix32(p.drop.seed + (pair0 + jj) * 0x9e3779b9U);
                const float k0 = ((hsh & 0xffffU) < t16) ? 0.f : p.drop.scale, k1 = ((hsh >> 16) < t16) ? 0.f : p.drop.scale;
                pd_keep[2 * jj] *= k0; pd_keep[2 * jj + 1] *= k1;
                if (g[2 * jj] != CUDART_INF_F) g[2 * jj] *= k0;
                if (g[2 * jj + 1] != CUDART_INF_F) g[2 * jj + 1] *= k1;
            }
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            if (g[jj] != CUDART_INF_F) D += s[jj] * g[jj];
        }
    }
    red[c * 128 + row] = D;
    tcgen05_fence_before();
    __syncthreads();
    D = red[row] + red[128 + row] + red[256 + row] + red[384 + row];
    if (live) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const int j = c * 32 + jj;
            const float pr = s[jj];
            float ds = 0.f, pd = 0.f;
            if (qok && j < Lk) {
                pd = pd_keep[jj];
                ds = (g[jj] == CUDART_INF_F) ? 0.f : pr * (g[jj] - D) * p.scale;   // 1/sqrt(dk) of dQ / dK folded in
            }
            s[jj] = pd;
            g[jj] = ds;
        }
    } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) { s[jj] = 0.f; g[jj] = 0.f; }
    }
    // every chunk (also the key-less ones) is written: the MN-major reads below touch all 128 key columns
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        st_row8(sm + BwdSmem::PD, row, c * 32 + q4 * 8, s + q4 * 8);
        st_row8(sm + BwdSmem::DS, row, c * 32 + q4 * 8, g + q4 * 8);
    }
    fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (t == 0) {
        tcgen05_fence_after();
        // dV[key, dk] = Pd^T dO : M = keys (MN-major A, two 64-key blocks TILE apart), K = query rows, N = dk
        const uint32_t idT = make_idesc(128, 64, true, true);
#pragma unroll
        for (int k = 0; k < 8; ++k)       // 16 query rows per MMA
            umma_bf16(tmem, make_smem_desc(base + BwdSmem::PD + k * 2048, TILE, 1024),
                      make_smem_desc(base + BwdSmem::DO + k * 2048, 8192, 1024), idT, k > 0 ? 1u : 0u);
        // dK[key, dk] = dS^T Q
#pragma unroll
        for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + 64, make_smem_desc(base + BwdSmem::DS + k * 2048, TILE, 1024),
                      make_smem_desc(base + BwdSmem::Q + k * 2048, 8192, 1024), idT, k > 0 ? 1u : 0u);
        // dQ[q, dk] = dS K : A = dS K-major over keys, B = K MN-major
        const uint32_t idQ = make_idesc(128, 64, false, true);
        const int nk = (Lk + 15) >> 4;
        for (int k = 0; k < nk; ++k)
            umma_bf16(tmem + 128, make_smem_desc(base + BwdSmem::DS + (k >> 2) * TILE + (k & 3) * 32, 16, 1024),
                      make_smem_desc(base + BwdSmem::K + k * 2048, 8192, 1024), idQ, k > 0 ? 1u : 0u);
        umma_commit(bars + 16);
    }
    mbar_wait(bars + 16, 0);
    __syncwarp();
    tcgen05_fence_after();
    // six 32-column result chunks over the 16 warps: chunk id = c (dV lo, dV hi, dK lo, dK hi), then c < 2: dQ lo / hi
    auto store_chunk = [&](uint32_t col, void* dst, int ld, int L, int coff) {
        float v[32];
        tmem_ld32(trow + col, v);
        if (row < L) {
            bf16* og = reinterpret_cast<bf16*>(dst) + ((size_t)b * L + row) * ld + h * 64 + coff;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                f8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] = v[g4 * 8 + e];
                st8(og + g4 * 8, o);
            }
        }
    };
    if (c < 2) {
        store_chunk(c * 32, bp.dV, bp.lddv, Lk, c * 32);
        store_chunk(128 + c * 32, bp.dQ, bp.lddq, Lq, c * 32);
    } else {
        store_chunk(64 + (c - 2) * 32, bp.dK, bp.lddk, Lk, (c - 2) * 32);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static int operand_map(const void* ptr, int ld, int rows, int H, CUtensorMap* out) {
    return get_tensor_map(ptr, (uint64_t)H * 64, (uint64_t)rows, (uint64_t)ld * 2, 64, 128, out);
}

static bool supported(const AttnParams& p) {
    return p.Lq <= 128 && p.Lk <= 128 && p.Lk >= 1 && (p.ldq % 8) == 0 && (p.ldk % 8) == 0 && (p.ldv % 8) == 0 &&
           (!p.mask || p.mask_rstride == 0 || p.Lq * p.Lk <= 16384);
}

static int launch_fwd(const AttnParams& p, cudaStream_t st) {
    CUtensorMap tq, tk, tv;
    GCT_TRY(operand_map(p.Q, p.ldq, p.B * p.Lq, p.H, &tq));
    GCT_TRY(operand_map(p.K, p.ldk, p.B * p.Lk, p.H, &tk));
    GCT_TRY(operand_map(p.V, p.ldv, p.B * p.Lk, p.H, &tv));
    static bool attr = false;
    if (!attr) {
        GCT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
        attr = true;
    }
    attn_fwd_tc_kernel<<<p.B * p.H, 256, FwdSmem::TOTAL, st>>>(tq, tk, tv, p);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

static int launch_bwd(const AttnBwdParams& bp, cudaStream_t st) {
    const AttnParams& p = bp.f;
    CUtensorMap tq, tk, tv, tdo;
    GCT_TRY(operand_map(p.Q, p.ldq, p.B * p.Lq, p.H, &tq));
    GCT_TRY(operand_map(p.K, p.ldk, p.B * p.Lk, p.H, &tk));
    GCT_TRY(operand_map(p.V, p.ldv, p.B * p.Lk, p.H, &tv));
    GCT_TRY(operand_map(bp.dO, bp.lddo, p.B * p.Lq, p.H, &tdo));
    static bool attr = false;
    if (!attr) {
        GCT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
        attr = true;
    }
    attn_bwd_tc_kernel<<<p.B * p.H, 512, BwdSmem::TOTAL, st>>>(tq, tk, tv, tdo, bp);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

}  // namespace atc
