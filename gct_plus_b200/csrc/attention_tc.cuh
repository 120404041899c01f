// tcgen05 attention for the training / teacher-forced path: L <= 128, d_k = 64, bf16 operands.
// One CTA per (batch, head), 256 threads: thread (q = warp % 4, half = warp / 4, lane) owns query row q*32+lane
// (= its TMEM lane) and every second 16-column chunk of the scores, which it reads from TMEM -- twice in the
// forward (row max, then exp / sum / dropout), once in the backward -- so the register footprint stays small and
// several CTAs are co-resident per SM (4 forward, 2 backward): one CTA's TMA / MMA latency hides behind another's math.
//
//   forward :  S = Q K^T          (UMMA 128 x NS x 64, A/B K-major)             -> TMEM cols [0,NS), NS = ceil16(Lk)
//              P = dropout(exp2(c*S - max)) as bf16 into swizzled smem (over the dead Q/K tiles)
//              O = P V            (UMMA 128 x 64 x NS, A = P K-major, B = V MN-major) -> TMEM cols [0,64) (S is dead)
//              out = O / sum, lse = max + log(sum)
//   backward:  S = Q K^T, dP = dO V^T -> TMEM;  D = rowsum(dO * O) (from the saved forward output, identical to
//              sum_j P dP' also under dropout);  P = exp(scale*S - lse), dS = P (dP' - D) scale, Pd = dropout(P) to smem;
//              then three UMMAs that reuse the SAME smem tiles under different descriptors:
//              dV = Pd^T dO (A = Pd MN-major, B = dO MN-major), dQ = dS K (A = dS K-major, B = K MN-major),
//              dK = dS^T Q (A = dS MN-major, B = Q MN-major).
// Q/K/V/dO tiles arrive by TMA (box 64 x ceil16(L) rows, SWIZZLE_128B); rows past this batch element's length belong to
// the next element (finite) or are zero-filled and are neutralised by zeroing the matching P / dS entries.  UMMA reads
// of M = 128 rows from a shorter tile run into the neighbouring tile (finite bits, rows never stored).
// The byte mask [B,(1|Lq),Lk] is packed to one bit per key while the TMA loads are in flight.
#pragma once
#include "attention.cuh"
#include "gemm_tc.cuh"
#include <math_constants.h>

extern int g_gct_attn_box;
extern int g_gct_attn_persist;                    // persistent, operand-prefetching kernels (gct_set_attention_persistent): bit 0 forward, bit 1 backward
extern unsigned long long* g_gct_attn_trace;      // debugging hook (gct_set_attention_trace): per-CTA phase timestamps

namespace atc {
using namespace tc;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kFill2 = -1.4426950408889634e9f;     // the reference's -1e9 fill, in log2 units

// phase i of tile `slot` (default: this CTA's index): [16*slot + i] = globaltimer (ns), [16*blockIdx + 8 + i] = SM clock; slot 7 of the first group = %smid
__device__ __forceinline__ void trace_mark(unsigned long long* trace, int i, int slot = -1) {
    if (trace != nullptr && threadIdx.x == 0) {
        if (slot < 0) slot = blockIdx.x;
        unsigned long long g, c;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
        asm volatile("mov.u64 %0, %clock64;" : "=l"(c));
        trace[(size_t)slot * 16 + i] = g;
        trace[(size_t)slot * 16 + 8 + i] = c;
        if (i == 0) {
            uint32_t sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            trace[(size_t)slot * 16 + 7] = sm;
        }
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of 16-byte chunk `chunk` (8 bf16 columns) of row `row` inside one [rows][64] bf16 block, 128B swizzle
__device__ __forceinline__ uint32_t swz16(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// bits[r*4 + w]: bit (j % 32) of word j / 32 is set iff key j is visible to query row r.  nrows = 1 for a broadcast
// key-padding row.  Loads are issued eight rows at a time so that their latencies overlap.
__device__ __forceinline__ void build_mask_bits(const uint8_t* __restrict__ mg, int nrows, int Lk, int rstride, uint32_t* bits,
                                                int warp, int lane, int nwarp) {
    const int nw = (Lk + 31) >> 5;
    for (int r0 = warp * 8; r0 < nrows; r0 += nwarp * 8) {
        for (int w = 0; w < nw; ++w) {
            const int j = w * 32 + lane;
            uint8_t m[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = (j < Lk && r0 + i < nrows) ? mg[(size_t)(r0 + i) * rstride + j] : (uint8_t)0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t word = __ballot_sync(0xffffffffu, m[i] != 0);
                if (lane == 0 && r0 + i < nrows) bits[(r0 + i) * 4 + w] = word;
            }
        }
    }
}
// the 16 mask bits of 16-column chunk c16
__device__ __forceinline__ uint32_t mask16(const uint4& mb, int c16) {
    const uint32_t w = (c16 < 4) ? ((c16 < 2) ? mb.x : mb.y) : ((c16 < 6) ? mb.z : mb.w);
    return w >> ((c16 & 1) << 4);
}
// raw scores -> masked scores in log2 units: visible -> c*s, masked -> fill, past the last key (TAIL chunk only) -> -inf
template <bool TAIL>
__device__ __forceinline__ void score16(float* v, uint32_t m16, float cs, int nvalid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 s = __fmul2_rn(make_float2(v[2 * i], v[2 * i + 1]), make_float2(cs, cs));
        v[2 * i] = (m16 & (1u << (2 * i))) ? s.x : kFill2;
        v[2 * i + 1] = (m16 & (2u << (2 * i))) ? s.y : kFill2;
        if (TAIL) {
            if (2 * i >= nvalid) v[2 * i] = -CUDART_INF_F;
            if (2 * i + 1 >= nvalid) v[2 * i + 1] = -CUDART_INF_F;
        }
    }
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}


// ---- the softmax work on one 16-column chunk of a query row, shared by the one-tile-per-CTA and the persistent kernels ----
// A chunk is classified per WARP (votes over its 32 rows) so that the branch is uniform around the warp-wide tcgen05.ld:
//   every valid column masked for every row (padded keys, the upper triangle of the causal mask): nothing to load -- the
//     probabilities are exactly 0 (exp2(fill - max) underflows) as long as the row has one visible key somewhere;
//   all 16 columns visible for every row: no mask / tail selects, c*s - max as one FMA;
//   otherwise the general path (the reference's masked_fill(-1e9) semantics, tail columns at -inf).
// [B200] the selects were ~6 of the ~27 SASS instructions per score of the backward's pass.
__device__ __forceinline__ uint32_t valid16(int nvalid) { return nvalid >= 16 ? 0xffffu : ((1u << nvalid) - 1u); }

// forward pass 1: running row maximum (mx in scaled log2 units; mx_raw over unscaled scores of all-visible chunks)
__device__ __forceinline__ void fwd_max_chunk(uint32_t taddr, uint32_t m16, int nvalid, float cs, float& mx, float& mx_raw) {
    const uint32_t vis = m16 & valid16(nvalid);
    if (__all_sync(0xffffffffu, vis == 0u)) {
        mx = fmaxf(mx, kFill2);
        return;
    }
    float v[16];
    tmem_ld16(taddr, v);
    if (__all_sync(0xffffffffu, vis == 0xffffu)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mx_raw = fmaxf(mx_raw, v[i]);
        return;
    }
    if (nvalid >= 16) score16<false>(v, m16, cs, 16);
    else score16<true>(v, m16, cs, nvalid);
#pragma unroll
    for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
}

// forward pass 2: pk[8] = dropout(exp2(c*s - mx)) as packed bf16 pairs, sum2 += the pre-dropout probabilities (even | odd columns)
__device__ __forceinline__ void fwd_exp_chunk(uint32_t taddr, uint32_t m16, int nvalid, float cs, float mx, const DropCtx& drop,
                                              uint32_t pair0, float2& sum2, uint32_t* pk) {
    const uint32_t vis = m16 & valid16(nvalid);
    if (__all_sync(0xffffffffu, vis == 0u && mx > 0.5f * kFill2)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = 0u;
        return;
    }
    float v[16];
    tmem_ld16(taddr, v);
    const bool all = __all_sync(0xffffffffu, vis == 0xffffu);
    if (!all) {
        if (nvalid >= 16) score16<false>(v, m16, cs, 16);
        else score16<true>(v, m16, cs, nvalid);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 sv = make_float2(v[2 * i], v[2 * i + 1]);
        const float2 d = all ? __ffma2_rn(sv, make_float2(cs, cs), make_float2(-mx, -mx)) : __fadd2_rn(sv, make_float2(-mx, -mx));
        float2 e = make_float2(ex2_approx(d.x), ex2_approx(d.y));
        sum2 = __fadd2_rn(sum2, e);
        e = __fmul2_rn(e, drop_mult_pair<true>(drop, pair0 + i));
        pk[i] = pack_bf16(e.x, e.y);
    }
}

// backward: P = exp2(c*s - lse), Pd = dropout(P) -> pk[8]; dS = P (dropout'(dP) - D) scale -> dk[8]; row sums of both
__device__ __forceinline__ void bwd_chunk(uint32_t ts, uint32_t tg, uint32_t m16, int nvalid, float cs, float lse2, float D, float scale,
                                          const DropCtx& drop, uint32_t pair0, bool qok, float log2_lk, uint32_t* pk, uint32_t* dk,
                                          float2& rs_p, float2& rs_s) {
    const uint32_t vis = m16 & valid16(nvalid);
    if (__all_sync(0xffffffffu, vis == 0u && lse2 > 0.5f * kFill2)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = dk[i] = 0u;
        return;
    }
    float s[16], g[16];
    tmem_ld16(ts, s);
    tmem_ld16(tg, g);
    const bool all = __all_sync(0xffffffffu, vis == 0xffffu);
    if (!qok) {          // rows past the sequence (this thread only): zero operand rows for the MMAs that follow
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = dk[i] = 0u;
        return;
    }
    if (!all) {
        if (nvalid >= 16) score16<false>(s, m16, cs, 16);
        else score16<true>(s, m16, cs, nvalid);
        if (lse2 <= 0.5f * kFill2) {
            // a row without one visible key (this thread only, practically never): every score is the fill value, softmax is
            // uniform over the Lk keys -- and lse = fill + log(Lk) has lost the log(Lk) to fp32 rounding.  Re-base the row:
            // scores 0 (tail columns stay at -inf), lse = log2(Lk).  Done ahead of the loop so that the loop body, which
            // decides the kernel's registers, is the same for every row.
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] -= kFill2;
            lse2 = log2_lk;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 sv = make_float2(s[2 * i], s[2 * i + 1]);
        const float2 d = all ? __ffma2_rn(sv, make_float2(cs, cs), make_float2(-lse2, -lse2)) : __fadd2_rn(sv, make_float2(-lse2, -lse2));
        const float2 pr = make_float2(ex2_approx(d.x), ex2_approx(d.y));
        const float2 keep = drop_mult_pair<true>(drop, pair0 + i);
        const float2 pd = __fmul2_rn(pr, keep);
        const float2 gd = __ffma2_rn(make_float2(g[2 * i], g[2 * i + 1]), keep, make_float2(-D, -D));
        float2 ds = __fmul2_rn(__fmul2_rn(pr, make_float2(scale, scale)), gd);     // 1/sqrt(dk) of dQ / dK folded in
        if (!all) {
            // no gradient through a masked score (the reference's masked_fill); a padded column has pr = 0 already
            if (!(m16 & (1u << (2 * i)))) ds.x = 0.f;
            if (!(m16 & (2u << (2 * i)))) ds.y = 0.f;
        }
        pk[i] = pack_bf16(pd.x, pd.y);
        dk[i] = pack_bf16(ds.x, ds.y);
        rs_p = __fadd2_rn(rs_p, pd);
        rs_s = __fadd2_rn(rs_s, ds);
    }
}

struct FwdLayout {
    int q, k, v, bits, red, bars, total;       // byte offsets from the 1 KB-aligned base; P aliases [0, 32768)
    __host__ __device__ FwdLayout(int RPq, int RPk) {
        q = 0; k = RPq * 128;
        int qk = k + RPk * 128;
        if (qk < 32768) qk = 32768;
        v = qk; bits = v + RPk * 128; red = bits + 2048; bars = red + 2048; total = bars + 64 + 1024;
    }
};

__global__ void __launch_bounds__(256, 4)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, int box_store, AttnParams p,
                   unsigned long long* trace) {
    pdl_wait();        // programmatic dependent launch: q / k / v (dO, O, lse) come from the kernels just before this one
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int Lq = p.Lq, Lk = p.Lk;
    const int RPq = (Lq + 15) & ~15, NS = (Lk + 15) & ~15;
    const FwdLayout L(RPq, NS);
    const uint32_t bars = base + L.bars;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.bars + 32);
    uint32_t* bits = reinterpret_cast<uint32_t*>(sm + L.bits);
    float* red = reinterpret_cast<float*>(sm + L.red);       // [2 kinds][2 halves][128 rows]
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    trace_mark(trace, 0);
    if (t == 32) {
        // the thread that initialises the barriers also starts the loads: they are in flight before the TMEM allocation
        for (int i = 0; i < 3; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bars, (uint32_t)(RPq + 2 * NS) * 128u);
        tma_load_2d(base + L.q, &tmQ, h * 64, b * Lq, bars);
        tma_load_2d(base + L.k, &tmK, h * 64, b * Lk, bars);
        tma_load_2d(base + L.v, &tmV, h * 64, b * Lk, bars);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();       // only now: a dependent that allocates TMEM in its prologue must not get ahead of this CTA's allocation
    trace_mark(trace, 1);
    const bool dense = p.mask_rstride != 0;
    if (p.mask) build_mask_bits(p.mask + (size_t)b * p.mask_bstride, dense ? Lq : 1, Lk, dense ? p.mask_rstride : 0, bits, warp, lane, 8);
    __syncthreads();
    mbar_wait(bars, 0);
    trace_mark(trace, 2);
    if (t == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc(128, NS, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, make_smem_desc(base + L.q + k * 32, 16, 1024), make_smem_desc(base + L.k + k * 32, 16, 1024), idesc,
                      k > 0 ? 1u : 0u);
        umma_commit(bars + 8);
    }
    // this thread's 16-column chunks: half, half + 2, ... (interleaved so that the chunks the mask empties -- padded keys, the
    // causal mask's upper triangle -- are shared between the two column halves)
    const int nchk = NS >> 4;
    const bool wactive = q * 32 < Lq;                 // warp-uniform: the warp owns at least one real query row
    uint4 mb = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (p.mask) mb = *reinterpret_cast<const uint4*>(bits + (dense ? min(row, Lq - 1) : 0) * 4);
    const float cs = p.scale * kLog2e;
    mbar_wait(bars + 8, 0);
    __syncwarp();            // tcgen05.ld is .sync.aligned: reconverge after the elected-thread branch / spin loop
    tcgen05_fence_after();
    trace_mark(trace, 3);
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    float mx = -CUDART_INF_F;
    if (wactive) {
        float mx_raw = -CUDART_INF_F;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            const int c16 = half + 2 * ci;
            if (c16 < nchk) fwd_max_chunk(trow + c16 * 16, mask16(mb, c16), Lk - c16 * 16, cs, mx, mx_raw);
        }
        mx = fmaxf(mx, mx_raw * cs);
    }
    red[half * 128 + row] = mx;
    __syncthreads();
    mx = fmaxf(red[row], red[128 + row]);
    float sum = 0.f;
    const uint32_t drow32 = (uint32_t)((((size_t)b * p.H + h) * Lq + row) * ((Lk + 1) & ~1));    // dropout index base
    if (wactive) {
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            const int c16 = half + 2 * ci;
            if (c16 < nchk) {
                uint32_t pk[8];
                const uint32_t pair0 = (drow32 + (uint32_t)(c16 * 16)) >> 1;      // drow32 is even
                fwd_exp_chunk(trow + c16 * 16, mask16(mb, c16), Lk - c16 * 16, cs, mx, p.drop, pair0, sum2, pk);
                uint8_t* pt = sm + (c16 >> 2) * 16384;
                *reinterpret_cast<uint4*>(pt + swz16(row, (c16 & 3) * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(pt + swz16(row, (c16 & 3) * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
        sum = sum2.x + sum2.y;
    }
    red[256 + half * 128 + row] = sum;
    __syncthreads();
    sum = red[256 + row] + red[256 + 128 + row];
    const float inv = 1.f / sum;
    if (p.probs && wactive) {          // get_attn: normalised pre-dropout probabilities (S is still intact in TMEM)
        const size_t prow = (((size_t)b * p.H + h) * Lq + row) * Lk;
#pragma unroll 1
        for (int c16 = half; c16 < nchk; c16 += 2) {
            float v[16];
            tmem_ld16(trow + c16 * 16, v);
            const int nvalid = Lk - c16 * 16;
            score16<true>(v, mask16(mb, c16), cs, nvalid);
            if (row < Lq) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (i < nvalid) p.probs[prow + c16 * 16 + i] = ex2_approx(v[i] - mx) * inv;
            }
        }
    }
    if (half == 0 && row < Lq && p.lse) p.lse[((size_t)b * p.H + h) * Lq + row] = (mx + __log2f(sum)) * kLn2;
    fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    trace_mark(trace, 4);
    if (t == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc(128, 64, false, true);
        const int nk = NS >> 4;
        for (int k = 0; k < nk; ++k)
            umma_bf16(tmem, make_smem_desc(base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      make_smem_desc(base + L.v + k * 2048, 8192, 1024), idesc, k > 0 ? 1u : 0u);
        umma_commit(bars + 16);
    }
    mbar_wait(bars + 16, 0);
    __syncwarp();
    tcgen05_fence_after();
    trace_mark(trace, 5);
    if (wactive) {
        // O is 64 columns: each half stores 32 of them
        float o32[32];
        tmem_ld32(trow + half * 32, o32);
        if (box_store) {
            // the output tile leaves as ONE TMA box (rows past Lq are outside the rank-3 map): a lane writing its own row to
            // global memory makes every 16-byte store of the warp hit a different line
            if (row < RPq) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(o32[g * 8 + 2 * e] * inv, o32[g * 8 + 2 * e + 1] * inv);
                    *reinterpret_cast<uint4*>(sm + L.q + swz16(row, half * 4 + g)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);    // Q / P are dead
                }
            }
        } else if (row < Lq) {
            bf16* og = reinterpret_cast<bf16*>(p.O) + ((size_t)b * Lq + row) * p.ldo + h * 64 + half * 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                f8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] = o32[g * 8 + e] * inv;
                st8(og + g * 8, o);
            }
        }
    }
    if (box_store) fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (box_store && t == 0) {
        tma_store_3d(&tmO, base + L.q, h * 64, 0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    trace_mark(trace, 6);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

struct BwdLayout {
    // BS = bytes of one [RPq][64] block.  Pd block 0 is written over the V tile (dead once dP = dO V^T has completed),
    // Pd block 1 follows it; LBOP = their distance (the MN-major descriptor's leading-dimension byte offset).
    int BS, LBOP, pd, ds, dO, q, k, v, bits, red, bars, csum, total;
    __host__ __device__ BwdLayout(int RPq, int RPk) {
        BS = RPq * 128;
        const int VB = RPk * 128;
        LBOP = VB > BS ? VB : BS;
        v = 0; pd = 0;
        ds = LBOP + BS; dO = ds + 2 * BS; q = dO + BS; k = q + BS; bits = k + VB;
        red = bits + 2048; bars = red + 1024; csum = bars + 64; total = csum + 768;
        // UMMA A operands always span 128 rows: keep the furthest such read (Q as A of S) inside the allocation
        if (total < q + 16384) total = q + 16384;
        total += 1024;
    }
};

__global__ void __launch_bounds__(256, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQ,
                   const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV, int box_io, AttnBwdParams bp,
                   unsigned long long* trace) {
    pdl_wait();        // programmatic dependent launch: q / k / v (dO, O, lse) come from the kernels just before this one
    const AttnParams& p = bp.f;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int Lq = p.Lq, Lk = p.Lk;
    const int RPq = (Lq + 15) & ~15, NS = (Lk + 15) & ~15;
    const BwdLayout L(RPq, NS);
    const uint32_t bars = base + L.bars;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.bars + 32);
    uint32_t* bits = reinterpret_cast<uint32_t*>(sm + L.bits);
    float* red = reinterpret_cast<float*>(sm + L.red);        // [2 halves][128 rows]
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    float* csum = reinterpret_cast<float*>(sm + L.csum);      // [dV | dK | dQ][64] column sums (bias gradients), see below
    const bool want_bsum = bp.bsum_q != nullptr || bp.bsum_k != nullptr || bp.bsum_v != nullptr;
    if (t < 192) csum[t] = 0.f;
    trace_mark(trace, 0);
    if (t == 32) {
        for (int i = 0; i < 3; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bars, (uint32_t)((box_io ? 3 : 2) * RPq + 2 * NS) * 128u);
        tma_load_2d(base + L.q, &tmQ, h * 64, b * Lq, bars);
        tma_load_2d(base + L.k, &tmK, h * 64, b * Lk, bars);
        tma_load_2d(base + L.v, &tmV, h * 64, b * Lk, bars);
        tma_load_2d(base + L.dO, &tmDO, h * 64, b * Lq, bars);
        if (box_io) tma_load_2d(base + L.ds, &tmO, h * 64, b * Lq, bars);      // forward output tile: parked where dS goes later
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
    pdl_launch_dependents();       // only now: a dependent that allocates TMEM in its prologue must not get ahead of this CTA's allocation
    // TMEM columns: S [0,128)  dP [128,256); after the softmax pass: dV [0,64)  dK [64,128)  dQ [128,192)
    const bool dense = p.mask_rstride != 0;
    if (p.mask) build_mask_bits(p.mask + (size_t)b * p.mask_bstride, dense ? Lq : 1, Lk, dense ? p.mask_rstride : 0, bits, warp, lane, 8);
    const bool qok = row < Lq;
    // forward output row (32 of its 64 columns) for D = rowsum(dO * O): requested before the TMA wait
    uint4 ov[4];
    if (qok && !box_io) {
        const uint4* og = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.O) + ((size_t)b * Lq + row) * p.ldo + h * 64 + half * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) ov[i] = og[i];
    }
    const float lse2 = qok ? p.lse[((size_t)b * p.H + h) * Lq + row] * kLog2e : 0.f;
    trace_mark(trace, 1);
    __syncthreads();
    mbar_wait(bars, 0);
    trace_mark(trace, 2);
    if (t == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc(128, NS, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, make_smem_desc(base + L.q + k * 32, 16, 1024), make_smem_desc(base + L.k + k * 32, 16, 1024), idesc,
                      k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + 128, make_smem_desc(base + L.dO + k * 32, 16, 1024), make_smem_desc(base + L.v + k * 32, 16, 1024),
                      idesc, k > 0 ? 1u : 0u);
        umma_commit(bars + 8);
    }
    float D = 0.f;
    if (qok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 dv = *reinterpret_cast<const uint4*>(sm + L.dO + swz16(row, half * 4 + i));
            if (box_io) ov[i] = *reinterpret_cast<const uint4*>(sm + L.ds + swz16(row, half * 4 + i));
            const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&dv);
            const __nv_bfloat162* o = reinterpret_cast<const __nv_bfloat162*>(&ov[i]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 af = __bfloat1622float2(a[e]), of = __bfloat1622float2(o[e]);
                D = fmaf(af.x, of.x, D);
                D = fmaf(af.y, of.y, D);
            }
        }
    }
    red[half * 128 + row] = D;
    __syncthreads();
    D = red[row] + red[128 + row];
    const int nchk = NS >> 4;                         // this thread's 16-column chunks: half, half + 2, ...
    const bool wlive = q * 32 < RPq;                  // warp-uniform: rows the MN-major (query-row K dimension) reads touch
    uint4 mb = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (p.mask) mb = *reinterpret_cast<const uint4*>(bits + (dense ? min(row, Lq - 1) : 0) * 4);
    const float cs = p.scale * kLog2e;
    const uint32_t drow32 = (uint32_t)((((size_t)b * p.H + h) * Lq + row) * ((Lk + 1) & ~1));    // dropout index base
    mbar_wait(bars + 8, 0);
    __syncwarp();
    tcgen05_fence_after();
    trace_mark(trace, 3);
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    if (wlive) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            const int c16 = half + 2 * ci;
            if (c16 < nchk) {
                uint32_t pk[8], dk[8];
                const uint32_t pair0 = (drow32 + (uint32_t)(c16 * 16)) >> 1;
                float2 rs_p = make_float2(0.f, 0.f), rs_s = rs_p;       // row sums: unused here
                bwd_chunk(trow + c16 * 16, trow + 128 + c16 * 16, mask16(mb, c16), Lk - c16 * 16, cs, lse2, D, p.scale, p.drop, pair0, qok, log2f((float)Lk), pk,
                          dk, rs_p, rs_s);
                const uint32_t blk = (uint32_t)(c16 >> 2) * (uint32_t)L.BS, blkp = (uint32_t)(c16 >> 2) * (uint32_t)L.LBOP;
                const int ch = (c16 & 3) * 2;
                if (row < RPq) {      // a block holds RPq rows: rows past it would land in the next block / tile
                    *reinterpret_cast<uint4*>(sm + L.pd + blkp + swz16(row, ch)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(sm + L.pd + blkp + swz16(row, ch + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    *reinterpret_cast<uint4*>(sm + L.ds + blk + swz16(row, ch)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
                    *reinterpret_cast<uint4*>(sm + L.ds + blk + swz16(row, ch + 1)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
                }
            }
        }
    }
    fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    trace_mark(trace, 4);
    if (t == 0) {
        tcgen05_fence_after();
        const int nq = RPq >> 4, nk = NS >> 4;
        // dV[key, dk] = Pd^T dO : M = keys (MN-major A, two 64-key blocks BS apart), K = query rows, N = dk
        const uint32_t idT = make_idesc(128, 64, true, true);
        for (int k = 0; k < nq; ++k)          // 16 query rows per MMA
            umma_bf16(tmem, make_smem_desc(base + L.pd + k * 2048, (uint32_t)L.LBOP, 1024),
                      make_smem_desc(base + L.dO + k * 2048, 8192, 1024), idT, k > 0 ? 1u : 0u);
        // dK[key, dk] = dS^T Q
        for (int k = 0; k < nq; ++k)
            umma_bf16(tmem + 64, make_smem_desc(base + L.ds + k * 2048, (uint32_t)L.BS, 1024),
                      make_smem_desc(base + L.q + k * 2048, 8192, 1024), idT, k > 0 ? 1u : 0u);
        // dQ[q, dk] = dS K : A = dS K-major over keys, B = K MN-major
        const uint32_t idQ = make_idesc(128, 64, false, true);
        for (int k = 0; k < nk; ++k)
            umma_bf16(tmem + 128, make_smem_desc(base + L.ds + (k >> 2) * L.BS + (k & 3) * 32, 16, 1024),
                      make_smem_desc(base + L.k + k * 2048, 8192, 1024), idQ, k > 0 ? 1u : 0u);
        umma_commit(bars + 16);
    }
    mbar_wait(bars + 16, 0);
    __syncwarp();
    tcgen05_fence_after();
    trace_mark(trace, 5);
    // each thread: 32 of the 64 columns of its dV / dK row (row = key) and of its dQ row (row = query)
    // `which` 0/1/2 = dV/dK/dQ.  With bias-gradient outputs requested the warp also reduces its 32 rows column-wise (butterfly
    // with halving: 31 shuffles leave lane j with the sum of column half*32 + j) into the CTA's csum slab.
    auto store_chunk = [&](uint32_t col, void* dst, int ld, int Lr, int which) {
        float v[32];
        tmem_ld32(trow + col + half * 32, v);
        if (box_io) {
            // park the tile in shared memory (dV over K, dK over V / Pd, dQ over Q: all dead once the last MMAs have completed);
            // it leaves as one TMA box per tensor below
            const int RPr = (Lr + 15) & ~15;
            uint8_t* tile = sm + (which == 0 ? L.k : (which == 1 ? L.v : L.q));
            if (row < RPr) {
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(v[g4 * 8 + 2 * e], v[g4 * 8 + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(tile + swz16(row, half * 4 + g4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        } else if (row < Lr) {
            bf16* og = reinterpret_cast<bf16*>(dst) + ((size_t)b * Lr + row) * ld + h * 64 + half * 32;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                f8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] = v[g4 * 8 + e];
                st8(og + g4 * 8, o);
            }
        }
        if (want_bsum) {
            if (row >= Lr) {        // rows past the sequence: not part of the tensor (UMMA's M = 128 reads ran into a neighbour tile)
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0.f;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const float send = up ? v[i] : v[i + off];
                    const float keep = up ? v[i + off] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            atomicAdd(&csum[which * 64 + half * 32 + lane], v[0]);
        }
    };
    if (q * 32 < Lk) {
        store_chunk(0, bp.dV, bp.lddv, Lk, 0);
        store_chunk(64, bp.dK, bp.lddk, Lk, 1);
    }
    if (q * 32 < Lq) store_chunk(128, bp.dQ, bp.lddq, Lq, 2);
    if (box_io) fence_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (box_io && t == 0) {
        tma_store_3d(&tmDV, base + L.k, h * 64, 0, b);
        tma_store_3d(&tmDK, base + L.v, h * 64, 0, b);
        tma_store_3d(&tmDQ, base + L.q, h * 64, 0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    trace_mark(trace, 6);
    if (want_bsum && t < 192) {
        float* dst = (t < 64) ? bp.bsum_v : (t < 128) ? bp.bsum_k : bp.bsum_q;
        if (dst) atomicAdd(dst + h * 64 + (t & 63), csum[t]);
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------------
// Persistent variants for L <= 96 (both sequence lengths), the shapes the training step runs (src 81, trg 80).
// [B200] per-CTA phase timestamps of the one-tile-per-CTA kernels above (scripts/trace_attention.py,
// profiles/r02_attention_trace.txt) show ~2.7 us of each CTA's ~10 us (backward) / ~6.7 us (forward) life spent before its
// operand tiles have landed, plus ~1 us between one CTA's exit and its successor's start.  Here a CTA keeps its TMEM
// allocation and walks a contiguous range of (batch, head) tiles; the next tile's operands are requested while the current
// tile is still computing, into shared-memory units whose contents are dead by then:
//   backward (2 CTAs / SM, nine units of U = 128 * ceil16(L) bytes): Q and K alternate between two pairs of units (the pair of
//   tile i+1 is loaded during tile i); V, dO and the saved O of tile i+1 are requested as soon as tile i's last MMAs have
//   completed (into the dS / dO units) and land while tile i's results are being staged and stored;
//   forward (3 CTAs / SM, five units): Q / K pairs as above, V of tile i+1 requested when tile i's P V has completed.
// Results leave as TMA boxes from dead units exactly as above; a unit is re-used only after `cp.async.bulk.wait_group.read`.
struct FwdPersistLayout {
    int U, q0, k0, ps, v, bits, red, bars, total;       // pair x: Q at q0 + x*ps, K at k0 + x*ps; P block 0 / 1 of a tile are written over its own Q / K units
    __host__ __device__ FwdPersistLayout(int RP) {
        U = RP * 128;
        q0 = 0; k0 = U; ps = 2 * U; v = 4 * U;
        bits = 5 * U; red = bits + 2048; bars = red + 2048; total = bars + 128;
        if (total < k0 + ps + 16384) total = k0 + ps + 16384;          // UMMA A operands span 128 rows from their base
        total += 1024;
    }
};

__global__ void __launch_bounds__(256, 3)
attn_fwd_tc_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, AttnParams p, int ntiles,
                           unsigned long long* trace) {
    pdl_wait();        // programmatic dependent launch: q / k / v (dO, O, lse) come from the kernels just before this one
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int Lq = p.Lq, Lk = p.Lk;
    const int RPq = (Lq + 15) & ~15, NS = (Lk + 15) & ~15;
    const FwdPersistLayout L(RPq > NS ? RPq : NS);
    const uint32_t bars = base + L.bars;
    const uint32_t bar_qk0 = bars, bar_v = bars + 16, bar_m1 = bars + 24, bar_m2 = bars + 32;     // bar_qk1 = bars + 8
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.bars + 64);
    uint32_t* bits = reinterpret_cast<uint32_t*>(sm + L.bits);
    float* red = reinterpret_cast<float*>(sm + L.red);       // [2 kinds][2 halves][128 rows]
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    const int first = (int)((long long)blockIdx.x * ntiles / gridDim.x), last = (int)((long long)(blockIdx.x + 1) * ntiles / gridDim.x);
    const int n = last - first;
    const uint32_t qk_bytes = (uint32_t)(RPq + NS) * 128u, v_bytes = (uint32_t)NS * 128u;
    if (t == 0 && n > 0) {
        for (int i = 0; i < 5; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int b0 = first / p.H, h0 = first % p.H;
        mbar_expect_tx(bar_qk0, qk_bytes);
        tma_load_2d(base + L.q0, &tmQ, h0 * 64, b0 * Lq, bar_qk0);
        tma_load_2d(base + L.k0, &tmK, h0 * 64, b0 * Lk, bar_qk0);
        mbar_expect_tx(bar_v, v_bytes);
        tma_load_2d(base + L.v, &tmV, h0 * 64, b0 * Lk, bar_v);
        if (n > 1) {
            const int b1 = (first + 1) / p.H, h1 = (first + 1) % p.H;
            mbar_expect_tx(bar_qk0 + 8, qk_bytes);
            tma_load_2d(base + (L.q0 + L.ps), &tmQ, h1 * 64, b1 * Lq, bar_qk0 + 8);
            tma_load_2d(base + (L.k0 + L.ps), &tmK, h1 * 64, b1 * Lk, bar_qk0 + 8);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();       // only now: a dependent that allocates TMEM in its prologue must not get ahead of this CTA's allocation
    const bool dense = p.mask_rstride != 0;
    const int nchk = NS >> 4;                         // this thread's 16-column chunks: half, half + 2, ...
    const bool wactive = q * 32 < Lq;                 // warp-uniform: the warp owns at least one real query row
    const float cs = p.scale * kLog2e;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    int b_prev = -1;
    for (int i = 0; i < n; ++i) {
        const int tile = first + i, b = tile / p.H, h = tile - b * p.H, pr = i & 1;
        const uint32_t ph = (uint32_t)(i & 1), phqk = (uint32_t)((i >> 1) & 1);
        const uint32_t uq = (uint32_t)(L.q0 + pr * L.ps), uk = (uint32_t)(L.k0 + pr * L.ps);
        trace_mark(trace, 0, tile);
        const bool newmask = p.mask != nullptr && b != b_prev;      // CTA-uniform
        b_prev = b;
        if (newmask) build_mask_bits(p.mask + (size_t)b * p.mask_bstride, dense ? Lq : 1, Lk, dense ? p.mask_rstride : 0, bits, warp, lane, 8);
        if (t == 0) {
            mbar_wait(bar_qk0 + 8 * pr, phqk);
            tcgen05_fence_after();
            trace_mark(trace, 2, tile);
            const uint32_t idesc = make_idesc(128, NS, false, false);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, make_smem_desc(base + uq + k * 32, 16, 1024), make_smem_desc(base + uk + k * 32, 16, 1024), idesc,
                          k > 0 ? 1u : 0u);
            umma_commit(bar_m1);
            // tile i-1's output box has left its unit (the other pair's Q unit): that pair is free for tile i+1's Q / K
            if (i > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (i > 0 && i + 1 < n) {
                const int b1 = (tile + 1) / p.H, h1 = (tile + 1) % p.H;
                const uint32_t bq = bar_qk0 + 8 * (pr ^ 1);
                mbar_expect_tx(bq, qk_bytes);
                tma_load_2d(base + (L.q0 + (pr ^ 1) * L.ps), &tmQ, h1 * 64, b1 * Lq, bq);
                tma_load_2d(base + (L.k0 + (pr ^ 1) * L.ps), &tmK, h1 * 64, b1 * Lk, bq);
            }
        }
        if (newmask) __syncthreads();
        uint4 mb = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (p.mask) mb = *reinterpret_cast<const uint4*>(bits + (dense ? min(row, Lq - 1) : 0) * 4);
        mbar_wait(bar_m1, ph);
        __syncwarp();            // tcgen05.ld is .sync.aligned: reconverge after the elected-thread branch / spin loop
        tcgen05_fence_after();
        trace_mark(trace, 3, tile);
        float mx = -CUDART_INF_F;
        if (wactive) {
            float mx_raw = -CUDART_INF_F;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const int c16 = half + 2 * ci;
                if (c16 < nchk) fwd_max_chunk(trow + c16 * 16, mask16(mb, c16), Lk - c16 * 16, cs, mx, mx_raw);
            }
            mx = fmaxf(mx, mx_raw * cs);
        }
        red[half * 128 + row] = mx;
        __syncthreads();
        mx = fmaxf(red[row], red[128 + row]);
        float sum = 0.f;
        const uint32_t drow32 = (uint32_t)((((size_t)b * p.H + h) * Lq + row) * ((Lk + 1) & ~1));    // dropout index base
        if (wactive) {
            float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const int c16 = half + 2 * ci;
                if (c16 < nchk) {
                    uint32_t pk[8];
                    const uint32_t pair0 = (drow32 + (uint32_t)(c16 * 16)) >> 1;      // drow32 is even
                    fwd_exp_chunk(trow + c16 * 16, mask16(mb, c16), Lk - c16 * 16, cs, mx, p.drop, pair0, sum2, pk);
                    if (row < RPq) {       // a unit holds ceil16(L) rows
                        uint8_t* pt = sm + ((c16 >> 2) ? uk : uq);
                        *reinterpret_cast<uint4*>(pt + swz16(row, (c16 & 3) * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(pt + swz16(row, (c16 & 3) * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
            }
            sum = sum2.x + sum2.y;
        }
        red[256 + half * 128 + row] = sum;
        __syncthreads();
        sum = red[256 + row] + red[256 + 128 + row];
        const float inv = 1.f / sum;
        if (p.probs && wactive) {          // get_attn: normalised pre-dropout probabilities (S is still intact in TMEM)
            const size_t prow = (((size_t)b * p.H + h) * Lq + row) * Lk;
#pragma unroll 1
            for (int c16 = half; c16 < nchk; c16 += 2) {
                float v[16];
                tmem_ld16(trow + c16 * 16, v);
                const int nvalid = Lk - c16 * 16;
                score16<true>(v, mask16(mb, c16), cs, nvalid);
                if (row < Lq) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (e < nvalid) p.probs[prow + c16 * 16 + e] = ex2_approx(v[e] - mx) * inv;
                }
            }
        }
        if (half == 0 && row < Lq && p.lse) p.lse[((size_t)b * p.H + h) * Lq + row] = (mx + __log2f(sum)) * kLn2;
        fence_async_smem();
        tcgen05_fence_before();
        __syncthreads();
        trace_mark(trace, 4, tile);
        if (t == 0) {
            mbar_wait(bar_v, ph);
            tcgen05_fence_after();
            const uint32_t idesc = make_idesc(128, 64, false, true);
            const int nk = NS >> 4;
            for (int k = 0; k < nk; ++k)
                umma_bf16(tmem, make_smem_desc(base + ((k >> 2) ? uk : uq) + (k & 3) * 32, 16, 1024),
                          make_smem_desc(base + L.v + k * 2048, 8192, 1024), idesc, k > 0 ? 1u : 0u);
            umma_commit(bar_m2);
        }
        mbar_wait(bar_m2, ph);
        __syncwarp();
        tcgen05_fence_after();
        trace_mark(trace, 5, tile);
        if (t == 0 && i + 1 < n) {         // V is dead: request the next tile's
            const int b1 = (tile + 1) / p.H, h1 = (tile + 1) % p.H;
            mbar_expect_tx(bar_v, v_bytes);
            tma_load_2d(base + L.v, &tmV, h1 * 64, b1 * Lk, bar_v);
        }
        if (wactive) {
            float o32[32];
            tmem_ld32(trow + half * 32, o32);
            if (row < RPq) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(o32[g * 8 + 2 * e] * inv, o32[g * 8 + 2 * e + 1] * inv);
                    *reinterpret_cast<uint4*>(sm + uq + swz16(row, half * 4 + g)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);    // Q / P are dead
                }
            }
        }
        fence_async_smem();
        tcgen05_fence_before();
        __syncthreads();
        tcgen05_fence_after();
        if (t == 0) {
            tma_store_3d(&tmO, base + uq, h * 64, 0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        trace_mark(trace, 6, tile);
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

struct BwdPersistLayout {
    // units of U bytes: dS block 0 / 1 (the next tile's saved O / V are parked there until the softmax pass overwrites them), dO,
    // the (Q, K) units, Pd block 0 / 1.  Staging of the results: dV over this tile's K, dK over Pd block 0, dQ over this tile's Q.
    // L <= 96: two (Q, K) pairs (nine units; the next tile's pair is loaded a whole tile ahead); 96 < L <= 112: nine units no
    // longer fit twice per SM, so one pair (seven units), reloaded as soon as this tile's dQ / dV boxes have left it.
    int U, ds, dO, q0, k0, ps, pd, bits, red, csum, bars, total;      // pair x: Q at q0 + x*ps, K at k0 + x*ps (ps = 0: one pair)
    __host__ __device__ BwdPersistLayout(int RP) {
        U = RP * 128;
        ds = 0; dO = 2 * U;
        if (RP <= 96) { q0 = 3 * U; k0 = 5 * U; ps = U; pd = 7 * U; bits = 9 * U; }
        else          { q0 = 3 * U; k0 = 4 * U; ps = 0; pd = 5 * U; bits = 7 * U; }
        red = bits + 2048; csum = red + 1024; bars = csum + 768; total = bars + 128;
        if (total < q0 + ps + 16384) total = q0 + ps + 16384;          // UMMA A operands span 128 rows from their base
        total += 1024;
    }
};

__global__ void __launch_bounds__(256, 2)
attn_bwd_tc_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                           const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQ,
                           const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV, AttnBwdParams bp, int ntiles,
                           unsigned long long* trace) {
    pdl_wait();        // programmatic dependent launch: q / k / v (dO, O, lse) come from the kernels just before this one
    const AttnParams& p = bp.f;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int Lq = p.Lq, Lk = p.Lk;
    const int RPq = (Lq + 15) & ~15, NS = (Lk + 15) & ~15;
    const BwdPersistLayout L(RPq > NS ? RPq : NS);
    const uint32_t U = (uint32_t)L.U;
    const uint32_t bars = base + L.bars;
    const uint32_t bar_qk0 = bars, bar_in = bars + 16, bar_m1 = bars + 24, bar_m2 = bars + 32;      // bar_qk1 = bars + 8
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.bars + 64);
    uint32_t* bits = reinterpret_cast<uint32_t*>(sm + L.bits);
    float* red = reinterpret_cast<float*>(sm + L.red);        // [2 halves][128 rows]
    float* csum = reinterpret_cast<float*>(sm + L.csum);      // [dV | dK | dQ][64] column sums (bias gradients)
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    const bool want_bsum = bp.bsum_q != nullptr || bp.bsum_k != nullptr || bp.bsum_v != nullptr;
    const int first = (int)((long long)blockIdx.x * ntiles / gridDim.x), last = (int)((long long)(blockIdx.x + 1) * ntiles / gridDim.x);
    const int n = last - first;
    const uint32_t qk_bytes = (uint32_t)(RPq + NS) * 128u, in_bytes = (uint32_t)(NS + 2 * RPq) * 128u;
    const bool dbl = L.ps != 0;                       // two (Q, K) pairs
    if (t < 192) csum[t] = 0.f;
    if (t == 0 && n > 0) {
        for (int i = 0; i < 5; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int b0 = first / p.H, h0 = first % p.H;
        mbar_expect_tx(bar_qk0, qk_bytes);
        tma_load_2d(base + L.q0, &tmQ, h0 * 64, b0 * Lq, bar_qk0);
        tma_load_2d(base + L.k0, &tmK, h0 * 64, b0 * Lk, bar_qk0);
        mbar_expect_tx(bar_in, in_bytes);
        tma_load_2d(base + L.ds + U, &tmV, h0 * 64, b0 * Lk, bar_in);
        tma_load_2d(base + L.dO, &tmDO, h0 * 64, b0 * Lq, bar_in);
        tma_load_2d(base + L.ds, &tmO, h0 * 64, b0 * Lq, bar_in);
        if (dbl && n > 1) {
            const int b1 = (first + 1) / p.H, h1 = (first + 1) % p.H;
            mbar_expect_tx(bar_qk0 + 8, qk_bytes);
            tma_load_2d(base + (L.q0 + L.ps), &tmQ, h1 * 64, b1 * Lq, bar_qk0 + 8);
            tma_load_2d(base + (L.k0 + L.ps), &tmK, h1 * 64, b1 * Lk, bar_qk0 + 8);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const bool qok = row < Lq;
    float lse_next = (qok && n > 0) ? p.lse[(size_t)first * Lq + row] : 0.f;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();       // only now: a dependent that allocates TMEM in its prologue must not get ahead of this CTA's allocation
    // TMEM columns: S [0,128)  dP [128,256); after the softmax pass: dV [0,64)  dK [64,128)  dQ [128,192)
    const bool dense = p.mask_rstride != 0;
    const int nchk = NS >> 4;                         // this thread's 16-column chunks: half, half + 2, ...
    const bool wlive = q * 32 < RPq;                  // warp-uniform: rows the MN-major (query-row K dimension) reads touch
    const float cs = p.scale * kLog2e;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    int b_prev = -1;
    for (int i = 0; i < n; ++i) {
        const int tile = first + i, b = tile / p.H, h = tile - b * p.H, pr = dbl ? (i & 1) : 0;
        const uint32_t ph = (uint32_t)(i & 1), phqk = dbl ? (uint32_t)((i >> 1) & 1) : ph;
        const uint32_t uq = (uint32_t)(L.q0 + pr * L.ps), uk = (uint32_t)(L.k0 + pr * L.ps);
        trace_mark(trace, 0, tile);
        const bool newmask = p.mask != nullptr && b != b_prev;      // CTA-uniform
        b_prev = b;
        if (newmask) build_mask_bits(p.mask + (size_t)b * p.mask_bstride, dense ? Lq : 1, Lk, dense ? p.mask_rstride : 0, bits, warp, lane, 8);
        const float lse2 = lse_next * kLog2e;
        if (qok && i + 1 < n) lse_next = p.lse[(size_t)(tile + 1) * Lq + row];       // consumed one tile later
        if (t == 0) {
            mbar_wait(bar_qk0 + 8 * pr, phqk);
            mbar_wait(bar_in, ph);
            tcgen05_fence_after();
            trace_mark(trace, 2, tile);
            const uint32_t idesc = make_idesc(128, NS, false, false);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, make_smem_desc(base + uq + k * 32, 16, 1024), make_smem_desc(base + uk + k * 32, 16, 1024), idesc,
                          k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem + 128, make_smem_desc(base + L.dO + k * 32, 16, 1024), make_smem_desc(base + L.ds + U + k * 32, 16, 1024),
                          idesc, k > 0 ? 1u : 0u);
            umma_commit(bar_m1);
            // tile i-1's result boxes have left their units (Pd block 0 and the other pair's Q / K units): the softmax pass may
            // write Pd again, and that pair is free for tile i+1's Q / K
            if (i > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (dbl && i > 0 && i + 1 < n) {
                const int b1 = (tile + 1) / p.H, h1 = (tile + 1) % p.H;
                const uint32_t bq = bar_qk0 + 8 * (pr ^ 1);
                mbar_expect_tx(bq, qk_bytes);
                tma_load_2d(base + (L.q0 + (pr ^ 1) * L.ps), &tmQ, h1 * 64, b1 * Lq, bq);
                tma_load_2d(base + (L.k0 + (pr ^ 1) * L.ps), &tmK, h1 * 64, b1 * Lk, bq);
            }
        }
        mbar_wait(bar_in, ph);       // dO and the saved O are in shared memory
        float D = 0.f;
        if (qok) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 dv = *reinterpret_cast<const uint4*>(sm + L.dO + swz16(row, half * 4 + c));
                const uint4 ov = *reinterpret_cast<const uint4*>(sm + L.ds + swz16(row, half * 4 + c));
                const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&dv);
                const __nv_bfloat162* o = reinterpret_cast<const __nv_bfloat162*>(&ov);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 af = __bfloat1622float2(a[e]), of = __bfloat1622float2(o[e]);
                    D = fmaf(af.x, of.x, D);
                    D = fmaf(af.y, of.y, D);
                }
            }
        }
        red[half * 128 + row] = D;
        __syncthreads();             // also: mask bits visible, thread 0 is past its wait_group.read, every reader of O is done
        D = red[row] + red[128 + row];
        uint4 mb = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (p.mask) mb = *reinterpret_cast<const uint4*>(bits + (dense ? min(row, Lq - 1) : 0) * 4);
        const uint32_t drow32 = (uint32_t)((((size_t)b * p.H + h) * Lq + row) * ((Lk + 1) & ~1));    // dropout index base
        mbar_wait(bar_m1, ph);
        __syncwarp();
        tcgen05_fence_after();
        trace_mark(trace, 3, tile);
        float2 rs_p = make_float2(0.f, 0.f), rs_s = make_float2(0.f, 0.f);     // this thread's part of rowsum(Pd), rowsum(dS) (even | odd columns)
        if (wlive) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                const int c16 = half + 2 * ci;
                if (c16 < nchk) {
                    uint32_t pk[8], dk[8];
                    const uint32_t pair0 = (drow32 + (uint32_t)(c16 * 16)) >> 1;
                    bwd_chunk(trow + c16 * 16, trow + 128 + c16 * 16, mask16(mb, c16), Lk - c16 * 16, cs, lse2, D, p.scale, p.drop, pair0, qok, log2f((float)Lk),
                              pk, dk, rs_p, rs_s);
                    const uint32_t blk = (uint32_t)(c16 >> 2) * U;
                    const int ch = (c16 & 3) * 2;
                    if (row < RPq) {      // a unit holds ceil16(L) rows
                        *reinterpret_cast<uint4*>(sm + L.pd + blk + swz16(row, ch)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(sm + L.pd + blk + swz16(row, ch + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        *reinterpret_cast<uint4*>(sm + L.ds + blk + swz16(row, ch)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
                        *reinterpret_cast<uint4*>(sm + L.ds + blk + swz16(row, ch + 1)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
                    }
                }
            }
            // Bias gradients of the v / k projections = column sums of dV / dK over the keys.  colsum(dV) = sum_q rowsum(Pd)[q] dO[q, :]
            // and colsum(dK) = sum_q rowsum(dS)[q] Q[q, :], so the row sums, written as two extra key columns (126, 127: one per
            // column half; L <= 96 leaves them free), make the dV / dK MMAs deliver the column sums in their rows 126 / 127.
            if (want_bsum && row < RPq) {
                const uint32_t off = U + swz16(row, 7) + (uint32_t)(6 + half) * 2u;
                *reinterpret_cast<__nv_bfloat16*>(sm + L.pd + off) = __float2bfloat16_rn(rs_p.x + rs_p.y);      // 0 for rows past the sequence
                *reinterpret_cast<__nv_bfloat16*>(sm + L.ds + off) = __float2bfloat16_rn(rs_s.x + rs_s.y);
            }
        }
        fence_async_smem();
        tcgen05_fence_before();
        __syncthreads();
        trace_mark(trace, 4, tile);
        if (t == 0) {
            tcgen05_fence_after();
            const int nq = RPq >> 4, nk = NS >> 4;
            // dV[key, dk] = Pd^T dO : M = keys (MN-major A, two 64-key blocks U apart), K = query rows, N = dk
            const uint32_t idT = make_idesc(128, 64, true, true);
            for (int k = 0; k < nq; ++k)          // 16 query rows per MMA
                umma_bf16(tmem, make_smem_desc(base + L.pd + k * 2048, U, 1024), make_smem_desc(base + L.dO + k * 2048, 8192, 1024), idT,
                          k > 0 ? 1u : 0u);
            // dK[key, dk] = dS^T Q
            for (int k = 0; k < nq; ++k)
                umma_bf16(tmem + 64, make_smem_desc(base + L.ds + k * 2048, U, 1024), make_smem_desc(base + uq + k * 2048, 8192, 1024), idT,
                          k > 0 ? 1u : 0u);
            // dQ[q, dk] = dS K : A = dS K-major over keys, B = K MN-major
            const uint32_t idQ = make_idesc(128, 64, false, true);
            for (int k = 0; k < nk; ++k)
                umma_bf16(tmem + 128, make_smem_desc(base + L.ds + (k >> 2) * U + (k & 3) * 32, 16, 1024),
                          make_smem_desc(base + uk + k * 2048, 8192, 1024), idQ, k > 0 ? 1u : 0u);
            umma_commit(bar_m2);
        }
        mbar_wait(bar_m2, ph);
        __syncwarp();
        tcgen05_fence_after();
        trace_mark(trace, 5, tile);
        if (t == 0 && i + 1 < n) {         // dS and dO are dead: request the next tile's V, dO and saved O
            const int b1 = (tile + 1) / p.H, h1 = (tile + 1) % p.H;
            mbar_expect_tx(bar_in, in_bytes);
            tma_load_2d(base + L.ds + U, &tmV, h1 * 64, b1 * Lk, bar_in);
            tma_load_2d(base + L.dO, &tmDO, h1 * 64, b1 * Lq, bar_in);
            tma_load_2d(base + L.ds, &tmO, h1 * 64, b1 * Lq, bar_in);
        }
        // each thread: 32 of the 64 columns of its dV / dK row (row = key) and of its dQ row (row = query), parked in a dead unit
        auto stage_chunk = [&](uint32_t col, uint32_t unit, int Lr, int which) {
            float v[32];
            tmem_ld32(trow + col + half * 32, v);
            const int RPr = (Lr + 15) & ~15;
            if (row < RPr) {
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(v[g4 * 8 + 2 * e], v[g4 * 8 + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(sm + unit + swz16(row, half * 4 + g4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            if (want_bsum && which == 2) {
                if (row >= Lr) {        // rows past the sequence: not part of the tensor
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0.f;
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int e = 0; e < off; ++e) {
                        const float send = up ? v[e] : v[e + off];
                        const float keep = up ? v[e + off] : v[e];
                        v[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                atomicAdd(&csum[which * 64 + half * 32 + lane], v[0]);
            }
        };
        if (q * 32 < Lk) {
            stage_chunk(0, uk, Lk, 0);
            stage_chunk(64, (uint32_t)L.pd, Lk, 1);
        }
        if (q * 32 < Lq) stage_chunk(128, uq, Lq, 2);
        if (want_bsum && q == 3) {         // rows 126 / 127 of dV and dK: the column sums (two partial sums each)
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                float v[32];
                tmem_ld32(trow + which * 64 + half * 32, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float x = v[e] + __shfl_down_sync(0xffffffffu, v[e], 1);
                    if (lane == 30) csum[which * 64 + half * 32 + e] = x;       // this warp owns these slots
                }
            }
        }
        fence_async_smem();
        tcgen05_fence_before();
        __syncthreads();
        tcgen05_fence_after();
        if (t == 0) {
            tma_store_3d(&tmDV, base + uk, h * 64, 0, b);
            tma_store_3d(&tmDK, base + L.pd, h * 64, 0, b);
            tma_store_3d(&tmDQ, base + uq, h * 64, 0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (!dbl && i + 1 < n) {           // one (Q, K) pair: reload it as soon as the dQ / dV boxes have been read out of it
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                const int b1 = (tile + 1) / p.H, h1 = (tile + 1) % p.H;
                mbar_expect_tx(bar_qk0, qk_bytes);
                tma_load_2d(base + L.q0, &tmQ, h1 * 64, b1 * Lq, bar_qk0);
                tma_load_2d(base + L.k0, &tmK, h1 * 64, b1 * Lk, bar_qk0);
            }
        }
        if (want_bsum && t < 192) {
            float* dst = (t < 64) ? bp.bsum_v : (t < 128) ? bp.bsum_k : bp.bsum_q;
            if (dst) atomicAdd(dst + h * 64 + (t & 63), csum[t]);
            csum[t] = 0.f;
        }
        trace_mark(trace, 6, tile);
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static int operand_map(const void* ptr, int ld, int rows, int H, int box_rows, CUtensorMap* out) {
    return get_tensor_map(ptr, (uint64_t)H * 64, (uint64_t)rows, (uint64_t)ld * 2, 64, (uint32_t)box_rows, out);
}

static bool supported(const AttnParams& p) {
    return p.Lq <= 128 && p.Lk <= 128 && p.Lk >= 1 && p.Lq >= 1 && (p.ldq % 8) == 0 && (p.ldk % 8) == 0 && (p.ldv % 8) == 0 &&
           (p.ldo % 8) == 0;
}

static int launch_fwd(const AttnParams& p, cudaStream_t st) {
    const int RPq = (p.Lq + 15) & ~15, RPk = (p.Lk + 15) & ~15;
    CUtensorMap tq, tk, tv;
    GCT_TRY(operand_map(p.Q, p.ldq, p.B * p.Lq, p.H, RPq, &tq));
    GCT_TRY(operand_map(p.K, p.ldk, p.B * p.Lk, p.H, RPk, &tk));
    GCT_TRY(operand_map(p.V, p.ldv, p.B * p.Lk, p.H, RPk, &tv));
    {
        static PerDeviceSize done_;
        if (!done_.cur()) {
            GCT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdLayout(128, 128).total));
            GCT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            done_.cur() = 1;
        }
    }
    CUtensorMap to = tq;
    int box_store = 0;
    if (g_gct_attn_box && (reinterpret_cast<uintptr_t>(p.O) & 15) == 0 && (p.ldo % 8) == 0) {
        GCT_TRY(get_tensor_map3(p.O, (uint64_t)p.H * 64, (uint64_t)p.Lq, (uint64_t)p.B, (uint64_t)p.ldo * 2, 64, (uint32_t)RPq, &to));
        box_store = 1;
    }
    if (box_store && (g_gct_attn_persist & 1) && RPq <= 96 && RPk <= 96) {
        static PerDeviceSize pdone_;
        if (!pdone_.cur()) {
            GCT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdPersistLayout(96).total));
            GCT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_persist_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            pdone_.cur() = 1;
        }
        const int ntiles = p.B * p.H, slots = 3 * sm_count();
        GCT_CUDA(launch_k(attn_fwd_tc_persist_kernel, dim3(ntiles < slots ? ntiles : slots), dim3(256), (size_t)(FwdPersistLayout(RPq > RPk ? RPq : RPk).total), st, true, 
            tq, tk, tv, to, p, ntiles, g_gct_attn_trace));
        GCT_LAUNCH_CHECK();
        return GCT_OK;
    }
    GCT_CUDA(launch_k(attn_fwd_tc_kernel, dim3(p.B * p.H), dim3(256), (size_t)(FwdLayout(RPq, RPk).total), st, true, tq, tk, tv, to, box_store, p, g_gct_attn_trace));
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

static int launch_bwd(const AttnBwdParams& bp, cudaStream_t st) {
    const AttnParams& p = bp.f;
    GCT_REQUIRE(p.O, "attention bwd (tcgen05): the forward output is required (D = rowsum(dO * O))");
    const int RPq = (p.Lq + 15) & ~15, RPk = (p.Lk + 15) & ~15;
    CUtensorMap tq, tk, tv, tdo;
    GCT_TRY(operand_map(p.Q, p.ldq, p.B * p.Lq, p.H, RPq, &tq));
    GCT_TRY(operand_map(p.K, p.ldk, p.B * p.Lk, p.H, RPk, &tk));
    GCT_TRY(operand_map(p.V, p.ldv, p.B * p.Lk, p.H, RPk, &tv));
    GCT_TRY(operand_map(bp.dO, bp.lddo, p.B * p.Lq, p.H, RPq, &tdo));
    {
        static PerDeviceSize done_;
        if (!done_.cur()) {
            GCT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdLayout(128, 128).total));
            GCT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            done_.cur() = 1;
        }
    }
    CUtensorMap to = tq, tdq = tq, tdk = tq, tdv = tq;
    int box_io = 0;
    auto ok16 = [](const void* ptr, int ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 8) == 0; };
    if (g_gct_attn_box && ok16(p.O, p.ldo) && ok16(bp.dQ, bp.lddq) && ok16(bp.dK, bp.lddk) && ok16(bp.dV, bp.lddv)) {
        GCT_TRY(operand_map(p.O, p.ldo, p.B * p.Lq, p.H, RPq, &to));
        GCT_TRY(get_tensor_map3(bp.dQ, (uint64_t)p.H * 64, (uint64_t)p.Lq, (uint64_t)p.B, (uint64_t)bp.lddq * 2, 64, (uint32_t)RPq, &tdq));
        GCT_TRY(get_tensor_map3(bp.dK, (uint64_t)p.H * 64, (uint64_t)p.Lk, (uint64_t)p.B, (uint64_t)bp.lddk * 2, 64, (uint32_t)RPk, &tdk));
        GCT_TRY(get_tensor_map3(bp.dV, (uint64_t)p.H * 64, (uint64_t)p.Lk, (uint64_t)p.B, (uint64_t)bp.lddv * 2, 64, (uint32_t)RPk, &tdv));
        box_io = 1;
    }
    if (box_io && (g_gct_attn_persist & 2) && RPq <= 112 && RPk <= 112) {
        static PerDeviceSize pdone_;
        if (!pdone_.cur()) {
            GCT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdPersistLayout(96).total));
            GCT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_persist_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            pdone_.cur() = 1;
        }
        const int ntiles = p.B * p.H, slots = 2 * sm_count();
        GCT_CUDA(launch_k(attn_bwd_tc_persist_kernel, dim3(ntiles < slots ? ntiles : slots), dim3(256), (size_t)(BwdPersistLayout(RPq > RPk ? RPq : RPk).total), st, true, 
            tq, tk, tv, tdo, to, tdq, tdk, tdv, bp, ntiles, g_gct_attn_trace));
        GCT_LAUNCH_CHECK();
        return GCT_OK;
    }
    GCT_CUDA(launch_k(attn_bwd_tc_kernel, dim3(p.B * p.H), dim3(256), (size_t)(BwdLayout(RPq, RPk).total), st, true, tq, tk, tv, tdo, to, tdq, tdk, tdv, box_io, bp, g_gct_attn_trace));
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

}  // namespace atc
