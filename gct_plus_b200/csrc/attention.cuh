// Scaled-dot-product attention (Model/sublayers.py:29-41) for L <= 256, d_k = 64.
//   scores = QK^T/sqrt(dk); masked_fill(mask==0, -1e9); softmax; dropout on probs; PV.
// One CTA per (batch, head): K and V are staged once in shared memory, each warp owns query rows.
// The mask is a byte array [B, (1|Lq), Lk] (row stride 0 broadcasts one key-padding row), so any
// mask the caller builds with get_src_mask / get_trg_mask is honoured exactly -- including the
// "all keys masked -> uniform softmax" behaviour of the -1e9 fill.
// fp32 math throughout; T is the storage type (float for the parity tier, bf16 otherwise).
#pragma once
#include "common.cuh"

constexpr int ATT_DK = 64;
constexpr int ATT_KPAD = ATT_DK + 1;     // conflict-free column reads of K in shared memory
constexpr int ATT_MAXJ = 8;              // keys per lane -> Lk <= 256

struct AttnParams {
    const void* Q; const void* K; const void* V;     // [B, L, ld] with head h at column h*64
    int ldq, ldk, ldv;
    const uint8_t* mask; long long mask_bstride; int mask_rstride;
    void* O; int ldo;                                 // [B, Lq, ldo]
    float* lse;                                       // [B, H, Lq]   (log-sum-exp of the masked scores)
    float* probs;                                     // optional [B, H, Lq, Lk] (pre-dropout), get_attn
    int B, H, Lq, Lk;
    float scale;
    DropCtx drop;
};

template <typename T>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(AttnParams p) {
    extern __shared__ float smf[];
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int Lk = p.Lk, Lq = p.Lq;
    float* Ks = smf;                          // [Lk][65]
    float* Vs = Ks + (size_t)Lk * ATT_KPAD;   // [Lk][64]
    float* Qs = Vs + (size_t)Lk * ATT_DK;     // [nwarp][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const T* Kg = reinterpret_cast<const T*>(p.K) + (size_t)b * Lk * p.ldk + h * ATT_DK;
    const T* Vg = reinterpret_cast<const T*>(p.V) + (size_t)b * Lk * p.ldv + h * ATT_DK;
    for (int i = threadIdx.x; i < Lk * (ATT_DK / 8); i += blockDim.x) {
        const int j = i / (ATT_DK / 8), c = (i % (ATT_DK / 8)) * 8;
        f8 kv = ld8(Kg + (size_t)j * p.ldk + c);
        f8 vv = ld8(Vg + (size_t)j * p.ldv + c);
#pragma unroll
        for (int e = 0; e < 8; ++e) { Ks[j * ATT_KPAD + c + e] = kv.v[e]; Vs[j * ATT_DK + c + e] = vv.v[e]; }
    }
    __syncthreads();
    const T* Qg = reinterpret_cast<const T*>(p.Q) + (size_t)b * Lq * p.ldq + h * ATT_DK;
    T* Og = reinterpret_cast<T*>(p.O) + (size_t)b * Lq * p.ldo + h * ATT_DK;
    float* qs = Qs + warp * ATT_DK;
    const int nj = (Lk + 31) >> 5;
    for (int i = warp; i < Lq; i += nwarp) {
        qs[lane] = to_f(Qg[(size_t)i * p.ldq + lane]);
        qs[lane + 32] = to_f(Qg[(size_t)i * p.ldq + lane + 32]);
        __syncwarp();
        float s[ATT_MAXJ];
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj) s[jj] = 0.f;
#pragma unroll 8
        for (int d = 0; d < ATT_DK; ++d) {
            const float qd = qs[d];
#pragma unroll
            for (int jj = 0; jj < ATT_MAXJ; ++jj)
                if (jj < nj) {
                    const int j = min(jj * 32 + lane, Lk - 1);
                    s[jj] = fmaf(qd, Ks[j * ATT_KPAD + d], s[jj]);
                }
        }
        const uint8_t* mrow = p.mask ? p.mask + (size_t)b * p.mask_bstride + (size_t)i * p.mask_rstride : nullptr;
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj)
            if (jj < nj) {
                const int j = jj * 32 + lane;
                if (j < Lk) {
                    s[jj] *= p.scale;
                    if (mrow && mrow[j] == 0) s[jj] = -1e9f;
                } else {
                    s[jj] = -INFINITY;
                }
                mx = fmaxf(mx, s[jj]);
            }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj)
            if (jj < nj) { s[jj] = expf(s[jj] - mx); sum += s[jj]; }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        if (p.lse && lane == 0) p.lse[((size_t)b * p.H + h) * Lq + i] = mx + logf(sum);
        const size_t prow = (((size_t)b * p.H + h) * Lq + i) * Lk;
        const size_t drow = (((size_t)b * p.H + h) * Lq + i) * ((Lk + 1) & ~1);   // dropout index base (even stride)
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj)
            if (jj < nj) {
                const int j = jj * 32 + lane;
                s[jj] *= inv;
                if (j < Lk) {
                    if (p.probs) p.probs[prow + j] = s[jj];
                    s[jj] = drop_apply(p.drop, drow + j, s[jj]);
                }
            }
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj)
            if (jj < nj) {
                const int lim = min(32, Lk - jj * 32);
                for (int l = 0; l < lim; ++l) {
                    const float pj = __shfl_sync(0xffffffffu, s[jj], l);
                    const int j = jj * 32 + l;
                    o0 = fmaf(pj, Vs[j * ATT_DK + lane], o0);
                    o1 = fmaf(pj, Vs[j * ATT_DK + lane + 32], o1);
                }
            }
        Og[(size_t)i * p.ldo + lane] = from_f<T>(o0);
        Og[(size_t)i * p.ldo + lane + 32] = from_f<T>(o1);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// backward: recompute P from the saved LSE.  Phase A (warp per query row): dP = dO V^T,
// D = sum(P_drop * dP), dS = P * (drop'(dP) - D), dQ = scale * dS K;  dS and P_drop are parked in
// shared memory.  Phase B (warp per key): dK = scale * dS^T Q, dV = P_drop^T dO -- no atomics.
// ------------------------------------------------------------------------------------------
struct AttnBwdParams {
    AttnParams f;                 // forward operands (O unused)
    const void* dO; int lddo;     // [B, Lq, lddo]
    void* dQ; void* dK; void* dV; // same layouts / pitches as Q, K, V
    int lddq, lddk, lddv;
    // optional: column sums of dQ / dK / dV over all rows = the bias gradients of the q / k / v projections ([H*64] each,
    // accumulated with atomics).  Honoured by the tcgen05 kernel only (atc::launch_bwd); other paths ignore them.
    float* bsum_q = nullptr; float* bsum_k = nullptr; float* bsum_v = nullptr;
};

template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(AttnBwdParams bp) {
    const AttnParams& p = bp.f;
    extern __shared__ float smf[];
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int Lk = p.Lk, Lq = p.Lq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    // shared: dS[Lq][Lk], Pd[Lq][Lk] (fp32), then T tiles K[Lk][64+2], V[Lk][64+2], Q[Lq][64], dO[Lq][64]
    float* dSs = smf;
    float* Pds = dSs + (size_t)Lq * Lk;
    T* Ks = reinterpret_cast<T*>(Pds + (size_t)Lq * Lk);
    constexpr int KP = ATT_DK + 2;                    // pad (keeps 4-byte alignment for bf16 pairs)
    T* Vs = Ks + (size_t)Lk * KP;
    T* Qs = Vs + (size_t)Lk * KP;
    T* dOs = Qs + (size_t)Lq * ATT_DK;
    const T* Qg = reinterpret_cast<const T*>(p.Q) + (size_t)b * Lq * p.ldq + h * ATT_DK;
    const T* Kg = reinterpret_cast<const T*>(p.K) + (size_t)b * Lk * p.ldk + h * ATT_DK;
    const T* Vg = reinterpret_cast<const T*>(p.V) + (size_t)b * Lk * p.ldv + h * ATT_DK;
    const T* dOg = reinterpret_cast<const T*>(bp.dO) + (size_t)b * Lq * bp.lddo + h * ATT_DK;
    for (int i = threadIdx.x; i < Lk * ATT_DK; i += blockDim.x) {
        const int j = i / ATT_DK, c = i % ATT_DK;
        Ks[j * KP + c] = Kg[(size_t)j * p.ldk + c];
        Vs[j * KP + c] = Vg[(size_t)j * p.ldv + c];
    }
    for (int i = threadIdx.x; i < Lq * ATT_DK; i += blockDim.x) {
        const int r = i / ATT_DK, c = i % ATT_DK;
        Qs[i] = Qg[(size_t)r * p.ldq + c];
        dOs[i] = dOg[(size_t)r * bp.lddo + c];
    }
    __syncthreads();
    const int nj = (Lk + 31) >> 5;
    T* dQg = reinterpret_cast<T*>(bp.dQ) + (size_t)b * Lq * bp.lddq + h * ATT_DK;
    for (int i = warp; i < Lq; i += nwarp) {
        float s[ATT_MAXJ], dp[ATT_MAXJ];
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj) { s[jj] = 0.f; dp[jj] = 0.f; }
#pragma unroll 4
        for (int d = 0; d < ATT_DK; ++d) {
            const float qd = to_f(Qs[i * ATT_DK + d]), gd = to_f(dOs[i * ATT_DK + d]);
#pragma unroll
            for (int jj = 0; jj < ATT_MAXJ; ++jj)
                if (jj < nj) {
                    const int j = min(jj * 32 + lane, Lk - 1);
                    s[jj] = fmaf(qd, to_f(Ks[j * KP + d]), s[jj]);
                    dp[jj] = fmaf(gd, to_f(Vs[j * KP + d]), dp[jj]);
                }
        }
        const uint8_t* mrow = p.mask ? p.mask + (size_t)b * p.mask_bstride + (size_t)i * p.mask_rstride : nullptr;
        const float lse = p.lse[((size_t)b * p.H + h) * Lq + i];
        const size_t prow = (((size_t)b * p.H + h) * Lq + i) * Lk;
        const size_t drow = (((size_t)b * p.H + h) * Lq + i) * ((Lk + 1) & ~1);   // dropout index base (even stride)
        float Dsum = 0.f;
        float pd[ATT_MAXJ];
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj)
            if (jj < nj) {
                const int j = jj * 32 + lane;
                float pr = 0.f;
                pd[jj] = 0.f;
                if (j < Lk) {
                    float sc = s[jj] * p.scale;
                    if (mrow && mrow[j] == 0) sc = -1e9f;
                    // a row without one visible key: every score is the fill value, softmax is uniform -- and lse = -1e9 + log(Lk)
                    // has lost the log(Lk) to fp32 rounding, so exp(sc - lse) would be 1
                    pr = (lse <= -5e8f) ? 1.f / (float)Lk : expf(sc - lse);
                    pd[jj] = drop_apply(p.drop, drow + j, pr);          // dropped probability used in PV
                    dp[jj] = drop_apply(p.drop, drow + j, dp[jj]);      // gradient through the same mask
                    Dsum += pr * dp[jj];
                }
                s[jj] = pr;
            }
        Dsum = warp_sum(Dsum);
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj)
            if (jj < nj) {
                const int j = jj * 32 + lane;
                if (j < Lk) {
                    float ds = s[jj] * (dp[jj] - Dsum);
                    if (lse <= -5e8f && mrow && mrow[j] == 0) ds = 0.f;     // masked_fill passes no gradient to a masked score
                    dSs[(size_t)i * Lk + j] = ds;
                    Pds[(size_t)i * Lk + j] = pd[jj];
                    s[jj] = ds;
                } else {
                    s[jj] = 0.f;
                }
            }
        // dQ_i = scale * sum_j dS_ij K_j
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < ATT_MAXJ; ++jj)
            if (jj < nj) {
                const int lim = min(32, Lk - jj * 32);
                for (int l = 0; l < lim; ++l) {
                    const float dsj = __shfl_sync(0xffffffffu, s[jj], l);
                    const int j = jj * 32 + l;
                    q0 = fmaf(dsj, to_f(Ks[j * KP + lane]), q0);
                    q1 = fmaf(dsj, to_f(Ks[j * KP + lane + 32]), q1);
                }
            }
        dQg[(size_t)i * bp.lddq + lane] = from_f<T>(q0 * p.scale);
        dQg[(size_t)i * bp.lddq + lane + 32] = from_f<T>(q1 * p.scale);
    }
    __syncthreads();
    T* dKg = reinterpret_cast<T*>(bp.dK) + (size_t)b * Lk * bp.lddk + h * ATT_DK;
    T* dVg = reinterpret_cast<T*>(bp.dV) + (size_t)b * Lk * bp.lddv + h * ATT_DK;
    for (int j = warp; j < Lk; j += nwarp) {
        float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int i = 0; i < Lq; ++i) {
            const float ds = dSs[(size_t)i * Lk + j], pdv = Pds[(size_t)i * Lk + j];
            k0 = fmaf(ds, to_f(Qs[i * ATT_DK + lane]), k0);
            k1 = fmaf(ds, to_f(Qs[i * ATT_DK + lane + 32]), k1);
            v0 = fmaf(pdv, to_f(dOs[i * ATT_DK + lane]), v0);
            v1 = fmaf(pdv, to_f(dOs[i * ATT_DK + lane + 32]), v1);
        }
        dKg[(size_t)j * bp.lddk + lane] = from_f<T>(k0 * p.scale);
        dKg[(size_t)j * bp.lddk + lane + 32] = from_f<T>(k1 * p.scale);
        dVg[(size_t)j * bp.lddv + lane] = from_f<T>(v0);
        dVg[(size_t)j * bp.lddv + lane + 32] = from_f<T>(v1);
    }
}

template <typename T>
static size_t attn_bwd_smem(int Lq, int Lk) {
    return (size_t)2 * Lq * Lk * 4 + ((size_t)2 * Lk * (ATT_DK + 2) + (size_t)2 * Lq * ATT_DK) * sizeof(T);
}
static size_t attn_fwd_smem(int Lk, int nwarp) {
    return ((size_t)Lk * ATT_KPAD + (size_t)Lk * ATT_DK + (size_t)nwarp * ATT_DK) * 4;
}

template <typename T>
static int launch_attn_fwd(const AttnParams& p, cudaStream_t st) {
    GCT_REQUIRE(p.Lk >= 1 && p.Lk <= 32 * ATT_MAXJ, "attention: Lk=%d outside [1,%d]", p.Lk, 32 * ATT_MAXJ);
    GCT_REQUIRE((p.ldk % 8) == 0 && (p.ldv % 8) == 0, "attention: K/V pitch must be a multiple of 8 elements");
    const size_t sm = attn_fwd_smem(p.Lk, 8);
    GCT_SMEM_LIMIT(attn_fwd_kernel<T>, sm);
    attn_fwd_kernel<T><<<p.B * p.H, 256, sm, st>>>(p);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

template <typename T>
static int launch_attn_bwd(const AttnBwdParams& bp, cudaStream_t st) {
    const AttnParams& p = bp.f;
    GCT_REQUIRE(p.Lk >= 1 && p.Lk <= 32 * ATT_MAXJ, "attention bwd: Lk=%d outside [1,%d]", p.Lk, 32 * ATT_MAXJ);
    const size_t sm = attn_bwd_smem<T>(p.Lq, p.Lk);
    if (sm > 227 * 1024) GCT_FAIL(GCT_ERR_UNSUPPORTED, "attention bwd: Lq=%d Lk=%d needs %zu B of shared memory", p.Lq, p.Lk, sm);
    GCT_SMEM_LIMIT(attn_bwd_kernel<T>, sm);
    attn_bwd_kernel<T><<<p.B * p.H, 256, sm, st>>>(bp);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
