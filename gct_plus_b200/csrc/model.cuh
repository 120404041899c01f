// Whole-model runtime: Vaetf / Cvaetf forward (Model/vaetf.py:154-182, Model/cvaetf.py:177-193),
// hand-written backward, and the KV-cached decoder.  Pure orchestration of the kernels in
// elementwise.cuh / gemm_*.cuh / attention.cuh / decode.cuh over caller-owned buffers.
#pragma once
#include "../../include/gct_b200.h"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "decode.cuh"
#include "decode_zattn.cuh"
#include "elementwise.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_rownorm.cuh"

extern int g_gct_rownorm;
extern int g_gct_simt_only;
extern int g_gct_simt_attn;
extern int g_gct_zattn;
extern int g_gct_ffn_classic;
extern int g_gct_attn_bias_separate;

template <typename T>
static int attn_fwd_dispatch(const AttnParams& p, cudaStream_t st) {
    if constexpr (sizeof(T) == 2) {
        if (!g_gct_simt_attn && atc::supported(p)) return atc::launch_fwd(p, st);
    }
    return launch_attn_fwd<T>(p, st);
}
template <typename T>
static int attn_bwd_dispatch(const AttnBwdParams& bp, cudaStream_t st) {
    if constexpr (sizeof(T) == 2) {
        if (!g_gct_simt_attn && atc::supported(bp.f) && bp.f.O) return atc::launch_bwd(bp, st);
    }
    return launch_attn_bwd<T>(bp, st);
}

// ------------------------------------------------------------------------------------------
// bump allocator over the caller's workspace (dry run with base == nullptr measures the size)
// ------------------------------------------------------------------------------------------
struct Bump {
    uint8_t* base; size_t off;
    explicit Bump(void* b) : base(reinterpret_cast<uint8_t*>(b)), off(0) {}
    void* take(size_t bytes) {
        off = (off + 255) & ~(size_t)255;
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
    template <typename U> U* arr(size_t n) { return reinterpret_cast<U*>(take(n * sizeof(U))); }
};

enum EncSlot { E_N1A, E_N1B, E_QKV_W, E_QKV_B, E_O_W, E_O_B, E_N2A, E_N2B, E_F1_W, E_F1_B, E_F2_W, E_F2_B };
enum DecSlot { D_N1A, D_N1B, D_QKV_W, D_QKV_B, D_O1_W, D_O1_B, D_N2A, D_N2B, D_Q2_W, D_Q2_B, D_KV2_W, D_KV2_B,
               D_O2_W, D_O2_B, D_N3A, D_N3B, D_F1_W, D_F1_B, D_F2_W, D_F2_B };

// dropout sites
enum Site { S_ENC_PE = 1, S_DEC_PE = 2, S_ENC_BASE = 16, S_DEC_BASE = 4096 };
enum EncSite { ES_ATTN, ES_DROP1, ES_FF, ES_DROP2, ES_COUNT };
enum DecSite { DS_ATTN1, DS_DROP1, DS_ATTN2, DS_DROP2, DS_FF, DS_DROP3, DS_COUNT };

template <typename T>
struct Model {
    gct_config_t c;
    gct_weights_t w;
    cudaStream_t st;
    int d, dff, H, lat, N, nc;
    DropCtx drop0;           // base dropout ctx (thresh = 0 when not training)

    int enc_slot(int l, int s) const { return GCT_NUM_GLOBAL_SLOTS + l * GCT_ENC_LAYER_SLOTS + s; }
    int dec_slot(int l, int s) const { return GCT_NUM_GLOBAL_SLOTS + N * GCT_ENC_LAYER_SLOTS + l * GCT_DEC_LAYER_SLOTS + s; }
    bool has(int slot) const { return w.slot_offsets_host[slot] >= 0; }
    const float* P(int slot) const { return w.params_f32 + w.slot_offsets_host[slot]; }
    float* G(int slot) const { return w.grads_f32 + w.slot_offsets_host[slot]; }
    // GEMM-operand view of a weight: fp32 master in the parity tier, bf16 shadow otherwise
    const T* WT(int slot) const {
        if constexpr (sizeof(T) == 4) return reinterpret_cast<const T*>(w.params_f32 + w.slot_offsets_host[slot]);
        else return reinterpret_cast<const T*>(w.params_bf16) + w.slot_offsets_host[slot];
    }
    DropCtx site(uint32_t s) const { return drop_site(drop0, s); }

    int init(const gct_config_t* cfg, const gct_weights_t* wts, uint32_t seed, int train, void* stream) {
        c = *cfg; w = *wts; st = reinterpret_cast<cudaStream_t>(stream);
        d = c.d_model; dff = c.d_ff; H = c.heads; lat = c.latent_dim; N = c.n_layers; nc = c.nconds;
        GCT_REQUIRE(d % 128 == 0 && d <= 1024, "d_model=%d must be a multiple of 128 and <= 1024", d);
        GCT_REQUIRE(d == H * 64, "heads*64 must equal d_model (d_k is fixed at 64), got d=%d H=%d", d, H);
        GCT_REQUIRE(dff % 64 == 0 && lat % 32 == 0, "d_ff %% 64 and latent %% 32 must be 0");
        GCT_REQUIRE(nc >= 0 && nc <= 8, "nconds=%d outside [0,8]", nc);
        GCT_REQUIRE(N >= 1 && N <= 16, "n_layers=%d outside [1,16]", N);
        GCT_REQUIRE(c.trg_vocab <= 96 && c.src_vocab >= 1 && c.src_vocab <= 96, "vocabularies must be <= 96 ids (got %d / %d)", c.src_vocab, c.trg_vocab);
        GCT_REQUIRE(w.params_f32 && w.slot_offsets_host, "weights missing");
        if (sizeof(T) == 2) GCT_REQUIRE(w.params_bf16, "bf16 shadow parameters missing");
        drop0.seed = seed; drop0.scale = 1.f; drop0.thresh = 0;
        if (train && c.dropout > 0.f) {
            drop0.thresh = (uint32_t)fmin(4294967295.0, (double)c.dropout * 4294967296.0);
            drop0.scale = 1.f / (1.f - c.dropout);
        }
        return GCT_OK;
    }

    // ---------------- GEMM dispatch ----------------
    int gemm(const T* A, bool a_mn, long long lda, const T* B, bool b_mn, long long ldb, int M, int Nn, int K,
             Epilogue e, int split_k = 1, int bn_hint = 0) {
        if (e.alpha == 0.f) e.alpha = 1.f;
        if constexpr (sizeof(T) == 4) {
            return launch_gemm_simt<T, T, T>(A, a_mn ? 1 : lda, a_mn ? lda : 1, B, b_mn ? 1 : ldb, b_mn ? ldb : 1, M, Nn, K,
                                             split_k, e, st);
        } else {
            if (g_gct_simt_only)
                return launch_gemm_simt<T, T, T>(A, a_mn ? 1 : lda, a_mn ? lda : 1, B, b_mn ? 1 : ldb, b_mn ? ldb : 1, M, Nn,
                                                 K, split_k, e, st);
            return tc::launch_gemm_tc(A, a_mn, lda, B, b_mn, ldb, M, Nn, K, split_k, bn_hint, e, st);
        }
    }
    static Epilogue epi(const float* bias, int ldc) {
        Epilogue e; memset(&e, 0, sizeof(e));
        e.bias = bias; e.ldc = ldc; e.alpha = 1.f; e.drop.thresh = 0; e.drop.scale = 1.f;
        return e;
    }
    // y = x W^T + b  -> T
    int linear_T(const T* x, int M, int K, int wslot, int bslot, int Nn, T* y) {
        Epilogue e = epi(P(bslot), Nn); e.outT = y;
        return gemm(x, false, K, WT(wslot), false, K, M, Nn, K, e);
    }
    // split-K factor for weight gradients.  A work item of the CTA-pair kernel is a 256 x 256 tile of dW over R / s token rows, and it
    // ends in 256 KB of fp32 atomics: the items should fill the 74 SM pairs once (small dW: the atomics of a second wave cost more
    // than its shorter main loop saves) or twice (12+ tiles), never a little more than a whole number of waves.
    // [B200] profiles/r02_sweep_wgrad_split.txt, R = 41 472: 512 x 512 41.9 us at s = 37 (two waves) -> 29.9 us at 18; 1024 x 512
    // 61.4 (19) -> 51.1 (9); 1536 x 512 73.7 (13) -> 65.4 (9); 2048 x 512 86.6 (10) -> 77.1 (9).
    int wgrad_split(int Mout, int Nout, int R) const {
        const int tiles = cdiv(Mout, 256) * cdiv(Nout, 256);
        int s = (tiles <= 8 ? 72 : 144) / tiles;
        const int kb = cdiv(R, 64);
        if (s > kb / 2) s = kb / 2;
        return s < 1 ? 1 : s;
    }
    // dW[Nout,Kin] += dY^T[Nout,R] X[R,Kin] ; db += colsum(dY)
    int wgrad(const T* dY, int ldy, const T* X, int ldx, int R, int Nout, int Kin, int wslot, int bslot, bool do_bias) {
        Epilogue e = epi(nullptr, Kin); e.out32 = G(wslot); e.flags = EPI_ACCUM;
        GCT_TRY(gemm(dY, true, ldy, X, true, ldx, Nout, Kin, R, e, wgrad_split(Nout, Kin, R)));
        if (do_bias) {
            const int gy = cdiv(Nout, 256);
            dim3 grid(max(1, min(cdiv(R, 64), 1184 / gy)), gy);   // 8 CTAs of 256 threads per SM: full occupancy (ncu: 4 per SM left DRAM at 4.0 TB/s)
            GCT_CUDA(launch_k(colsum_kernel<T>, dim3(grid), dim3(256), (size_t)(0), st, true, dY, R, Nout, ldy, G(bslot)));
            GCT_LAUNCH_CHECK();
        }
        return GCT_OK;
    }
    int norm_fwd(const float* x, int aslot, int bslot, T* y, float* y32, int rows) {
        const int nv = d / 128;
        dim3 grid(cdiv(rows, 8));
#define GCT_NORM_CASE(NV) case NV: GCT_CUDA(launch_k(norm_fwd_kernel<T, NV>, grid, dim3(256), 0, st, true, x, P(aslot), P(bslot), y, y32, rows, 1e-6f)); break;
        switch (nv) { GCT_NORM_CASE(1) GCT_NORM_CASE(2) GCT_NORM_CASE(3) GCT_NORM_CASE(4) GCT_NORM_CASE(5) GCT_NORM_CASE(6)
                      GCT_NORM_CASE(7) GCT_NORM_CASE(8) default: GCT_FAIL(GCT_ERR_UNSUPPORTED, "norm width %d", d); }
#undef GCT_NORM_CASE
        GCT_LAUNCH_CHECK();
        return GCT_OK;
    }
    // x_out = res + dropout(A W^T + bias) followed by (normT [, norm32]) = Norm(x_out): one kernel (gemm_rownorm.cuh) when the
    // bf16 tier runs at d_model = 512 with enough row tiles to fill the machine, otherwise the residual GEMM + norm_fwd pair
    bool rownorm_ok(int M) const {
        if (sizeof(T) != 2 || d != tc::RN_N || g_gct_simt_only || g_gct_rownorm == 0) return false;
        return g_gct_rownorm == 2 || cdiv(M, 128) >= 96;
    }
    int linear_res_norm(const T* A, int M, int K, const T* Wp, const float* bias, const float* res32, float* out32, DropCtx drop,
                        int aslot, int bslot, T* normT, float* norm32) {
        if constexpr (sizeof(T) == 2) {
            if (rownorm_ok(M)) {
                tc::RowNormParams rp;
                rp.bias = bias; rp.res32 = res32; rp.out32 = out32; rp.alpha = P(aslot); rp.beta = P(bslot); rp.norm32 = norm32;
                rp.drop = drop; rp.eps = 1e-6f; rp.M = M; rp.K = K;
                return tc::launch_gemm_rownorm(A, K, Wp, K, normT, rp, st);
            }
        }
        Epilogue e = epi(bias, d);
        e.res32 = res32; e.out32 = out32; e.drop = drop;
        GCT_TRY(gemm(A, false, K, Wp, false, K, M, d, K, e));
        return norm_fwd(out32, aslot, bslot, normT, norm32, M);
    }
    // dropT / dc / dropsum: fused prologue of the consumer (see norm_bwd_kernel); null = plain Norm backward
    int norm_bwd(const float* x, int aslot, int bslot, const float* dy, const float* add, float* dx, int rows,
                 T* dropT = nullptr, DropCtx dc = DropCtx{0, 0, 1.f}, float* dropsum = nullptr) {
        const int nv = d / 128;
        dim3 grid(min(cdiv(rows, 8), 148 * 3));
        const size_t sm = (size_t)8 * 3 * d * sizeof(float);
#define GCT_NORMB_CASE(NV) case NV: if (sm > 48 * 1024) GCT_SMEM_LIMIT((norm_bwd_kernel<T, NV>), sm); GCT_CUDA(launch_k(norm_bwd_kernel<T, NV>, dim3(grid), dim3(256), (size_t)(sm), st, true, x, P(aslot), dy, add, dx, G(aslot), G(bslot), rows, 1e-6f, dropT, dc, dropsum)); break;
        switch (nv) { GCT_NORMB_CASE(1) GCT_NORMB_CASE(2) GCT_NORMB_CASE(3) GCT_NORMB_CASE(4) GCT_NORMB_CASE(5) GCT_NORMB_CASE(6)
                      GCT_NORMB_CASE(7) GCT_NORMB_CASE(8) default: GCT_FAIL(GCT_ERR_UNSUPPORTED, "norm width %d", d); }
#undef GCT_NORMB_CASE
        GCT_LAUNCH_CHECK();
        return GCT_OK;
    }
    // out(T)[r,c] = dropmask(site)*in ; optional bias grad
    int cast_drop(const float* in, T* out, int rows, int cols, DropCtx dc, float* colsum) {
        const int gy = cdiv(cols, 256);
        dim3 grid(max(1, min(cdiv(rows, 32), 592 / gy)), gy);
        GCT_CUDA(launch_k(cast_drop_colsum_kernel<T>, dim3(grid), dim3(256), (size_t)(0), st, true, in, out, rows, cols, dc, colsum));
        GCT_LAUNCH_CHECK();
        return GCT_OK;
    }
    int attention(const T* q, int ldq, const T* k, const T* v, int ldkv, const uint8_t* mask, long long mb, int mr, T* out,
                  float* lse, float* probs, int B, int Lq, int Lk, DropCtx dc) {
        AttnParams p;
        p.Q = q; p.K = k; p.V = v; p.ldq = ldq; p.ldk = ldkv; p.ldv = ldkv; p.mask = mask; p.mask_bstride = mb;
        p.mask_rstride = mr; p.O = out; p.ldo = d; p.lse = lse; p.probs = probs; p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
        p.scale = 0.125f; p.drop = dc;
        return attn_fwd_dispatch<T>(p, st);
    }
    // gq / gk / gv: bias gradients of the q / k / v projections (column sums of dq / dk / dv), produced inside the tcgen05
    // kernel's write-out, by separate column-sum launches on the other paths
    int attention_bwd(const T* q, int ldq, const T* k, const T* v, int ldkv, const uint8_t* mask, long long mb, int mr,
                      const float* lse, const T* out, const T* dO, T* dq, int lddq, T* dk, T* dv, int lddkv, int B, int Lq, int Lk,
                      DropCtx dc, float* gq = nullptr, float* gk = nullptr, float* gv = nullptr) {
        AttnBwdParams bp;
        AttnParams& p = bp.f;
        p.Q = q; p.K = k; p.V = v; p.ldq = ldq; p.ldk = ldkv; p.ldv = ldkv; p.mask = mask; p.mask_bstride = mb;
        p.mask_rstride = mr; p.O = const_cast<T*>(out); p.ldo = d; p.lse = const_cast<float*>(lse); p.probs = nullptr; p.B = B; p.H = H;
        p.Lq = Lq; p.Lk = Lk; p.scale = 0.125f; p.drop = dc;
        bp.dO = dO; bp.lddo = d; bp.dQ = dq; bp.dK = dk; bp.dV = dv; bp.lddq = lddq; bp.lddk = lddkv; bp.lddv = lddkv;
        bool fused = false;
        if constexpr (sizeof(T) == 2) fused = !g_gct_simt_attn && !g_gct_attn_bias_separate && atc::supported(bp.f) && bp.f.O != nullptr;
        if (fused) { bp.bsum_q = gq; bp.bsum_k = gk; bp.bsum_v = gv; }
        GCT_TRY(attn_bwd_dispatch<T>(bp, st));
        if (!fused) {
            auto cs = [&](const T* g, int rows, int ld, float* dst) -> int {
                if (!dst) return GCT_OK;
                const int gy = cdiv(d, 256);
                dim3 grid(max(1, min(cdiv(rows, 64), 1184 / gy)), gy);
                GCT_CUDA(launch_k(colsum_kernel<T>, dim3(grid), dim3(256), (size_t)(0), st, true, g, rows, d, ld, dst));
                GCT_LAUNCH_CHECK();
                return GCT_OK;
            };
            GCT_TRY(cs(dq, B * Lq, lddq, gq));
            GCT_TRY(cs(dk, B * Lk, lddkv, gk));
            GCT_TRY(cs(dv, B * Lk, lddkv, gv));
        }
        return GCT_OK;
    }
};

// ------------------------------------------------------------------------------------------
// saved activations
// ------------------------------------------------------------------------------------------
template <typename T>
struct EncLayerAct { T* a1; float* a1_32; T* qkv; T* att; float* lse; float* x1; T* a2; float* a2_32; T* hpre; T* g; float* xout; };
template <typename T>
struct DecLayerAct { T* a1; T* qkv; T* att1; float* lse1; float* y1; T* a2; T* q2; T* kv2; T* att2; float* lse2; float* y2;
                     T* a3; T* hpre; T* g; float* yout; };

template <typename T>
struct Acts {
    int B, S, Tt, Se, Sm, Ld, Me, Mm, Md, Vpad;
    float* x0; EncLayerAct<T> enc[16]; T* xe; float* mulv;
    T* zpad; T* mem; float* y0; DecLayerAct<T> dec[16]; T* yd;
    uint8_t* cross_mask;     // [B, Sm]
    size_t bytes;

    void carve(const gct_config_t& c, int B_, int S_, int T_, void* ws) {
        Bump bp(ws);
        B = B_; S = S_; Tt = T_;
        const int d = c.d_model, dff = c.d_ff, lat = c.latent_dim, H = c.heads, nc = c.nconds;
        Se = nc + S; Sm = Se + ((c.use_cond2lat && nc > 0 && !c.use_cond2dec) ? nc : 0);
        Ld = Tt + ((c.use_cond2dec && nc > 0) ? nc : 0);
        Me = B * Se; Mm = B * Sm; Md = B * Ld;
        Vpad = (c.trg_vocab + 31) / 32 * 32;
        x0 = bp.arr<float>((size_t)Me * d);
        for (int l = 0; l < c.n_layers; ++l) {
            EncLayerAct<T>& a = enc[l];
            a.a1 = bp.arr<T>((size_t)Me * d); a.a1_32 = bp.arr<float>((size_t)Me * d);
            a.qkv = bp.arr<T>((size_t)Me * 3 * d); a.att = bp.arr<T>((size_t)Me * d);
            a.lse = bp.arr<float>((size_t)B * H * Se); a.x1 = bp.arr<float>((size_t)Me * d);
            a.a2 = bp.arr<T>((size_t)Me * d); a.a2_32 = bp.arr<float>((size_t)Me * d);
            a.hpre = bp.arr<T>((size_t)Me * dff); a.g = bp.arr<T>((size_t)Me * dff);
            a.xout = bp.arr<float>((size_t)Me * d);
        }
        xe = bp.arr<T>((size_t)Me * d);
        mulv = bp.arr<float>((size_t)Me * 2 * lat);
        zpad = bp.arr<T>((size_t)Mm * lat);
        mem = bp.arr<T>((size_t)Mm * d);
        y0 = bp.arr<float>((size_t)Md * d);
        for (int l = 0; l < c.n_layers; ++l) {
            DecLayerAct<T>& a = dec[l];
            a.a1 = bp.arr<T>((size_t)Md * d); a.qkv = bp.arr<T>((size_t)Md * 3 * d); a.att1 = bp.arr<T>((size_t)Md * d);
            a.lse1 = bp.arr<float>((size_t)B * H * Ld); a.y1 = bp.arr<float>((size_t)Md * d);
            a.a2 = bp.arr<T>((size_t)Md * d); a.q2 = bp.arr<T>((size_t)Md * d); a.kv2 = bp.arr<T>((size_t)Mm * 2 * d);
            a.att2 = bp.arr<T>((size_t)Md * d); a.lse2 = bp.arr<float>((size_t)B * H * Ld); a.y2 = bp.arr<float>((size_t)Md * d);
            a.a3 = bp.arr<T>((size_t)Md * d); a.hpre = bp.arr<T>((size_t)Md * dff); a.g = bp.arr<T>((size_t)Md * dff);
            a.yout = bp.arr<float>((size_t)Md * d);
        }
        yd = bp.arr<T>((size_t)Md * d);
        cross_mask = bp.arr<uint8_t>((size_t)B * Sm);
        bytes = bp.off + 256;
    }
};

__global__ void cross_mask_kernel(const uint8_t* __restrict__ src_mask, int B, int Se, int Sm, uint8_t* __restrict__ out) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Sm) return;
    const int b = i / Sm, j = i % Sm, off = Sm - Se;
    out[i] = (j < off) ? 1 : src_mask[(size_t)b * Se + (j - off)];
}
// z (fp32, caller supplied) -> zpad (T) with `off` leading rows per batch left untouched
template <typename T>
__global__ void zpad_kernel(const float* __restrict__ z, int B, int Se, int Sm, int lat, T* __restrict__ zpad) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * Sm * lat) return;
    const int c = (int)(i % lat);
    const size_t row = i / lat;
    const int b = (int)(row / Sm), j = (int)(row % Sm), off = Sm - Se;
    zpad[i] = (j < off) ? from_f<T>(0.f) : from_f<T>(z[((size_t)b * Se + (j - off)) * lat + c]);
}
template <typename T>
__global__ void zero_rows_kernel(T* __restrict__ x, int B, int L, int nrows, int cols) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * nrows * cols) return;
    const int c = (int)(i % cols);
    const size_t r = i / cols;
    const int b = (int)(r / nrows), j = (int)(r % nrows);
    x[((size_t)b * L + j) * cols + c] = from_f<T>(0.f);
}
// dz[b,s,:] = dzpad[b, off+s, :] (+ dz_ext)
__global__ void dz_gather_kernel(const float* __restrict__ dzpad, const float* __restrict__ dz_ext, int B, int Se, int Sm,
                                 int lat, float* __restrict__ dz) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * Se * lat) return;
    const int c = (int)(i % lat);
    const size_t row = i / lat;
    const int b = (int)(row / Se), s = (int)(row % Se), off = Sm - Se;
    float v = dzpad[((size_t)b * Sm + off + s) * lat + c];
    if (dz_ext) v += dz_ext[i];
    dz[i] = v;
}
// fp32 [rows, V] -> T [rows, Vpad] zero padded
template <typename T>
__global__ void pad_cast_kernel(const float* __restrict__ in, int rows, int V, int Vpad, T* __restrict__ out) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * Vpad) return;
    const int c = (int)(i % Vpad);
    const size_t r = i / Vpad;
    out[i] = from_f<T>(c < V ? in[r * V + c] : 0.f);
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <typename T>
static int model_forward(Model<T>& m, const gct_io_t& io, Acts<T>& A) {
    const int d = m.d, dff = m.dff, lat = m.lat, nc = m.nc, N = m.N;
    const int B = A.B, Se = A.Se, Sm = A.Sm, Ld = A.Ld, Me = A.Me, Mm = A.Mm, Md = A.Md;
    cudaStream_t st = m.st;
    const float sqd = sqrtf((float)d);

    if (io.run_encoder) {
        GCT_REQUIRE(io.src && io.src_mask, "forward: src / src_mask missing");
        GCT_REQUIRE(nc == 0 || io.econds, "forward: econds missing");
        GCT_REQUIRE(Se <= 200, "encoder length %d exceeds the 200-row positional table", Se);
        GCT_CUDA(launch_k(embed_pe_kernel, dim3(Me), dim3(128), (size_t)(0), st, true, io.src, A.S, m.P(GCT_SLOT_ENC_EMB), m.c.src_vocab, io.econds,
                                            nc ? m.P(GCT_SLOT_ENC_C2E_W) : nullptr, nc ? m.P(GCT_SLOT_ENC_C2E_B) : nullptr, nc,
                                            m.P(GCT_SLOT_ENC_PE), A.x0, d, sqd, m.site(S_ENC_PE)));
        GCT_LAUNCH_CHECK();
        const float* xin = A.x0;
        for (int l = 0; l < N; ++l) {
            EncLayerAct<T>& a = A.enc[l];
            const uint32_t sb = S_ENC_BASE + l * ES_COUNT;
            GCT_TRY(m.norm_fwd(xin, m.enc_slot(l, E_N1A), m.enc_slot(l, E_N1B), a.a1, a.a1_32, Me));
            GCT_TRY(m.linear_T(a.a1, Me, d, m.enc_slot(l, E_QKV_W), m.enc_slot(l, E_QKV_B), 3 * d, a.qkv));
            float* probs = io.enc_attn ? io.enc_attn + (size_t)l * B * m.H * Se * Se : nullptr;
            GCT_TRY(m.attention(a.qkv, 3 * d, a.qkv + d, a.qkv + 2 * d, 3 * d, io.src_mask, Se, 0, a.att, a.lse, probs, B, Se,
                                Se, m.site(sb + ES_ATTN)));
            // x1 = a1 + drop(att Wo^T + bo) (residual on the normalised stream, layers.py:23-28) ; a2 = Norm2(x1)
            GCT_TRY(m.linear_res_norm(a.att, Me, d, m.WT(m.enc_slot(l, E_O_W)), m.P(m.enc_slot(l, E_O_B)), a.a1_32, a.x1,
                                      m.site(sb + ES_DROP1), m.enc_slot(l, E_N2A), m.enc_slot(l, E_N2B), a.a2, a.a2_32));
            {   // g = drop(gelu(a2 W1^T + b1))
                Epilogue e = Model<T>::epi(m.P(m.enc_slot(l, E_F1_B)), dff);
                e.flags = g_gct_ffn_classic ? EPI_GELU : (EPI_GELU | EPI_GELU_GRAD); e.aux_out = a.hpre; e.outT = a.g; e.drop = m.site(sb + ES_FF);
                GCT_TRY(m.gemm(a.a2, false, d, m.WT(m.enc_slot(l, E_F1_W)), false, d, Me, dff, d, e));
            }
            {   // xout = a2 + drop(g W2^T + b2)
                Epilogue e = Model<T>::epi(m.P(m.enc_slot(l, E_F2_B)), d);
                e.res32 = a.a2_32; e.out32 = a.xout; e.drop = m.site(sb + ES_DROP2);
                GCT_TRY(m.gemm(a.g, false, dff, m.WT(m.enc_slot(l, E_F2_W)), false, dff, Me, d, dff, e));
            }
            xin = a.xout;
        }
        GCT_TRY(m.norm_fwd(xin, GCT_SLOT_ENC_NORM_A, GCT_SLOT_ENC_NORM_B, A.xe, nullptr, Me));
        {   // mu | log_var heads as one N = 2*lat GEMM (fp32 output)
            Epilogue e = Model<T>::epi(m.P(GCT_SLOT_MULV_B), 2 * lat);
            e.out32 = A.mulv;
            GCT_TRY(m.gemm(A.xe, false, d, m.WT(GCT_SLOT_MULV_W), false, d, Me, 2 * lat, d, e));
        }
        GCT_REQUIRE(io.mu && io.log_var && io.z, "forward: mu/log_var/z outputs missing");
        const size_t n = (size_t)Me * lat;
        GCT_CUDA(launch_k(reparam_fwd_kernel<T>, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), st, true, A.mulv, io.eps, Me, lat, Se, Sm, io.mu, io.log_var, io.z,
                                                            io.run_decoder ? A.zpad : nullptr));
        GCT_LAUNCH_CHECK();
    }
    if (!io.run_decoder) return GCT_OK;

    GCT_REQUIRE(io.trg && io.trg_mask && io.logits, "forward: trg / trg_mask / logits missing");
    GCT_REQUIRE(Ld <= 200, "decoder length %d exceeds the 200-row positional table", Ld);
    if (!io.run_encoder) {
        GCT_REQUIRE(io.z_in && io.src_mask, "decode: z_in / src_mask missing");
        const size_t n = (size_t)Mm * lat;
        GCT_CUDA(launch_k(zpad_kernel<T>, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), st, true, io.z_in, B, Se, Sm, lat, A.zpad));
        GCT_LAUNCH_CHECK();
    } else if (Sm > Se) {
        GCT_CUDA(launch_k(zero_rows_kernel<T>, dim3(cdiv((size_t)B * (Sm - Se) * lat, 256)), dim3(256), (size_t)(0), st, true, A.zpad, B, Sm, Sm - Se, lat));
        GCT_LAUNCH_CHECK();
    }
    GCT_CUDA(launch_k(cross_mask_kernel, dim3(cdiv(B * Sm, 256)), dim3(256), (size_t)(0), st, true, io.src_mask, B, Se, Sm, A.cross_mask));
    GCT_LAUNCH_CHECK();
    GCT_TRY(m.linear_T(A.zpad, Mm, lat, GCT_SLOT_FCZ_W, GCT_SLOT_FCZ_B, d, A.mem));
    const bool c2d = m.c.use_cond2dec && nc > 0, c2l = m.c.use_cond2lat && nc > 0 && !c2d;
    if (Sm > Se) {
        // rows [0, nc) of the memory: cond2lat tokens when that path is active (cvaetf.py:107-110).  With
        // cond2dec AND cond2lat set the reference still extends src_mask by nc ones (cvaetf.py:114-116) but
        // does not prepend tokens -- that combination is rejected by the host wrapper.
        GCT_REQUIRE(c2l && io.dconds, "cond2lat memory rows need dconds");
        GCT_CUDA(launch_k(cond_tokens_kernel<T>, dim3(B * nc), dim3(128), (size_t)(0), st, true, io.dconds, m.P(GCT_SLOT_DEC_C2L_W), m.P(GCT_SLOT_DEC_C2L_B), nc, d, A.mem, Sm));
        GCT_LAUNCH_CHECK();
    }
    GCT_REQUIRE(!c2d || io.dconds, "cond2dec needs dconds");
    GCT_CUDA(launch_k(embed_pe_kernel, dim3(Md), dim3(128), (size_t)(0), st, true, io.trg, A.Tt, m.P(GCT_SLOT_DEC_EMB), m.c.trg_vocab, io.dconds,
                                        c2d ? m.P(GCT_SLOT_DEC_C2D_W) : nullptr, c2d ? m.P(GCT_SLOT_DEC_C2D_B) : nullptr,
                                        c2d ? nc : 0, m.P(GCT_SLOT_DEC_PE), A.y0, d, sqd, m.site(S_DEC_PE)));
    GCT_LAUNCH_CHECK();
    const float* yin = A.y0;
    for (int l = 0; l < N; ++l) {
        DecLayerAct<T>& a = A.dec[l];
        const uint32_t sb = S_DEC_BASE + l * DS_COUNT;
        GCT_TRY(m.norm_fwd(yin, m.dec_slot(l, D_N1A), m.dec_slot(l, D_N1B), a.a1, nullptr, Md));
        GCT_TRY(m.linear_T(a.a1, Md, d, m.dec_slot(l, D_QKV_W), m.dec_slot(l, D_QKV_B), 3 * d, a.qkv));
        float* pr1 = io.dec_attn1 ? io.dec_attn1 + (size_t)l * B * m.H * Ld * Ld : nullptr;
        GCT_TRY(m.attention(a.qkv, 3 * d, a.qkv + d, a.qkv + 2 * d, 3 * d, io.trg_mask, (long long)Ld * Ld, Ld, a.att1, a.lse1,
                            pr1, B, Ld, Ld, m.site(sb + DS_ATTN1)));
        GCT_TRY(m.linear_res_norm(a.att1, Md, d, m.WT(m.dec_slot(l, D_O1_W)), m.P(m.dec_slot(l, D_O1_B)), yin, a.y1,
                                  m.site(sb + DS_DROP1), m.dec_slot(l, D_N2A), m.dec_slot(l, D_N2B), a.a2, nullptr));
        GCT_TRY(m.linear_T(a.a2, Md, d, m.dec_slot(l, D_Q2_W), m.dec_slot(l, D_Q2_B), d, a.q2));
        GCT_TRY(m.linear_T(A.mem, Mm, d, m.dec_slot(l, D_KV2_W), m.dec_slot(l, D_KV2_B), 2 * d, a.kv2));
        float* pr2 = io.dec_attn2 ? io.dec_attn2 + (size_t)l * B * m.H * Ld * Sm : nullptr;
        GCT_TRY(m.attention(a.q2, d, a.kv2, a.kv2 + d, 2 * d, A.cross_mask, Sm, 0, a.att2, a.lse2, pr2, B, Ld, Sm,
                            m.site(sb + DS_ATTN2)));
        GCT_TRY(m.linear_res_norm(a.att2, Md, d, m.WT(m.dec_slot(l, D_O2_W)), m.P(m.dec_slot(l, D_O2_B)), a.y1, a.y2,
                                  m.site(sb + DS_DROP2), m.dec_slot(l, D_N3A), m.dec_slot(l, D_N3B), a.a3, nullptr));
        {
            Epilogue e = Model<T>::epi(m.P(m.dec_slot(l, D_F1_B)), dff);
            e.flags = g_gct_ffn_classic ? EPI_GELU : (EPI_GELU | EPI_GELU_GRAD); e.aux_out = a.hpre; e.outT = a.g; e.drop = m.site(sb + DS_FF);
            GCT_TRY(m.gemm(a.a3, false, d, m.WT(m.dec_slot(l, D_F1_W)), false, d, Md, dff, d, e));
        }
        {
            Epilogue e = Model<T>::epi(m.P(m.dec_slot(l, D_F2_B)), d);
            e.res32 = a.y2; e.out32 = a.yout; e.drop = m.site(sb + DS_DROP3);
            GCT_TRY(m.gemm(a.g, false, dff, m.WT(m.dec_slot(l, D_F2_W)), false, dff, Md, d, dff, e));
        }
        yin = a.yout;
    }
    GCT_TRY(m.norm_fwd(yin, GCT_SLOT_DEC_NORM_A, GCT_SLOT_DEC_NORM_B, A.yd, nullptr, Md));
    {
        Epilogue e = Model<T>::epi(m.P(GCT_SLOT_OUT_B), m.c.trg_vocab);
        e.out32 = io.logits;
        GCT_TRY(m.gemm(A.yd, false, d, m.WT(GCT_SLOT_OUT_W), false, d, Md, m.c.trg_vocab, d, e));
    }
    return GCT_OK;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <typename T>
struct BwdScratch {
    float *dxa, *dxb, *dxc;       // fp32 [Mmax, d] ping-pong gradients of the residual stream
    float* dmem;                  // fp32 [Mm, d]
    T *dyT, *dattT, *dqkvT, *dhT; // operand-typed gradient buffers
    T* dkv2T; T* dmemT; T* dlogT; T* dmulvT;
    float* dzpad; float* dz;
    size_t bytes;
    void carve(const gct_config_t& c, const Acts<T>& A, void* ws) {
        Bump bp(ws);
        const int d = c.d_model, dff = c.d_ff, lat = c.latent_dim;
        const size_t Mmax = (size_t)((A.Me > A.Md ? A.Me : A.Md) > A.Mm ? (A.Me > A.Md ? A.Me : A.Md) : A.Mm);
        dxa = bp.arr<float>(Mmax * d); dxb = bp.arr<float>(Mmax * d); dxc = bp.arr<float>(Mmax * d);
        dmem = bp.arr<float>((size_t)A.Mm * d);
        dyT = bp.arr<T>(Mmax * d); dattT = bp.arr<T>(Mmax * d); dqkvT = bp.arr<T>(Mmax * 3 * d); dhT = bp.arr<T>(Mmax * dff);
        dkv2T = bp.arr<T>((size_t)A.Mm * 2 * d); dmemT = bp.arr<T>((size_t)A.Mm * d);
        dlogT = bp.arr<T>((size_t)A.Md * A.Vpad); dmulvT = bp.arr<T>((size_t)A.Me * 2 * lat);
        dzpad = bp.arr<float>((size_t)A.Mm * lat); dz = bp.arr<float>((size_t)A.Me * lat);
        bytes = bp.off + 256;
    }
};

// FFN block backward shared by encoder and decoder layers.
//   forward:  pre = a W1^T + b1 ; g = dropF(gelu(pre)) ; hpre := keepF*gelu'(pre) ; out = res + dropO(g W2^T + b2)
//   in: dout (fp32 grad of out).  out: dA (fp32) = dHpre W1 (+ add_to_dA)
template <typename T>
static int ffn_backward(Model<T>& m, BwdScratch<T>& S, int M, const float* dout, const T* a, const T* hpre, const T* g,
                        int f1w, int f1b, int f2w, int f2b, DropCtx drop_out, DropCtx drop_ff, const float* add_to_dA,
                        float* dA, bool dyT_ready = false) {
    const int d = m.d, dff = m.dff;
    // dyT_ready: the Norm backward that produced dout already wrote S.dyT = dropO'(dout) and the b2 gradient
    if (!dyT_ready) GCT_TRY(m.cast_drop(dout, S.dyT, M, d, drop_out, m.G(f2b)));
    GCT_TRY(m.wgrad(S.dyT, d, g, dff, M, d, dff, f2w, f2b, false));
    {   // dHpre = (dY W2) * [keepF * gelu'(pre)]  -- the bracket was saved by the forward epilogue (EPI_GELU_GRAD)
        Epilogue e = Model<T>::epi(nullptr, dff);
        e.aux_in = hpre; e.outT = S.dhT;
        if (g_gct_ffn_classic) { e.flags = EPI_DGELU; e.drop = drop_ff; }      // hpre = pre-activation: regenerate mask, gelu'
        else e.flags = EPI_MUL_AUX;                                             // hpre = keepF * gelu'(pre)
        GCT_TRY(m.gemm(S.dyT, false, d, m.WT(f2w), true, dff, M, dff, d, e));
    }
    GCT_TRY(m.wgrad(S.dhT, dff, a, d, M, dff, d, f1w, f1b, true));
    {
        Epilogue e = Model<T>::epi(nullptr, d);
        e.res32 = add_to_dA; e.out32 = dA;
        GCT_TRY(m.gemm(S.dhT, false, dff, m.WT(f1w), true, d, M, d, dff, e));
    }
    return GCT_OK;
}

// Called by model_backward when a group of gradients is final: stage k < N = decoder layer N-1-k, stage N+k = encoder layer
// N-1-k, stage 2N = everything (embeddings, heads, final norms).  The data-parallel backward (gct_backward_dp) uses it to
// start the gradient exchange of that group on the communication stream while the rest of the backward runs.
struct StageHook {
    virtual int done(int stage) = 0;
    virtual ~StageHook() {}
};

template <typename T>
static int model_backward(Model<T>& m, const gct_io_t& io, Acts<T>& A, BwdScratch<T>& S, const float* dlogits,
                          const float* dmu, const float* dlv, const float* dz_ext, StageHook* hook = nullptr) {
    const int d = m.d, lat = m.lat, nc = m.nc, N = m.N;
    const int B = A.B, Se = A.Se, Sm = A.Sm, Ld = A.Ld, Me = A.Me, Mm = A.Mm, Md = A.Md, V = m.c.trg_vocab, Vpad = A.Vpad;
    cudaStream_t st = m.st;
    const float sqd = sqrtf((float)d);
    GCT_REQUIRE(m.w.grads_f32, "backward: gradient buffer missing");
    const bool c2d = m.c.use_cond2dec && nc > 0, c2l = m.c.use_cond2lat && nc > 0 && !c2d;
    const float* dz_from_dec = nullptr;

    if (io.run_decoder && dlogits) {
        // ---- vocabulary projection ----
        GCT_CUDA(launch_k(pad_cast_kernel<T>, dim3(cdiv((size_t)Md * Vpad, 256)), dim3(256), (size_t)(0), st, true, dlogits, Md, V, Vpad, S.dlogT));
        GCT_LAUNCH_CHECK();
        GCT_TRY(m.wgrad(S.dlogT, Vpad, A.yd, d, Md, V, d, GCT_SLOT_OUT_W, GCT_SLOT_OUT_B, true));
        {
            Epilogue e = Model<T>::epi(nullptr, d); e.out32 = S.dxa;
            GCT_TRY(m.gemm(S.dlogT, false, Vpad, m.WT(GCT_SLOT_OUT_W), true, d, Md, d, V, e));
        }
        const float* ylast = A.dec[N - 1].yout;
        float* dy = S.dxb;          // gradient of the residual stream entering the final norm
        // every Norm backward below also emits S.dyT = T(dropout'(dx)) + the bias gradient for the projection that
        // consumes it next (FFN linear_2 / out-projections), replacing a separate cast pass over the fp32 gradient
        GCT_TRY(m.norm_bwd(ylast, GCT_SLOT_DEC_NORM_A, GCT_SLOT_DEC_NORM_B, S.dxa, nullptr, dy, Md, S.dyT,
                           m.site(S_DEC_BASE + (N - 1) * DS_COUNT + DS_DROP3), m.G(m.dec_slot(N - 1, D_F2_B))));
        GCT_CUDA(cudaMemsetAsync(S.dmem, 0, (size_t)Mm * d * sizeof(float), st));
        float* other = S.dxa;       // free buffer
        float* third = S.dxc;
        for (int l = N - 1; l >= 0; --l) {
            DecLayerAct<T>& a = A.dec[l];
            const float* yin = (l == 0) ? A.y0 : A.dec[l - 1].yout;
            const uint32_t sb = S_DEC_BASE + l * DS_COUNT;
            // FFN: dA3 -> other ; dY2 = norm3_bwd(y2, dA3) + dy -> third
            GCT_TRY(ffn_backward(m, S, Md, dy, a.a3, a.hpre, a.g, m.dec_slot(l, D_F1_W), m.dec_slot(l, D_F1_B),
                                 m.dec_slot(l, D_F2_W), m.dec_slot(l, D_F2_B), m.site(sb + DS_DROP3), m.site(sb + DS_FF), nullptr,
                                 other, true));
            GCT_TRY(m.norm_bwd(a.y2, m.dec_slot(l, D_N3A), m.dec_slot(l, D_N3B), other, dy, third, Md, S.dyT, m.site(sb + DS_DROP2),
                               m.G(m.dec_slot(l, D_O2_B))));
            // cross attention
            float* dY2 = third;
            GCT_TRY(m.wgrad(S.dyT, d, a.att2, d, Md, d, d, m.dec_slot(l, D_O2_W), m.dec_slot(l, D_O2_B), false));
            {
                Epilogue e = Model<T>::epi(nullptr, d); e.outT = S.dattT;
                GCT_TRY(m.gemm(S.dyT, false, d, m.WT(m.dec_slot(l, D_O2_W)), true, d, Md, d, d, e));
            }
            T* dq2 = S.dqkvT;                      // [Md, d]
            GCT_TRY(m.attention_bwd(a.q2, d, a.kv2, a.kv2 + d, 2 * d, A.cross_mask, Sm, 0, a.lse2, a.att2, S.dattT, dq2, d, S.dkv2T,
                                    S.dkv2T + d, 2 * d, B, Ld, Sm, m.site(sb + DS_ATTN2), m.G(m.dec_slot(l, D_Q2_B)),
                                    m.G(m.dec_slot(l, D_KV2_B)), m.G(m.dec_slot(l, D_KV2_B)) + d));
            GCT_TRY(m.wgrad(dq2, d, a.a2, d, Md, d, d, m.dec_slot(l, D_Q2_W), m.dec_slot(l, D_Q2_B), false));
            GCT_TRY(m.wgrad(S.dkv2T, 2 * d, A.mem, d, Mm, 2 * d, d, m.dec_slot(l, D_KV2_W), m.dec_slot(l, D_KV2_B), false));
            {   // dmem += dkv2 Wkv2
                Epilogue e = Model<T>::epi(nullptr, d); e.res32 = S.dmem; e.out32 = S.dmem;
                GCT_TRY(m.gemm(S.dkv2T, false, 2 * d, m.WT(m.dec_slot(l, D_KV2_W)), true, d, Mm, d, 2 * d, e));
            }
            {   // dA2 = dq2 Wq2 -> other
                Epilogue e = Model<T>::epi(nullptr, d); e.out32 = other;
                GCT_TRY(m.gemm(dq2, false, d, m.WT(m.dec_slot(l, D_Q2_W)), true, d, Md, d, d, e));
            }
            // dY1 = norm2_bwd(y1, dA2) + dY2 -> dy
            GCT_TRY(m.norm_bwd(a.y1, m.dec_slot(l, D_N2A), m.dec_slot(l, D_N2B), other, dY2, dy, Md, S.dyT, m.site(sb + DS_DROP1),
                               m.G(m.dec_slot(l, D_O1_B))));
            // self attention
            GCT_TRY(m.wgrad(S.dyT, d, a.att1, d, Md, d, d, m.dec_slot(l, D_O1_W), m.dec_slot(l, D_O1_B), false));
            {
                Epilogue e = Model<T>::epi(nullptr, d); e.outT = S.dattT;
                GCT_TRY(m.gemm(S.dyT, false, d, m.WT(m.dec_slot(l, D_O1_W)), true, d, Md, d, d, e));
            }
            GCT_TRY(m.attention_bwd(a.qkv, 3 * d, a.qkv + d, a.qkv + 2 * d, 3 * d, io.trg_mask, (long long)Ld * Ld, Ld, a.lse1,
                                    a.att1, S.dattT, S.dqkvT, 3 * d, S.dqkvT + d, S.dqkvT + 2 * d, 3 * d, B, Ld, Ld, m.site(sb + DS_ATTN1),
                                    m.G(m.dec_slot(l, D_QKV_B)), m.G(m.dec_slot(l, D_QKV_B)) + d, m.G(m.dec_slot(l, D_QKV_B)) + 2 * d));
            GCT_TRY(m.wgrad(S.dqkvT, 3 * d, a.a1, d, Md, 3 * d, d, m.dec_slot(l, D_QKV_W), m.dec_slot(l, D_QKV_B), false));
            {
                Epilogue e = Model<T>::epi(nullptr, d); e.out32 = other;
                GCT_TRY(m.gemm(S.dqkvT, false, 3 * d, m.WT(m.dec_slot(l, D_QKV_W)), true, d, Md, d, 3 * d, e));
            }
            // dYin = norm1_bwd(yin, dA1) + dY1 -> third ; rotate
            if (l > 0)
                GCT_TRY(m.norm_bwd(yin, m.dec_slot(l, D_N1A), m.dec_slot(l, D_N1B), other, dy, third, Md, S.dyT,
                                   m.site(S_DEC_BASE + (l - 1) * DS_COUNT + DS_DROP3), m.G(m.dec_slot(l - 1, D_F2_B))));
            else
                GCT_TRY(m.norm_bwd(yin, m.dec_slot(l, D_N1A), m.dec_slot(l, D_N1B), other, dy, third, Md));
            float* t = dy; dy = third; third = t;
            if (hook) GCT_TRY(hook->done(N - 1 - l));
        }
        // decoder embedding (+ cond2dec tokens)
        {
            dim3 grid(cdiv((long long)B * A.Tt, EMB_BWD_ROWS), cdiv(d, 128));
            GCT_CUDA(launch_k(embed_bwd_kernel, dim3(grid), dim3(128), (size_t)((size_t)m.c.trg_vocab * 128 * sizeof(float)), st, true, io.trg, B, A.Tt, c2d ? nc : 0, dy, d, sqd, m.site(S_DEC_PE),
                                                   m.G(GCT_SLOT_DEC_EMB), m.c.trg_vocab));
            GCT_LAUNCH_CHECK();
            if (c2d) {
                dim3 g2(nc, cdiv(d, 128), cdiv(B, COND_BWD_BATCH));
                GCT_CUDA(launch_k(cond_embed_bwd_kernel, dim3(g2), dim3(128), (size_t)(0), st, true, dy, B, Ld, nc, d, io.dconds, sqd, m.site(S_DEC_PE), 1,
                                                          m.G(GCT_SLOT_DEC_C2D_W), m.G(GCT_SLOT_DEC_C2D_B)));
                GCT_LAUNCH_CHECK();
            }
        }
        // memory: cond2lat tokens, fc_z
        if (c2l) {
            dim3 g2(nc, cdiv(d, 128), cdiv(B, COND_BWD_BATCH));
            GCT_CUDA(launch_k(cond_embed_bwd_kernel, dim3(g2), dim3(128), (size_t)(0), st, true, S.dmem, B, Sm, nc, d, io.dconds, 1.f, m.site(0), 0,
                                                      m.G(GCT_SLOT_DEC_C2L_W), m.G(GCT_SLOT_DEC_C2L_B)));
            GCT_LAUNCH_CHECK();
        }
        DropCtx nodrop; nodrop.seed = 0; nodrop.thresh = 0; nodrop.scale = 1.f;
        GCT_TRY(m.cast_drop(S.dmem, S.dmemT, Mm, d, nodrop, nullptr));
        if (Sm > Se) {
            GCT_CUDA(launch_k(zero_rows_kernel<T>, dim3(cdiv((size_t)B * (Sm - Se) * d, 256)), dim3(256), (size_t)(0), st, true, S.dmemT, B, Sm, Sm - Se, d));
            GCT_LAUNCH_CHECK();
        }
        GCT_TRY(m.wgrad(S.dmemT, d, A.zpad, lat, Mm, d, lat, GCT_SLOT_FCZ_W, GCT_SLOT_FCZ_B, true));
        {
            Epilogue e = Model<T>::epi(nullptr, lat); e.out32 = S.dzpad;
            GCT_TRY(m.gemm(S.dmemT, false, d, m.WT(GCT_SLOT_FCZ_W), true, lat, Mm, lat, d, e));
        }
        GCT_CUDA(launch_k(dz_gather_kernel, dim3(cdiv((size_t)Me * lat, 256)), dim3(256), (size_t)(0), st, true, S.dzpad, dz_ext, B, Se, Sm, lat, S.dz));
        GCT_LAUNCH_CHECK();
        dz_from_dec = S.dz;
    } else {
        dz_from_dec = dz_ext;
    }
    if (!io.run_encoder) return GCT_OK;
    if (!dz_from_dec && !dmu && !dlv) return GCT_OK;

    // ---- latent heads ----
    GCT_CUDA(launch_k(reparam_bwd_kernel<T>, dim3(cdiv((size_t)Me * lat, 256)), dim3(256), (size_t)(0), st, true, dz_from_dec, dmu, dlv, io.eps, io.log_var, Me, lat, S.dmulvT));
    GCT_LAUNCH_CHECK();
    GCT_TRY(m.wgrad(S.dmulvT, 2 * lat, A.xe, d, Me, 2 * lat, d, GCT_SLOT_MULV_W, GCT_SLOT_MULV_B, true));
    {
        Epilogue e = Model<T>::epi(nullptr, d); e.out32 = S.dxa;
        GCT_TRY(m.gemm(S.dmulvT, false, 2 * lat, m.WT(GCT_SLOT_MULV_W), true, d, Me, d, 2 * lat, e));
    }
    float* dx = S.dxb;
    GCT_TRY(m.norm_bwd(A.enc[N - 1].xout, GCT_SLOT_ENC_NORM_A, GCT_SLOT_ENC_NORM_B, S.dxa, nullptr, dx, Me, S.dyT,
                       m.site(S_ENC_BASE + (N - 1) * ES_COUNT + ES_DROP2), m.G(m.enc_slot(N - 1, E_F2_B))));
    float* other = S.dxa;
    float* third = S.dxc;
    for (int l = N - 1; l >= 0; --l) {
        EncLayerAct<T>& a = A.enc[l];
        const float* xin = (l == 0) ? A.x0 : A.enc[l - 1].xout;
        const uint32_t sb = S_ENC_BASE + l * ES_COUNT;
        // xout = a2_32 + drop2(ffn(a2)) : dA2 = dHpre W1 + dx (residual on the normalised stream)
        GCT_TRY(ffn_backward(m, S, Me, dx, a.a2, a.hpre, a.g, m.enc_slot(l, E_F1_W), m.enc_slot(l, E_F1_B), m.enc_slot(l, E_F2_W),
                             m.enc_slot(l, E_F2_B), m.site(sb + ES_DROP2), m.site(sb + ES_FF), dx, other, true));
        GCT_TRY(m.norm_bwd(a.x1, m.enc_slot(l, E_N2A), m.enc_slot(l, E_N2B), other, nullptr, third, Me, S.dyT, m.site(sb + ES_DROP1),
                           m.G(m.enc_slot(l, E_O_B))));   // dX1
        GCT_TRY(m.wgrad(S.dyT, d, a.att, d, Me, d, d, m.enc_slot(l, E_O_W), m.enc_slot(l, E_O_B), false));
        {
            Epilogue e = Model<T>::epi(nullptr, d); e.outT = S.dattT;
            GCT_TRY(m.gemm(S.dyT, false, d, m.WT(m.enc_slot(l, E_O_W)), true, d, Me, d, d, e));
        }
        GCT_TRY(m.attention_bwd(a.qkv, 3 * d, a.qkv + d, a.qkv + 2 * d, 3 * d, io.src_mask, Se, 0, a.lse, a.att, S.dattT, S.dqkvT, 3 * d,
                                S.dqkvT + d, S.dqkvT + 2 * d, 3 * d, B, Se, Se, m.site(sb + ES_ATTN), m.G(m.enc_slot(l, E_QKV_B)),
                                m.G(m.enc_slot(l, E_QKV_B)) + d, m.G(m.enc_slot(l, E_QKV_B)) + 2 * d));
        GCT_TRY(m.wgrad(S.dqkvT, 3 * d, a.a1, d, Me, 3 * d, d, m.enc_slot(l, E_QKV_W), m.enc_slot(l, E_QKV_B), false));
        {   // dA1 = dqkv Wqkv + dX1
            Epilogue e = Model<T>::epi(nullptr, d); e.res32 = third; e.out32 = other;
            GCT_TRY(m.gemm(S.dqkvT, false, 3 * d, m.WT(m.enc_slot(l, E_QKV_W)), true, d, Me, d, 3 * d, e));
        }
        if (l > 0)
            GCT_TRY(m.norm_bwd(xin, m.enc_slot(l, E_N1A), m.enc_slot(l, E_N1B), other, nullptr, dx, Me, S.dyT,
                               m.site(S_ENC_BASE + (l - 1) * ES_COUNT + ES_DROP2), m.G(m.enc_slot(l - 1, E_F2_B))));
        else
            GCT_TRY(m.norm_bwd(xin, m.enc_slot(l, E_N1A), m.enc_slot(l, E_N1B), other, nullptr, dx, Me));
        if (hook) GCT_TRY(hook->done(N + N - 1 - l));
    }
    {
        dim3 grid(cdiv((long long)B * A.S, EMB_BWD_ROWS), cdiv(d, 128));
        GCT_CUDA(launch_k(embed_bwd_kernel, dim3(grid), dim3(128), (size_t)((size_t)m.c.src_vocab * 128 * sizeof(float)), st, true, io.src, B, A.S, nc, dx, d, sqd, m.site(S_ENC_PE), m.G(GCT_SLOT_ENC_EMB), m.c.src_vocab));
        GCT_LAUNCH_CHECK();
        if (nc > 0) {
            dim3 g2(nc, cdiv(d, 128), cdiv(B, COND_BWD_BATCH));
            GCT_CUDA(launch_k(cond_embed_bwd_kernel, dim3(g2), dim3(128), (size_t)(0), st, true, dx, B, Se, nc, d, io.econds, sqd, m.site(S_ENC_PE), 1,
                                                      m.G(GCT_SLOT_ENC_C2E_W), m.G(GCT_SLOT_ENC_C2E_B)));
            GCT_LAUNCH_CHECK();
        }
    }
    return GCT_OK;
}

// ------------------------------------------------------------------------------------------
// KV-cached decoder
// ------------------------------------------------------------------------------------------
template <typename T>
struct DecodeWs {
    int B, Lz, Sm, Lmax;
    bool zmode;                     // cross-attention in latent space (decode_zattn.cuh): no per-layer K/V of the latent rows
    int nck;                        // condition rows of the memory (use_cond2lat), 0 otherwise
    int KZ;                         // width of the latent-space query / context vectors: H*lat (+ d with condition rows)
    T* zpad; T* mem; T* kx; T* vx;  // K/V form: zpad [B][Sm][lat] (cond rows zero), memory, cross keys / values [N][B*Sm][d]
    T* kc; T* vc;                   // [N][B][Lmax][d]
    float* x; T* xn; T* qkv; T* att; T* q2; T* hbuf; float* logits;
    uint8_t* key_valid; uint8_t* cross_mask; uint8_t* done;
    // latent-space form: zlat [B][Lz][lat]; folded projections (per layer) and their scratch
    T* zlat;
    T* wqz; float* bqz; T* woz; float* boz;      // [N][KZ][d], [N][KZ], [N][d][KZ], [N][d]
    float* mk; float* mv; float* tvec; T* wtmp;  // (Wk Wz) [d][lat], (Wv Wz) [d][lat], Wv bz + bv [d], Woz in head-major order [d][H*lat]
    T* ucond; T* kvc;                            // condition rows minus bz [B*nc][d]; their per-layer (K | V) [N][B*nc][2d]
    T* qz; T* zbar;                              // [B][KZ]
    size_t bytes;
    static bool want_zmode(const gct_config_t& c, int Lz_) {
        (void)Lz_;
        const int nck_ = (c.use_cond2lat && c.nconds > 0 && !c.use_cond2dec) ? c.nconds : 0;
        return sizeof(T) == 2 && g_gct_zattn && zattn_supported(c.latent_dim, c.heads, nck_);
    }
    void carve(const gct_config_t& c, int B_, int Lz_, int max_len, void* ws) {
        Bump bp(ws);
        B = B_; Lz = Lz_; Lmax = max_len;
        const int d = c.d_model, nc = c.nconds, N = c.n_layers, lat = c.latent_dim, HL = c.heads * c.latent_dim;
        nck = (c.use_cond2lat && nc > 0 && !(c.use_cond2dec)) ? nc : 0;
        Sm = Lz + nck;
        zmode = want_zmode(c, Lz_);
        KZ = HL + (nck ? d : 0);
        zpad = mem = kx = vx = zlat = nullptr;
        wqz = woz = qz = zbar = wtmp = ucond = kvc = nullptr; bqz = boz = mk = mv = tvec = nullptr;
        if (!zmode) {
            zpad = bp.arr<T>((size_t)B * Sm * lat);
            mem = bp.arr<T>((size_t)B * Sm * d);
            kx = bp.arr<T>((size_t)N * B * Sm * d);
            vx = bp.arr<T>((size_t)N * B * Sm * d);
        } else {
            zlat = bp.arr<T>((size_t)B * Lz * lat);
            wqz = bp.arr<T>((size_t)N * KZ * d); bqz = bp.arr<float>((size_t)N * KZ);
            woz = bp.arr<T>((size_t)N * d * KZ); boz = bp.arr<float>((size_t)N * d);
            mk = bp.arr<float>((size_t)d * lat); mv = bp.arr<float>((size_t)d * lat); tvec = bp.arr<float>(d);
            wtmp = bp.arr<T>((size_t)d * HL);
            qz = bp.arr<T>((size_t)B * KZ); zbar = bp.arr<T>((size_t)B * KZ);
            if (nck) { ucond = bp.arr<T>((size_t)B * nck * d); kvc = bp.arr<T>((size_t)N * B * nck * 2 * d); }
        }
        kc = bp.arr<T>((size_t)N * B * Lmax * d);
        vc = bp.arr<T>((size_t)N * B * Lmax * d);
        x = bp.arr<float>((size_t)B * d); xn = bp.arr<T>((size_t)B * d); qkv = bp.arr<T>((size_t)B * 3 * d);
        att = bp.arr<T>((size_t)B * d); q2 = bp.arr<T>((size_t)B * d); hbuf = bp.arr<T>((size_t)B * c.d_ff);
        logits = bp.arr<float>((size_t)B * c.trg_vocab);
        key_valid = bp.arr<uint8_t>((size_t)B * Lmax); cross_mask = bp.arr<uint8_t>((size_t)B * Sm);
        done = bp.arr<uint8_t>((size_t)B);
        bytes = bp.off + 256;
    }
};

// dst[r, c] = T(scale * src[r, c])   (row pitches in elements)
template <typename T>
__global__ void scale_copy2d_kernel(const float* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld, int rows, int cols,
                                    float scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[(size_t)r * dst_ld + c] = from_f<T>(scale * src[(size_t)r * src_ld + c]);
}
// head-major [rows][H*lat] -> (dim, head) order: dst[r, a*H + h] = src[r, h*lat + a]
template <typename T>
__global__ void head_to_dim_major_kernel(const T* __restrict__ src, T* __restrict__ dst, int dst_ld, int rows, int H, int lat) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * H * lat) return;
    const int r = (int)(i / (H * lat)), c = (int)(i % (H * lat));
    const int a = c / H, h = c % H;
    dst[(size_t)r * dst_ld + c] = src[(size_t)r * H * lat + (size_t)h * lat + a];
}
// u[b*nc + j, c] = embed_cond2lat(dconds)[b, j, c] - bz[c]   (condition rows of the memory relative to fc_z's bias)
template <typename T>
__global__ void cond_u_kernel(const float* __restrict__ conds, const float* __restrict__ W, const float* __restrict__ Bv,
                              const float* __restrict__ bz, int nc, int d, T* __restrict__ u) {
    const int b = blockIdx.x / nc, j = blockIdx.x % nc;
    T* o = u + (size_t)blockIdx.x * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float* w = W + ((size_t)j * d + c) * nc;
        float v = Bv[(size_t)j * d + c] - bz[c];
        for (int k = 0; k < nc; ++k) v = fmaf(w[k], conds[(size_t)b * nc + k], v);
        o[c] = from_f<T>(v);
    }
}

// Folds fc_z and the cross-attention k / v / out projections of every decoder layer into the two latent-space
// operands of decode_zattn.cuh.  fp32 SIMT GEMMs over the master weights (about 0.8 GFLOP, once per decode call);
// results are rounded once to the operand type.  nn.Linear weights are [out, in] row-major.
template <typename T>
static int decode_zprep(Model<T>& m, DecodeWs<T>& W) {
    const int d = m.d, lat = m.lat, N = m.N, H = m.H, HL = H * lat, KZ = W.KZ;
    cudaStream_t st = m.st;
    const float* Wz = m.P(GCT_SLOT_FCZ_W);      // [d, lat]
    const float* bz = m.P(GCT_SLOT_FCZ_B);
    for (int l = 0; l < N; ++l) {
        const float* Wq = m.P(m.dec_slot(l, D_Q2_W)); const float* bq = m.P(m.dec_slot(l, D_Q2_B));
        const float* Wk = m.P(m.dec_slot(l, D_KV2_W)); const float* Wv = Wk + (size_t)d * d;
        const float* bv = m.P(m.dec_slot(l, D_KV2_B)) + d;
        const float* Wo = m.P(m.dec_slot(l, D_O2_W)); const float* bo = m.P(m.dec_slot(l, D_O2_B));
        T* wqz = W.wqz + (size_t)l * KZ * d; float* bqz = W.bqz + (size_t)l * KZ;
        T* woz = W.woz + (size_t)l * d * KZ;
        {   // mk[c, a] = sum_m Wk[c, m] Wz[m, a] ;  mv likewise
            Epilogue e = Model<T>::epi(nullptr, lat); e.out32 = W.mk;
            GCT_TRY((launch_gemm_simt<float, float, float>(Wk, d, 1, Wz, 1, lat, d, lat, d, 1, e, st)));
            e.out32 = W.mv;
            GCT_TRY((launch_gemm_simt<float, float, float>(Wv, d, 1, Wz, 1, lat, d, lat, d, 1, e, st)));
        }
        {   // tvec = Wv bz + bv ;  boz = Wo tvec + bo
            Epilogue e = Model<T>::epi(bv, 1); e.flags = EPI_BIAS_ROW; e.out32 = W.tvec;
            GCT_TRY((launch_gemm_simt<float, float, float>(Wv, d, 1, bz, 0, 1, d, 1, d, 1, e, st)));
            Epilogue e2 = Model<T>::epi(bo, 1); e2.flags = EPI_BIAS_ROW; e2.out32 = W.boz + (size_t)l * d;
            GCT_TRY((launch_gemm_simt<float, float, float>(Wo, d, 1, W.tvec, 0, 1, d, 1, d, 1, e2, st)));
        }
        for (int h = 0; h < H; ++h) {
            const float* mkh = W.mk + (size_t)h * 64 * lat;
            const float* mvh = W.mv + (size_t)h * 64 * lat;
            {   // Wqz[h*lat + a, n] = (1/8) sum_i mk[h*64+i, a] Wq[h*64+i, n]
                Epilogue e = Model<T>::epi(nullptr, d); e.alpha = 0.125f; e.outT = wqz + (size_t)h * lat * d;
                GCT_TRY((launch_gemm_simt<float, float, T>(mkh, 1, lat, Wq + (size_t)h * 64 * d, 1, d, lat, d, 64, 1, e, st)));
                // bqz[h*lat + a] = (1/8) sum_i mk[h*64+i, a] bq[h*64+i]
                Epilogue eb = Model<T>::epi(nullptr, 1); eb.alpha = 0.125f; eb.out32 = bqz + (size_t)h * lat;
                GCT_TRY((launch_gemm_simt<float, float, float>(mkh, 1, lat, bq + h * 64, 0, 1, lat, 1, 64, 1, eb, st)));
            }
            {   // Woz (head-major scratch)[n, h*lat + a] = sum_i Wo[n, h*64+i] mv[h*64+i, a]
                Epilogue e = Model<T>::epi(nullptr, HL); e.outT = W.wtmp + (size_t)h * lat;
                GCT_TRY((launch_gemm_simt<float, float, T>(Wo + h * 64, d, 1, mvh, 1, lat, d, lat, 64, 1, e, st)));
            }
        }
        // the kernel writes zbar in (dim, head) order: permute the contraction index of Woz to match
        head_to_dim_major_kernel<T><<<cdiv((size_t)d * HL, 256), 256, 0, st>>>(W.wtmp, woz, KZ, d, H, lat);
        GCT_LAUNCH_CHECK();
        if (W.nck) {
            // raw queries / 8 as extra output columns of the q GEMM, Wo as extra contraction columns of the out GEMM
            scale_copy2d_kernel<T><<<cdiv((size_t)d * d, 256), 256, 0, st>>>(Wq, d, wqz + (size_t)HL * d, d, d, d, 0.125f);
            GCT_LAUNCH_CHECK();
            scale_copy2d_kernel<float><<<cdiv(d, 256), 256, 0, st>>>(bq, d, bqz + HL, d, 1, d, 0.125f);
            GCT_LAUNCH_CHECK();
            scale_copy2d_kernel<T><<<cdiv((size_t)d * d, 256), 256, 0, st>>>(Wo, d, woz + HL, KZ, d, d, 1.f);
            GCT_LAUNCH_CHECK();
            // (K | V) of the condition rows: u [B*nc, d] x (Wk ; Wv)^T, no bias (the constants live in boz / cancel in the softmax)
            Epilogue e = Model<T>::epi(nullptr, 2 * d); e.outT = W.kvc + (size_t)l * W.B * W.nck * 2 * d;
            GCT_TRY(m.gemm(W.ucond, false, d, m.WT(m.dec_slot(l, D_KV2_W)), false, d, W.B * W.nck, 2 * d, d, e));
        }
    }
    return GCT_OK;
}

template <typename T>
static int decode_begin(Model<T>& m, const gct_decode_t& D, DecodeWs<T>& W) {
    const int d = m.d, lat = m.lat, nc = m.nc, N = m.N, B = W.B, Sm = W.Sm, Lz = W.Lz;
    cudaStream_t st = m.st;
    GCT_REQUIRE(!(m.c.use_cond2dec && nc > 0), "KV-cached decode does not cover use_cond2dec (host falls back to re-decode)");
    GCT_REQUIRE(D.zs && D.src_mask && D.ys && D.status, "decode: zs / src_mask / ys / status missing");
    GCT_REQUIRE(D.prefix_len >= 1 && D.prefix_len <= D.max_len, "decode: bad prefix length");
    GCT_REQUIRE(D.max_len <= 200 && Sm <= DEC_MAX_KEYS, "decode: max_len %d > 200 or memory length %d > %d", D.max_len, Sm, DEC_MAX_KEYS);
    GCT_CUDA(launch_k(cross_mask_kernel, dim3(cdiv(B * Sm, 256)), dim3(256), (size_t)(0), st, true, D.src_mask, B, Lz, Sm, W.cross_mask));
    GCT_LAUNCH_CHECK();
    if (W.zmode) {
        GCT_CUDA(launch_k(zpad_kernel<T>, dim3(cdiv((size_t)B * Lz * lat, 256)), dim3(256), (size_t)(0), st, true, D.zs, B, Lz, Lz, lat, W.zlat));      // plain cast: latent rows only
        GCT_LAUNCH_CHECK();
        if (W.nck) {
            GCT_REQUIRE(D.dconds, "decode: dconds missing");
            cond_u_kernel<T><<<B * nc, 128, 0, st>>>(D.dconds, m.P(GCT_SLOT_DEC_C2L_W), m.P(GCT_SLOT_DEC_C2L_B), m.P(GCT_SLOT_FCZ_B), nc, d, W.ucond);
            GCT_LAUNCH_CHECK();
        }
        GCT_TRY(decode_zprep(m, W));
    } else {
        GCT_CUDA(launch_k(zpad_kernel<T>, dim3(cdiv((size_t)B * Sm * lat, 256)), dim3(256), (size_t)(0), st, true, D.zs, B, Lz, Sm, lat, W.zpad));
        GCT_LAUNCH_CHECK();
        GCT_TRY(m.linear_T(W.zpad, B * Sm, lat, GCT_SLOT_FCZ_W, GCT_SLOT_FCZ_B, d, W.mem));
        if (Sm > Lz) {
            GCT_REQUIRE(D.dconds, "decode: dconds missing");
            GCT_CUDA(launch_k(cond_tokens_kernel<T>, dim3(B * nc), dim3(128), (size_t)(0), st, true, D.dconds, m.P(GCT_SLOT_DEC_C2L_W), m.P(GCT_SLOT_DEC_C2L_B), nc, d, W.mem, Sm));
            GCT_LAUNCH_CHECK();
        }
        for (int l = 0; l < N; ++l) {
            // K and V of the memory as two contiguous [B*Sm, d] slabs (rows [0,d) / [d,2d) of the fused k;v weight)
            const int ws = m.dec_slot(l, D_KV2_W), bs = m.dec_slot(l, D_KV2_B);
            Epilogue ek = Model<T>::epi(m.P(bs), d); ek.outT = W.kx + (size_t)l * B * Sm * d;
            GCT_TRY(m.gemm(W.mem, false, d, m.WT(ws), false, d, B * Sm, d, d, ek));
            Epilogue ev = Model<T>::epi(m.P(bs) + d, d); ev.outT = W.vx + (size_t)l * B * Sm * d;
            GCT_TRY(m.gemm(W.mem, false, d, m.WT(ws) + (size_t)d * d, false, d, B * Sm, d, d, ev));
        }
    }
    GCT_CUDA(cudaMemsetAsync(W.done, 0, B, st));
    GCT_CUDA(cudaMemsetAsync(D.status, 0, 2 * sizeof(int), st));
    GCT_CUDA(cudaMemsetAsync(W.key_valid, 0, (size_t)B * W.Lmax, st));
    return GCT_OK;
}

// one position: reads ys[:, pos], writes ys[:, pos+1] (sampled, or left untouched while inside the prefix)
template <typename T>
static int decode_one(Model<T>& m, const gct_decode_t& D, DecodeWs<T>& W, int pos, int step, bool sample) {
    // B = rows the step kernels run on (active-row decode: the compact rows of D.rowmap), Bp = rows of the caches / ys / uniforms
    const int d = m.d, dff = m.dff, N = m.N, Bp = W.B, B = D.n_active > 0 ? D.n_active : W.B, Sm = W.Sm, Lmax = W.Lmax;
    const int* rowmap = D.n_active > 0 ? D.rowmap : nullptr;
    const uint8_t* skip = (D.skip_done && sample) ? W.done : nullptr;
    cudaStream_t st = m.st;
    GCT_CUDA(launch_k(decode_embed_kernel, dim3(cdiv(B, 4)), dim3(128), 0, st, true, (const int64_t*)D.ys, D.max_len, pos, m.P(GCT_SLOT_DEC_EMB),
                      m.c.trg_vocab, m.P(GCT_SLOT_DEC_PE), 0, d, sqrtf((float)d), m.c.pad_id, W.x, W.key_valid, Lmax, B, rowmap));
    for (int l = 0; l < N; ++l) {
        GCT_TRY(m.norm_fwd(W.x, m.dec_slot(l, D_N1A), m.dec_slot(l, D_N1B), W.xn, nullptr, B));
        GCT_TRY(m.linear_T(W.xn, B, d, m.dec_slot(l, D_QKV_W), m.dec_slot(l, D_QKV_B), 3 * d, W.qkv));
        {
            DecAttnParams p;
            p.q = W.qkv; p.ldq = 3 * d; p.knew = W.qkv + d; p.vnew = W.qkv + 2 * d; p.ldnew = 3 * d;
            p.kcache = W.kc + (size_t)l * Bp * Lmax * d; p.vcache = W.vc + (size_t)l * Bp * Lmax * d;
            p.cache_bstride = (long long)Lmax * d; p.pitch = d; p.n_cached = pos; p.key_valid = W.key_valid; p.kv_stride = Lmax;
            p.out = W.att; p.ldo = d; p.H = m.H; p.scale = 0.125f; p.rowmap = rowmap; p.done = skip; p.rows_phys = Bp;
            GCT_TRY(launch_decode_attn<T>(p, B, st));
        }
        GCT_TRY(m.linear_res_norm(W.att, B, d, m.WT(m.dec_slot(l, D_O1_W)), m.P(m.dec_slot(l, D_O1_B)), W.x, W.x, m.site(0),
                                  m.dec_slot(l, D_N2A), m.dec_slot(l, D_N2B), W.xn, nullptr));
        if constexpr (sizeof(T) == 2) {
            if (W.zmode) {
                // cross-attention in latent space (decode_zattn.cuh): q and out projections carry the folded k / v / fc_z
                const int KZ = W.KZ;
                {
                    Epilogue e = Model<T>::epi(W.bqz + (size_t)l * KZ, KZ); e.outT = W.qz;
                    GCT_TRY(m.gemm(W.xn, false, d, W.wqz + (size_t)l * KZ * d, false, d, B, KZ, d, e));
                }
                ZAttnParams zp;
                zp.qz = W.qz; zp.ldq = KZ; zp.z = W.zlat; zp.z_bstride = (long long)W.Lz * m.lat; zp.key_valid = W.cross_mask + W.nck;
                zp.kv_stride = Sm; zp.n_keys = W.Lz; zp.kvc = W.nck ? W.kvc + (size_t)l * Bp * W.nck * 2 * d : nullptr; zp.nc = W.nck;
                zp.out = W.zbar; zp.ldo = KZ; zp.H = m.H; zp.B = B; zp.rowmap = rowmap; zp.done = skip;
                GCT_TRY(launch_decode_zattn(zp, m.lat, st));
                GCT_TRY(m.linear_res_norm(W.zbar, B, KZ, W.woz + (size_t)l * d * KZ, W.boz + (size_t)l * d, W.x, W.x, m.site(0),
                                          m.dec_slot(l, D_N3A), m.dec_slot(l, D_N3B), W.xn, nullptr));
            }
        }
        if (!W.zmode) {
            GCT_TRY(m.linear_T(W.xn, B, d, m.dec_slot(l, D_Q2_W), m.dec_slot(l, D_Q2_B), d, W.q2));
            {
                DecAttnParams p;
                p.q = W.q2; p.ldq = d; p.knew = nullptr; p.vnew = nullptr; p.ldnew = 0;
                p.kcache = W.kx + (size_t)l * Bp * Sm * d; p.vcache = W.vx + (size_t)l * Bp * Sm * d;
                p.cache_bstride = (long long)Sm * d; p.pitch = d; p.n_cached = Sm;
                p.key_valid = W.cross_mask; p.kv_stride = Sm; p.out = W.att; p.ldo = d; p.H = m.H; p.scale = 0.125f;
                p.rowmap = rowmap; p.done = skip;
                GCT_TRY(launch_decode_attn<T>(p, B, st));
            }
            GCT_TRY(m.linear_res_norm(W.att, B, d, m.WT(m.dec_slot(l, D_O2_W)), m.P(m.dec_slot(l, D_O2_B)), W.x, W.x, m.site(0),
                                      m.dec_slot(l, D_N3A), m.dec_slot(l, D_N3B), W.xn, nullptr));
        }
        {
            Epilogue e = Model<T>::epi(m.P(m.dec_slot(l, D_F1_B)), dff); e.flags = EPI_GELU; e.outT = W.hbuf;
            GCT_TRY(m.gemm(W.xn, false, d, m.WT(m.dec_slot(l, D_F1_W)), false, d, B, dff, d, e));
        }
        // (chosen by the call's PHYSICAL batch: an active-row decode that shrinks below 1025 rows keeps the residual form, so a row's
        // arithmetic does not depend on how many other rows are still running)
        if (Bp <= 1024) {  // x += hbuf W2^T + b2 : in-place accumulate lets the long-K GEMM use split-K (bias from split 0 only)
            Epilogue e = Model<T>::epi(m.P(m.dec_slot(l, D_F2_B)), d); e.out32 = W.x; e.flags = EPI_ACCUM;
            GCT_TRY(m.gemm(W.hbuf, false, dff, m.WT(m.dec_slot(l, D_F2_W)), false, dff, B, d, dff, e, 2));
        } else {           // enough row tiles to fill the machine: residual form, served by the specialised epilogue
            Epilogue e = Model<T>::epi(m.P(m.dec_slot(l, D_F2_B)), d); e.res32 = W.x; e.out32 = W.x;
            GCT_TRY(m.gemm(W.hbuf, false, dff, m.WT(m.dec_slot(l, D_F2_W)), false, dff, B, d, dff, e));
        }
    }
    if (!sample) return GCT_OK;      // prefix position: only the caches were needed
    GCT_TRY(m.norm_fwd(W.x, GCT_SLOT_DEC_NORM_A, GCT_SLOT_DEC_NORM_B, W.xn, nullptr, B));
    {
        Epilogue e = Model<T>::epi(m.P(GCT_SLOT_OUT_B), m.c.trg_vocab); e.out32 = W.logits;
        GCT_TRY(m.gemm(W.xn, false, d, m.WT(GCT_SLOT_OUT_W), false, d, B, m.c.trg_vocab, d, e));
    }
    SampleParams sp;
    sp.logits = W.logits; sp.ld = m.c.trg_vocab; sp.V = m.c.trg_vocab; sp.ys = D.ys; sp.ys_stride = D.max_len; sp.pos = pos;
    sp.forced = D.forced ? D.forced + pos + 1 : nullptr; sp.forced_stride = D.max_len;
    sp.uniforms = D.uniforms ? D.uniforms + (size_t)step * Bp : nullptr;
    sp.seed = D.seed; sp.step = step; sp.greedy = D.greedy; sp.eos_id = D.eos_id; sp.done = W.done; sp.n_done = D.status;
    sp.first_all_done = D.status + 1; sp.B = B; sp.rowmap = rowmap; sp.Bphys = Bp; sp.skip_done = D.skip_done; sp.pad_id = m.c.pad_id;
    const size_t step_off = (size_t)step * Bp * m.c.trg_vocab;
    sp.probs_out = D.probs_out ? D.probs_out + step_off : nullptr;
    sp.logits_out = D.logits_out ? D.logits_out + step_off : nullptr;
    GCT_CUDA(launch_k(decode_sample_kernel, dim3(cdiv(B, 4)), dim3(128), 0, st, true, sp));
    return GCT_OK;
}
