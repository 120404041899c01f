// Decode-time cross-attention evaluated in LATENT space (bf16 tier, memories without condition rows).
//
// The decoder memory is an affine image of the latent code, mem_j = Wz z_j + bz (Model/vaetf.py:93-95,
// Model/cvaetf.py:99-101), so for head h
//     score_hj = q_h . (Wk_h mem_j + bk_h) / 8 = [(Wk_h Wz)^T q_h] . z_j / 8 + const(h)     (the constant cancels in the softmax)
//     out_h    = sum_j p_hj (Wv_h mem_j + bv_h) = (Wv_h Wz) zbar_h + (Wv_h bz + bv_h),        zbar_h = sum_j p_hj z_j.
// Instead of streaming per-layer K/V projections of the memory ([keys][512] x 2 per layer and batch row) each step,
// the step reads the shared latent rows ([keys][LAT], 4x narrower, the same for every layer) and the two projections
// are folded into the query / output GEMMs once per decode (decode_zprep in model.cuh):
//     qz   = xn Wqz^T + bqz,   Wqz[h*LAT + a, :] = (1/8) sum_i (Wk Wz)[h*64+i, a] Wq[h*64+i, :]           [B, H*LAT]
//     zbar = this kernel                                                                                [B, H*LAT]
//     x   += zbar Woz^T + boz, Woz[:, h*LAT + a] = sum_i Wo[:, h*64+i] (Wv Wz)[h*64+i, a],  boz = Wo (Wv bz + bv) + bo.
// Algorithmic HBM bytes per batch row and layer: keys*LAT*2 + 2*H*LAT*2  (vs 2*keys*512*2 for the K/V form).
//
// One warp per batch row, tensor-core math through mma.sync.m16n8k16 (the problem per row is 16 x keys x LAT -- far too
// small for a tcgen05 tile): S = Qz Z^T with heads as the M dimension, softmax on the accumulator fragments, zbar = P Z
// with the probabilities re-used as the A fragment (FlashAttention-2 register layout).
#pragma once
#include "common.cuh"

constexpr int ZA_MAX_KEYS = 64;
constexpr int ZA_WARPS = 4;

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

struct ZAttnParams {
    const bf16* qz; int ldq;            // [B, ldq], head h at column h*LAT (already scaled by 1/sqrt(dk))
    const bf16* z; long long z_bstride; // [B][keys][LAT]
    const uint8_t* key_valid; int kv_stride;
    int n_keys;                         // keys per row (<= ZA_MAX_KEYS)
    bf16* out; int ldo;                 // [B, ldo]
    int H; int B;
};

template <int LAT>
__global__ void __launch_bounds__(ZA_WARPS * 32)
decode_zattn_kernel(ZAttnParams p) {
    constexpr int ROWB = LAT * 2 + 16;          // padded smem row: conflict-free fragment loads and ldmatrix rows
    constexpr int KS = LAT / 16;                // k-steps of the score GEMM
    constexpr int NTO = LAT / 8;                // n-tiles of the output GEMM
    extern __shared__ __align__(16) uint8_t za_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x * ZA_WARPS + warp;
    pdl_wait();
    pdl_launch_dependents();
    if (b >= p.B) return;
    uint8_t* zs = za_smem + (size_t)warp * ZA_MAX_KEYS * ROWB;
    const uint32_t zs_u = (uint32_t)__cvta_generic_to_shared(zs);
    // ---- attendable keys: bit j of (vlo, vhi); keys after the last attendable one are never fetched
    const uint8_t* valid = p.key_valid + (size_t)b * p.kv_stride;
    const uint32_t vlo = __ballot_sync(0xffffffffu, lane < p.n_keys && valid[lane] != 0);
    const uint32_t vhi = __ballot_sync(0xffffffffu, lane + 32 < p.n_keys && valid[lane + 32] != 0);
    int nrow = vhi ? 64 - __clz(vhi) : (vlo ? 32 - __clz(vlo) : 0);
    const bool none = nrow == 0;                // nothing attendable: uniform softmax over all keys (masked_fill -1e9 semantics)
    if (none) nrow = p.n_keys;
    const int nrow16 = (nrow + 15) & ~15;
    // ---- latent rows -> shared memory (16-byte async copies), zero rows up to the next multiple of 16
    const bf16* zg = p.z + (size_t)b * p.z_bstride;
    constexpr int CPR = LAT / 8;                // 16-byte chunks per row
    for (int i = lane; i < nrow16 * CPR; i += 32) {
        const int r = i / CPR, c = i % CPR;
        if (r < nrow) cp_async16(zs_u + r * ROWB + c * 16, zg + (size_t)r * LAT + c * 8);
        else *reinterpret_cast<uint4*>(zs + r * ROWB + c * 16) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // ---- query fragments (A operand: row = head): a0/a2 rows g, a1/a3 rows g + 8
    uint32_t qa[KS][4];
    {
        const bf16* q0 = p.qz + (size_t)b * p.ldq + (size_t)g * LAT + 2 * t;
        const bf16* q1 = q0 + (size_t)8 * LAT;
        const bool h0 = g < p.H, h1 = g + 8 < p.H;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            qa[ks][0] = h0 ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16) : 0u;
            qa[ks][2] = h0 ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8) : 0u;
            qa[ks][1] = h1 ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16) : 0u;
            qa[ks][3] = h1 ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8) : 0u;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- S = Qz Z^T : n-tile nt covers keys nt*8 .. nt*8+7
    constexpr int NT = ZA_MAX_KEYS / 8;
    float s[NT][4];
    const int ntiles = nrow16 >> 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        if (nt < ntiles) {
            const uint8_t* zr = zs + (nt * 8 + g) * ROWB + 4 * t;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(zr + ks * 32);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(zr + ks * 32 + 16);
                mma_bf16_16816(s[nt], qa[ks], b0, b1);
            }
        }
    }
    // ---- masked softmax over keys; this thread holds keys nt*8 + 2t, +1 of head rows g (c0,c1) and g+8 (c2,c3)
    constexpr float kL2e = 1.4426950408889634f;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        if (nt < ntiles) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = nt * 8 + 2 * t + e;
                const bool ok = (j < 32) ? ((vlo >> j) & 1u) : ((vhi >> (j - 32)) & 1u);
                float v0 = s[nt][e] * kL2e, v1 = s[nt][2 + e] * kL2e;
                if (!ok) { v0 = -1e9f * kL2e; v1 = -1e9f * kL2e; }
                if (j >= nrow) { v0 = -INFINITY; v1 = -INFINITY; }
                s[nt][e] = v0; s[nt][2 + e] = v1;
                mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
            }
        }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[NT][2];                          // bf16x2 probabilities: [nt][0] row g, [nt][1] row g+8
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        pa[nt][0] = pa[nt][1] = 0u;
        if (nt < ntiles) {
            const float p00 = ex2_approx(s[nt][0] - mx0), p01 = ex2_approx(s[nt][1] - mx0);
            const float p10 = ex2_approx(s[nt][2] - mx1), p11 = ex2_approx(s[nt][3] - mx1);
            l0 += p00 + p01; l1 += p10 + p11;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(p00, p01), h1 = __floats2bfloat162_rn(p10, p11);
            pa[nt][0] = *reinterpret_cast<uint32_t*>(&h0);
            pa[nt][1] = *reinterpret_cast<uint32_t*>(&h1);
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- zbar = P Z : k-step kk covers keys kk*16 .. +15 (A = probability fragments), n-tile = 8 latent dims
    float o[NTO][4];
#pragma unroll
    for (int n = 0; n < NTO; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    const int ksteps = nrow16 >> 4;
    // ldmatrix.x4.trans: lanes 0-7 / 8-15 address keys kk*16+0..7 / +8..15 of dims n0..n0+7, lanes 16-31 the same keys of dims n0+8..
    const uint32_t lm_row = zs_u + (uint32_t)((lane & 15) * ROWB + (lane >> 4) * 16);
#pragma unroll
    for (int kk = 0; kk < ZA_MAX_KEYS / 16; ++kk) {
        if (kk < ksteps) {
            const uint32_t a[4] = {pa[2 * kk][0], pa[2 * kk][1], pa[2 * kk + 1][0], pa[2 * kk + 1][1]};
#pragma unroll
            for (int n2 = 0; n2 < NTO / 2; ++n2) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(lm_row + (uint32_t)(kk * 16 * ROWB + n2 * 32), b0, b1, b2, b3);
                mma_bf16_16816(o[2 * n2], a, b0, b1);
                mma_bf16_16816(o[2 * n2 + 1], a, b2, b3);
            }
        }
    }
    // ---- normalise and store: rows g / g+8 = heads, columns n*8 + 2t, +1
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    bf16* o0 = p.out + (size_t)b * p.ldo + (size_t)g * LAT + 2 * t;
    bf16* o1 = o0 + (size_t)8 * LAT;
#pragma unroll
    for (int n = 0; n < NTO; ++n) {
        if (g < p.H) *reinterpret_cast<__nv_bfloat162*>(o0 + n * 8) = __floats2bfloat162_rn(o[n][0] * i0, o[n][1] * i0);
        if (g + 8 < p.H) *reinterpret_cast<__nv_bfloat162*>(o1 + n * 8) = __floats2bfloat162_rn(o[n][2] * i1, o[n][3] * i1);
    }
}

template <int LAT>
static int launch_decode_zattn_lat(const ZAttnParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)ZA_WARPS * ZA_MAX_KEYS * (LAT * 2 + 16);
    GCT_SMEM_LIMIT(decode_zattn_kernel<LAT>, smem);
    GCT_CUDA(launch_k(decode_zattn_kernel<LAT>, dim3((p.B + ZA_WARPS - 1) / ZA_WARPS), dim3(ZA_WARPS * 32), smem, st, true, p));
    return GCT_OK;
}
static bool zattn_supported(int lat, int H, int n_keys) { return (lat == 128 || lat == 64) && H >= 1 && H <= 16 && n_keys >= 1 && n_keys <= ZA_MAX_KEYS; }
static int launch_decode_zattn(const ZAttnParams& p, int lat, cudaStream_t st) {
    GCT_REQUIRE(zattn_supported(lat, p.H, p.n_keys), "latent-space cross-attention: lat=%d H=%d keys=%d unsupported", lat, p.H, p.n_keys);
    return lat == 128 ? launch_decode_zattn_lat<128>(p, st) : launch_decode_zattn_lat<64>(p, st);
}
