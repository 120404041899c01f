// Decode-time cross-attention evaluated in LATENT space (bf16 tier; every sampler: vaetf, pvaetf, scavaetf, pscavaetf).
//
// The decoder memory is an affine image of the latent code, mem_j = Wz z_j + bz (Model/vaetf.py:93-95,
// Model/cvaetf.py:99-101), so for head h
//     score_hj = q_h . (Wk_h mem_j + bk_h) / 8 = [(Wk_h Wz)^T q_h] . z_j / 8 + C_h,   C_h = q_h . (Wk_h bz + bk_h) / 8
//     value_j  = Wv_h mem_j + bv_h             = (Wv_h Wz) z_j + (Wv_h bz + bv_h).
// With use_cond2lat the memory carries `nc` extra rows c_i = embed_cond2lat(dconds)_i in front (Model/cvaetf.py:107-116); they
// are not images of a latent, but relative to the same constants they are ordinary keys / values of u_i = c_i - bz:
//     score_hi = q_h . (Wk_h u_i) / 8 + C_h,      value_i = Wv_h u_i + (Wv_h bz + bv_h).
// C_h cancels in the softmax and the value constant factors out (the probabilities sum to one), so per layer
//     qz   = xn Wqz^T + bqz          Wqz[h*LAT + a, :] = (1/8) sum_i (Wk Wz)[h*64+i, a] Wq[h*64+i, :]        [B, H*LAT]
//     q8   = (xn Wq^T + bq) / 8      (only with condition rows)                                          [B, d]
//     zbar_h = sum_j p_hj z_j ,  ac_h = sum_i p_hi (Wv u_i)_h        <- this kernel, one softmax over latent + condition keys
//     x   += zbar Woz^T + ac Wo^T + boz,   Woz[:, (h, a)] = sum_i Wo[:, h*64+i] (Wv Wz)[h*64+i, a],  boz = Wo (Wv bz + bv) + bo.
// The two projections of every layer are folded into the query / output GEMMs once per decode (decode_zprep in model.cuh);
// the per-layer K/V of the condition rows (nc x 2d per batch row) are computed once per decode as well.
// Algorithmic HBM bytes per batch row and layer: keys*LAT*2 + 2*H*LAT*2 (+ nc*2d*2 + 2d*2)   vs   2*(nc+keys)*d*2 in K/V form.
//
// Kernel: one warp per batch row, no shared memory.  Keys are consumed 16 at a time with an online softmax; both GEMMs are
// mma.sync.m16n8k16 with the 8 heads as the N dimension (a tcgen05 tile would be > 90 % padding at 16 x 8 x 128 per chunk):
//     S^T[key][head]    = Z[key][dim] . Q^T[dim][head]          A = latent rows straight from global memory (16-byte loads),
//     zbar^T[dim][head] = Z^T[dim][key] . P^T[key][head]        A = the SAME registers transposed in place by movmatrix.
// A lane (g = lane/4, t = lane%4) loads z[key g (and g+8)][32*blk + 8t .. +8): the contraction index of the first GEMM is a
// free permutation, so those four registers serve as two k-steps as they are (the query fragments are loaded with the same
// permutation); movmatrix.trans of each register yields the A fragment of the second GEMM with the output dims permuted --
// which lands every lane on 4-byte pieces of contiguous 32-byte runs of zbar in (dim, head) order.
#pragma once
#include "common.cuh"

constexpr int ZA_WARPS = 4;

__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 8x8 b16 tile held one register per lane (lane (g,t): row g, columns 2t, 2t+1) -> its transpose in the same layout
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

struct ZAttnParams {
    const bf16* qz; int ldq;            // [B, ldq]: columns [0, H*LAT) latent-space queries, head h at h*LAT (already scaled by 1/8);
                                        //           with condition rows, columns [H*LAT, H*LAT + 64H) = q / 8
    const bf16* z; long long z_bstride; // [B][n_keys][LAT] latent rows
    const uint8_t* key_valid; int kv_stride;   // validity of latent key j of row b at key_valid[b*kv_stride + j]
    int n_keys;
    const bf16* kvc; int nc;            // condition rows: [B][nc][2*64H] = (Wk u_i | Wv u_i), or nc = 0
    bf16* out; int ldo;                 // [B, ldo]: columns [0, LAT*H) zbar in (dim, head) order: a*H + h; then 64H columns ac
    int H; int B;
    // active-row decode: qz / out rows are COMPACT, z / key_valid / kvc belong to the physical row rowmap[b]; finished rows return
    const int* rowmap = nullptr;
    const uint8_t* done = nullptr;
};

// SM = true (round 2, late): the row's latent block is staged in shared memory by bulk async copies issued BEFORE the programmatic
// dependency wait -- the latent rows, masks and condition K/V are constants of the decode call, only the queries come from the
// preceding GEMM -- so for the CTAs of the first wave the bytes arrive while that GEMM drains, and every later warp has its whole
// row (not two 16-key chunks) in flight after one round trip.  The warp's buffer is private (no CTA-level synchronisation); the
// first 32 keys are requested at once, the rest as soon as the mask has told where the row ends.  The MMA fragments are then read
// from shared memory with the same 16-byte-per-lane pattern the global loads used (2-way bank conflicts: irrelevant at 4 KB per
// chunk), which frees the 64 registers of the second in-flight chunk: 4 CTAs per SM.
template <int LAT, int MINB, bool SM, int NW = ZA_WARPS>
__global__ void __launch_bounds__(NW * 32, MINB)
decode_zattn_kernel(ZAttnParams p) {
    constexpr int NB = LAT / 32;                // 32-dim blocks of a latent row (one 16-byte load per lane and block)
    constexpr float kL2e = 1.4426950408889634f;
    extern __shared__ __align__(128) uint8_t za_smem[];
    __shared__ __align__(8) uint64_t za_bars[NW][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x * NW + warp;
    if constexpr (!SM) {
        pdl_wait();
        pdl_launch_dependents();
    }
    const bool live = b < p.B;
    const int bp = live ? (p.rowmap ? p.rowmap[b] : b) : 0;
    const bool run = live && !(p.done != nullptr && p.done[bp] != 0);
    if constexpr (!SM) { if (!run) return; }
    const int H = p.H, d = 64 * H, nkeys = p.n_keys;
    const bf16* zg = p.z + (size_t)bp * p.z_bstride + t * 8;
    const uint8_t* valid = p.key_valid + (size_t)bp * p.kv_stride;
    const bf16* qrow = p.qz + (size_t)b * p.ldq;
    // ---- SM: request the row before waiting for the producer of the queries
    const uint32_t zs = tc::smem_u32(za_smem) + (uint32_t)warp * (uint32_t)nkeys * (LAT * 2);
    const uint32_t bar_a = tc::smem_u32(&za_bars[warp][0]), bar_b = bar_a + 8;
    int nrow_sm = 0;
    if constexpr (SM) {
        if (run) {
            if (lane == 0) {
                tc::mbar_init(bar_a, 1);
                tc::mbar_init(bar_b, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                const uint32_t n0 = (uint32_t)min(nkeys, 32) * (LAT * 2);
                tc::mbar_expect_tx(bar_a, n0);
                bulk_g2s(zs, p.z + (size_t)bp * p.z_bstride, n0, bar_a);
            }
            for (int j0 = 0; j0 < nkeys; j0 += 32) {
                const uint32_t bal = __ballot_sync(0xffffffffu, j0 + lane < nkeys && valid[j0 + lane] != 0);
                if (bal) nrow_sm = j0 + 32 - __clz(bal);
            }
            if (nrow_sm == 0 && p.nc == 0) nrow_sm = nkeys;
            if (nrow_sm > 32 && lane == 0) {
                const uint32_t n1 = (uint32_t)(nrow_sm - 32) * (LAT * 2);
                tc::mbar_expect_tx(bar_b, n1);
                bulk_g2s(zs + 32 * (LAT * 2), p.z + (size_t)bp * p.z_bstride + (size_t)32 * LAT, n1, bar_b);
            }
        }
        pdl_wait();
        pdl_launch_dependents();
        if (!run) return;
    }

    // The first TWO 16-key chunks are requested before anything else (their addresses depend on nothing but n_keys): with the
    // query fragments and the validity bytes that is one round trip to memory for every row of up to 32 keys, two for the rest
    // (ncu, first version with one chunk in flight: 47 % of the stall samples were long-scoreboard waits, 22 % occupancy).
    uint4 za[NB], zb[NB], na[SM ? 1 : NB], nb[SM ? 1 : NB];
    auto load_chunk = [&](int chunk, uint4* da, uint4* db, int limit) {
        const int ka = chunk * 16 + g, kb = ka + 8;
        const bool ia = ka < limit, ib = kb < limit;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            if constexpr (SM) {
                da[k] = ia ? lds16(zs + (uint32_t)ka * (LAT * 2) + k * 64 + t * 16) : make_uint4(0, 0, 0, 0);
                db[k] = ib ? lds16(zs + (uint32_t)kb * (LAT * 2) + k * 64 + t * 16) : make_uint4(0, 0, 0, 0);
            } else {
                da[k] = ia ? ldg_nc16(zg + (size_t)ka * LAT + k * 32) : make_uint4(0, 0, 0, 0);
                db[k] = ib ? ldg_nc16(zg + (size_t)kb * LAT + k * 32) : make_uint4(0, 0, 0, 0);
            }
        }
    };
    if constexpr (!SM) {
        load_chunk(0, za, zb, nkeys);
        load_chunk(1, na, nb, nkeys);
    }
    // query fragments: lane (g,t) = head g, dims 32*blk + 8t .. +8 (same permutation as the latent rows)
    uint4 q[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) q[k] = (g < H) ? ldg_nc16(qrow + (size_t)g * LAT + k * 32 + t * 8) : make_uint4(0, 0, 0, 0);
    // last attendable latent key: later ones are never fetched
    int nrow = nrow_sm;
    if constexpr (!SM) {
        for (int j0 = 0; j0 < nkeys; j0 += 32) {
            const uint32_t bal = __ballot_sync(0xffffffffu, j0 + lane < nkeys && valid[j0 + lane] != 0);
            if (bal) nrow = j0 + 32 - __clz(bal);
        }
        if (nrow == 0 && p.nc == 0) nrow = nkeys;   // nothing attendable: every score is -1e9 -> uniform softmax over all keys (masked_fill semantics)
    }

    // running softmax state of heads 2t (index 0) and 2t+1 (index 1); l is a per-lane partial, reduced over g at the end
    float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
    // ---- condition rows: plain dot products; lane owns 16-byte piece(s) pc = lane + 32 j of the 64H-wide row, head pc / 8
    constexpr int MAXC = 8;
    float sc[2][MAXC];                          // log2-scaled score of condition row i for the head of piece j, replicated over its 8 lanes
    const int npieces = 8 * H;                  // 16-byte pieces per 64H row
    if (p.nc > 0) {
        const bf16* kv = p.kvc + (size_t)bp * p.nc * 2 * d;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pc = lane + 32 * j;
            f8 qf;
#pragma unroll
            for (int e = 0; e < 8; ++e) qf.v[e] = 0.f;
            if (pc < npieces) qf = ld8(qrow + (size_t)H * LAT + pc * 8);
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                sc[j][i] = -INFINITY;
                if (i < p.nc) {
                    float dsum = 0.f;
                    if (pc < npieces) {
                        const f8 kf = ld8(kv + (size_t)i * 2 * d + pc * 8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) dsum = fmaf(qf.v[e], kf.v[e], dsum);
                    }
                    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                    dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                    dsum += __shfl_xor_sync(0xffffffffu, dsum, 4);
                    sc[j][i] = dsum * kL2e;      // condition rows are always attendable (src_mask is extended by ones, cvaetf.py:114-116)
                }
            }
        }
        // fold them into the running state of this lane's two heads
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int h = 2 * t + e;
            float mx = -INFINITY, vals[MAXC];
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                const float v0 = __shfl_sync(0xffffffffu, sc[0][i], (h & 3) * 8), v1 = __shfl_sync(0xffffffffu, sc[1][i], (h & 3) * 8);
                vals[i] = (i < p.nc && h < H) ? ((h >> 2) ? v1 : v0) : -INFINITY;
                mx = fmaxf(mx, vals[i]);
            }
            if (h < H) {
                mrun[e] = mx;
                if (g == 0) {                   // counted once (l is summed over the eight g lanes later)
#pragma unroll
                    for (int i = 0; i < MAXC; ++i) if (i < p.nc) lrun[e] += ex2_approx(vals[i] - mx);
                }
            }
        }
    }

    float acc[2 * NB][4];
#pragma unroll
    for (int i = 0; i < 2 * NB; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    // one 16-key chunk: scores, online softmax update, context accumulation
    auto consume = [&](const uint4* ca, const uint4* cb, int c0) {
        const int ka = c0 + g, kb = ka + 8;
        const bool ina = ka < nrow, inb = kb < nrow;
        const bool oka = ina && valid[ka] != 0, okb = inb && valid[kb] != 0;
        // ---- S^T = Z Q^T (16 keys x 8 heads), k permuted: two k-steps per 32-dim block, two accumulators to halve the chain
        float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            mma_bf16_16816(s, ca[k].x, cb[k].x, ca[k].y, cb[k].y, q[k].x, q[k].y);
            mma_bf16_16816(s2, ca[k].z, cb[k].z, ca[k].w, cb[k].w, q[k].z, q[k].w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) s[e] += s2[e];
        // s[0], s[1]: key ka, heads 2t, 2t+1 ; s[2], s[3]: key kb
        float pa[2], pb[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float va = ina ? (oka ? s[e] * kL2e : -1e9f * kL2e) : -INFINITY;
            const float vb = inb ? (okb ? s[2 + e] * kL2e : -1e9f * kL2e) : -INFINITY;
            float mx = fmaxf(va, vb);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
            const float mn = fmaxf(mrun[e], mx);            // finite: the chunk holds at least one in-range key
            const float corr = ex2_approx(mrun[e] - mn);    // 0 on the first chunk (mrun = -inf)
            pa[e] = ex2_approx(va - mn);
            pb[e] = ex2_approx(vb - mn);
            lrun[e] = lrun[e] * corr + pa[e] + pb[e];
            mrun[e] = mn;
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) { acc[i][e] *= corr; acc[i][2 + e] *= corr; }
        }
        // ---- zbar^T += Z^T P^T : B = transposed probability tiles, A = transposed latent tiles (two 8-dim groups per m-tile)
        const uint32_t pb0 = movmatrix_trans(pack_bf16x2(pa[0], pa[1]));
        const uint32_t pb1 = movmatrix_trans(pack_bf16x2(pb[0], pb[1]));
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            mma_bf16_16816(acc[2 * k], movmatrix_trans(ca[k].x), movmatrix_trans(ca[k].y), movmatrix_trans(cb[k].x), movmatrix_trans(cb[k].y), pb0, pb1);
            mma_bf16_16816(acc[2 * k + 1], movmatrix_trans(ca[k].z), movmatrix_trans(ca[k].w), movmatrix_trans(cb[k].z), movmatrix_trans(cb[k].w), pb0, pb1);
        }
    };
    if constexpr (SM) {
        for (int c0 = 0; c0 < nrow; c0 += 16) {
            if (c0 == 0) tc::mbar_wait(bar_a, 0);
            if (c0 == 32) tc::mbar_wait(bar_b, 0);
            load_chunk(c0 >> 4, za, zb, nrow);
            consume(za, zb, c0);
        }
    }
    // two register buffers, each refilled with the chunk two ahead as soon as it has been consumed
    for (int c0 = 0; !SM && c0 < nrow; c0 += 32) {
        consume(za, zb, c0);
        if (c0 + 32 < nrow) load_chunk((c0 >> 4) + 2, za, zb, nrow);
        if (c0 + 16 < nrow) {
            consume(na, nb, c0 + 16);
            if (c0 + 48 < nrow) load_chunk((c0 >> 4) + 3, na, nb, nrow);
        }
    }
    // ---- normalise: l over the eight key slots (g); every lane ends with the totals of its two heads
    float inv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        float l = lrun[e];
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        inv[e] = 1.f / l;
    }
    // acc[2k + m][0..1]: dim 32k + 8(g/2) + 4m + (g&1) (+2 for [2..3]), heads 2t, 2t+1  ->  out[dim*H + head]
    bf16* orow = p.out + (size_t)b * p.ldo;
    const int dbase = 8 * (g >> 1) + (g & 1);
#pragma unroll
    for (int i = 0; i < 2 * NB; ++i) {
        const int dim0 = 32 * (i >> 1) + 4 * (i & 1) + dbase;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int dim = dim0 + 2 * r;
            const float v0 = acc[i][2 * r] * inv[0], v1 = acc[i][2 * r + 1] * inv[1];
            bf16* o = orow + (size_t)dim * H + 2 * t;
            if (2 * t + 1 < H && (H & 1) == 0) *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v0, v1);
            else {
                if (2 * t < H) o[0] = __float2bfloat16(v0);
                if (2 * t + 1 < H) o[1] = __float2bfloat16(v1);
            }
        }
    }
    // ---- condition values: ac[head of the piece][8 dims] = sum_i p_i (Wv u_i)
    if (p.nc > 0) {
        const bf16* kv = p.kvc + (size_t)bp * p.nc * 2 * d + d;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pc = lane + 32 * j;
            const int h = (pc >> 3) < H ? (pc >> 3) : 0;
            // final (max, 1/l) of head h live on the lanes with t = h / 2 (any g), component h & 1
            const float m0 = __shfl_sync(0xffffffffu, mrun[0], h >> 1), m1 = __shfl_sync(0xffffffffu, mrun[1], h >> 1);
            const float i0 = __shfl_sync(0xffffffffu, inv[0], h >> 1), i1 = __shfl_sync(0xffffffffu, inv[1], h >> 1);
            const float mh = (h & 1) ? m1 : m0, ih = (h & 1) ? i1 : i0;
            if (pc < npieces) {
                f8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] = 0.f;
#pragma unroll
                for (int i = 0; i < MAXC; ++i) {
                    if (i < p.nc) {
                        const float pi = ex2_approx(sc[j][i] - mh) * ih;
                        const f8 vf = ld8(kv + (size_t)i * 2 * d + pc * 8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o.v[e] = fmaf(pi, vf.v[e], o.v[e]);
                    }
                }
                st8(orow + (size_t)LAT * H + pc * 8, o);
            }
        }
    }
}

extern int g_za_cfg;      // tuning knob: 0 = rows staged in shared memory (default); 3 / 4 = rows straight from global memory into registers,
                          // compiled for that many resident CTAs per SM (3: 160 registers, no spills; 4: 128 registers)

template <int LAT>
static int launch_decode_zattn_lat(const ZAttnParams& p, cudaStream_t st) {
    const dim3 grid((p.B + ZA_WARPS - 1) / ZA_WARPS), block(ZA_WARPS * 32);
    const size_t wbytes = (size_t)p.n_keys * LAT * 2;          // one row's latent block = one warp's staging buffer
    const size_t smem = (size_t)ZA_WARPS * wbytes;
    if (g_za_cfg == 4) GCT_CUDA(launch_k(decode_zattn_kernel<LAT, 4, false>, grid, block, 0, st, true, p));
    else if (g_za_cfg == 3 || smem > 100 * 1024) GCT_CUDA(launch_k(decode_zattn_kernel<LAT, 3, false>, grid, block, 0, st, true, p));
    else if (g_za_cfg != 5 && 5 * (3 * wbytes + 1280) <= 228 * 1024 && 4 * (4 * wbytes + 1280) > 228 * 1024) {
        // rows staged in shared memory, 3 warps per CTA: five CTAs = 15 warps per SM where four CTAs of 4 warps do not fit
        // (53 .. 59 keys of 128 latent dims; the kernel is bound by how many rows an SM has in flight)
        auto kern = decode_zattn_kernel<LAT, 5, true, 3>;
        GCT_SMEM_LIMIT(kern, 3 * wbytes);
        GCT_CUDA(launch_k(kern, dim3((p.B + 2) / 3), dim3(96), 3 * wbytes, st, true, p));
    } else {
        // rows staged in shared memory (default): 4 warps x n_keys x LAT x 2 bytes per CTA, three or four CTAs per SM
        auto kern = decode_zattn_kernel<LAT, 4, true>;
        GCT_SMEM_LIMIT(kern, smem);
        GCT_CUDA(launch_k(kern, grid, block, smem, st, true, p));
    }
    return GCT_OK;
}
static bool zattn_supported(int lat, int H, int nc) { return (lat == 128 || lat == 64 || lat == 32) && H >= 1 && H <= 8 && nc >= 0 && nc <= 8; }
static int launch_decode_zattn(const ZAttnParams& p, int lat, cudaStream_t st) {
    GCT_REQUIRE(zattn_supported(lat, p.H, p.nc) && p.n_keys >= 1, "latent-space cross-attention: lat=%d H=%d nc=%d keys=%d unsupported", lat, p.H, p.nc, p.n_keys);
    return lat == 128 ? launch_decode_zattn_lat<128>(p, st) : lat == 64 ? launch_decode_zattn_lat<64>(p, st) : launch_decode_zattn_lat<32>(p, st);
}
