// KV-cached autoregressive decode kernels (replace the O(T^2) loop of
// Inference/sampling_tool.py:140-184): single-query attention over a cache, and the on-device
// next-token sampler (softmax -> argmax / inverse-CDF multinomial -> append -> <eos> bookkeeping).
// HBM-bound: each (batch, head) streams its K/V rows once with 16-byte loads.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

struct DecAttnParams {
    const void* q; int ldq;                  // [B, ldq], head h at column h*64
    const void* knew; const void* vnew; int ldnew;   // this step's K/V rows (self-attention) or null
    void* kcache; void* vcache;              // [B, Lmax, pitch] ; pitch in elements
    long long cache_bstride; int pitch;
    int n_cached;                            // keys already in the cache (self: pos ; cross: Sm)
    const uint8_t* key_valid; int kv_stride; // [B, kv_stride] 1 = attend (0 -> score -1e9)
    void* out; int ldo;                      // [B, ldo]
    int H; float scale;
    // active-row decode (gct_decode_t.rowmap / skip_done): q / knew / vnew / out are indexed by the COMPACT row i, the caches and
    // key_valid by the physical row rowmap[i]; a row whose `done` flag is set is not touched at all (no cache append, out stale)
    const int* rowmap = nullptr;
    const uint8_t* done = nullptr;
    long long rows_phys = 0;                 // rows of the caches (0 = B): the extent of the tensor maps' row dimension
};

// One CTA per batch row: a producer warp streams that row's K and V cache slabs (contiguous [keys][H*64]) through a
// DA_NS-stage shared-memory ring with 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx), so the bytes in
// flight are set by the ring depth, not by registers or warp count; H consumer warps (one per head) run an online
// softmax over the chunks.  lane = (g = lane/8 : key slot inside the chunk, sub = lane%8 : 8-dim slice of the head).
// Keys after the row's last attendable key are never fetched (cross-attention latents are padded per batch).
constexpr int DEC_MAX_KEYS = 256;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// DA_CHUNK keys per ring stage (each stage holds a K chunk and a V chunk), DA_NS stages, DA_ROWS batch rows per CTA
// (the copy ring keeps streaming across them).
template <typename T, int DA_CHUNK, int DA_NS, int DA_ROWS>
__global__ void __launch_bounds__(288)
decode_attn_kernel(DecAttnParams p, int B) {
    extern __shared__ __align__(128) uint8_t da_smem[];
    __shared__ uint8_t valid_s[DA_ROWS][DEC_MAX_KEYS];
    __shared__ __align__(8) uint64_t bars[2 * DA_NS];
    __shared__ int nrow_s[DA_ROWS];
    const int b0 = blockIdx.x * DA_ROWS;
    const int nrows_cta = min(DA_ROWS, B - b0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = p.H, rowlen = H * 64;
    const uint32_t rowbytes = rowlen * sizeof(T);
    const uint32_t chunk_bytes = DA_CHUNK * rowbytes;
    const uint32_t ring = tc::smem_u32(da_smem);
    const uint32_t bar0 = tc::smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int s = 0; s < DA_NS; ++s) { tc::mbar_init(bar0 + 8 * s, 1); tc::mbar_init(bar0 + 8 * (DA_NS + s), H); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < DA_ROWS) nrow_s[threadIdx.x] = 0;
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();
    const int nc = p.n_cached;
    const int nall = nc + (p.knew ? 1 : 0);
    auto phys = [&](int r) { return p.rowmap ? p.rowmap[b0 + r] : b0 + r; };
    auto skipped = [&](int r) { return p.done != nullptr && p.done[phys(r)] != 0; };      // uniform over the CTA
    if (DA_ROWS == 1 && skipped(0)) return;
    for (int r = 0; r < nrows_cta; ++r) {
        if (DA_ROWS > 1 && skipped(r)) continue;
        const uint8_t* valid = p.key_valid + (size_t)phys(r) * p.kv_stride;
        for (int j = threadIdx.x; j < nall; j += blockDim.x) {
            const uint8_t v = valid[j];
            valid_s[r][j] = v;
            if (v && j < nc) atomicMax(&nrow_s[r], j + 1);
        }
    }
    __syncthreads();
    // cached keys that must be fetched per row; if none is attendable every score is -1e9 and the softmax is uniform
    // over all of them (the reference's masked_fill behaviour), so fetch them all in that corner case
    auto row_keys = [&](int r) { return (nrow_s[r] == 0 && !(p.knew && valid_s[r][nc])) ? nc : nrow_s[r]; };

    if (warp == H) {
        // ---------------- producer: one continuous stream of chunks over the CTA's rows ----------------
        if (lane == 0) {
            int g = 0;
            for (int r = 0; r < nrows_cta; ++r) {
                if (DA_ROWS > 1 && skipped(r)) continue;
                const int nrow = row_keys(r);
                const T* kslab = reinterpret_cast<const T*>(p.kcache) + (size_t)phys(r) * p.cache_bstride;
                const T* vslab = reinterpret_cast<const T*>(p.vcache) + (size_t)phys(r) * p.cache_bstride;
                for (int c = 0; c * DA_CHUNK < nrow; ++c, ++g) {
                    const int s = g % DA_NS;
                    tc::mbar_wait(bar0 + 8 * (DA_NS + s), ((g / DA_NS) & 1) ^ 1);
                    const int nk = min(DA_CHUNK, nrow - c * DA_CHUNK);
                    const uint32_t bytes = nk * rowbytes;
                    const uint32_t dst = ring + s * 2 * chunk_bytes;
                    tc::mbar_expect_tx(bar0 + 8 * s, 2 * bytes);
                    bulk_g2s(dst, kslab + (size_t)c * DA_CHUNK * p.pitch, bytes, bar0 + 8 * s);
                    bulk_g2s(dst + chunk_bytes, vslab + (size_t)c * DA_CHUNK * p.pitch, bytes, bar0 + 8 * s);
                }
            }
        }
        return;
    }
    // ---------------- consumers: warp = head ----------------
    const int h = warp, gq = lane >> 3, sub = lane & 7;
    const int col = h * 64 + sub * 8;
    const unsigned gmask = 0xFFu << (gq * 8);
    int g = 0;
    for (int r = 0; r < nrows_cta; ++r) {
        if (DA_ROWS > 1 && skipped(r)) continue;
        const int b = b0 + r;
        const int nrow = row_keys(r);
        const int nchunks = (nrow + DA_CHUNK - 1) / DA_CHUNK;
        const uint8_t* vld = valid_s[r];
        T* kslab = reinterpret_cast<T*>(p.kcache) + (size_t)phys(r) * p.cache_bstride;
        T* vslab = reinterpret_cast<T*>(p.vcache) + (size_t)phys(r) * p.cache_bstride;
        const f8 q = ld8(reinterpret_cast<const T*>(p.q) + (size_t)b * p.ldq + col);
        f8 kn, vn;
        if (p.knew) {
            kn = ld8(reinterpret_cast<const T*>(p.knew) + (size_t)b * p.ldnew + col);
            vn = ld8(reinterpret_cast<const T*>(p.vnew) + (size_t)b * p.ldnew + col);
            if (gq == 0) {       // append this step's K/V row to the cache
                st8(kslab + (size_t)nc * p.pitch + col, kn);
                st8(vslab + (size_t)nc * p.pitch + col, vn);
            }
        }
        float m = -INFINITY, l = 0.f;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % DA_NS;
            tc::mbar_wait(bar0 + 8 * s, (g / DA_NS) & 1);
            const T* Ks = reinterpret_cast<const T*>(da_smem + (size_t)s * 2 * chunk_bytes);
            const T* Vs = reinterpret_cast<const T*>(da_smem + (size_t)s * 2 * chunk_bytes + chunk_bytes);
            const int nk = min(DA_CHUNK, nrow - c * DA_CHUNK);
            constexpr int KPG = DA_CHUNK / 4;        // keys per 8-lane group per chunk
            float sc[KPG];
            float cm = -INFINITY;
#pragma unroll
            for (int i = 0; i < KPG; ++i) {
                const int jj = gq + 4 * i;
                sc[i] = -INFINITY;
                if (jj < nk) {      // uniform within the 8-lane group
                    const f8 kf = ld8(Ks + (size_t)jj * rowlen + col);
                    float d = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) d = fmaf(q.v[e], kf.v[e], d);
                    d += __shfl_xor_sync(gmask, d, 1);
                    d += __shfl_xor_sync(gmask, d, 2);
                    d += __shfl_xor_sync(gmask, d, 4);
                    sc[i] = vld[c * DA_CHUNK + jj] ? d * p.scale : -1e9f;
                }
                cm = fmaxf(cm, sc[i]);
            }
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 8));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 16));
            const float mn = fmaxf(m, cm);
            const float corr = __expf(m - mn);
            l *= corr;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
            for (int i = 0; i < KPG; ++i) {
                const int jj = gq + 4 * i;
                if (jj < nk) {
                    const float pj = __expf(sc[i] - mn);
                    const f8 vf = ld8(Vs + (size_t)jj * rowlen + col);
                    l += pj;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf.v[e], acc[e]);
                }
            }
            m = mn;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (DA_NS + s));
        }
        if (p.knew) {
            // this step's own key: every lane computes the same score (cheap), slot 0 adds the value
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(q.v[e], kn.v[e], d);
            d += __shfl_xor_sync(gmask, d, 1);
            d += __shfl_xor_sync(gmask, d, 2);
            d += __shfl_xor_sync(gmask, d, 4);
            const float sn = vld[nc] ? d * p.scale : -1e9f;
            const float mn = fmaxf(m, sn);
            const float corr = __expf(m - mn);
            l *= corr;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] *= corr;
            if (gq == 0) {
                const float pj = __expf(sn - mn);
                l += pj;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vn.v[e], acc[e]);
            }
            m = mn;
        }
        // l counted each key once per lane of its group: reduce over the four key slots
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        const float inv = 1.f / l;
        f8 o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float a = acc[e];
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            o.v[e] = a * inv;
        }
        if (gq == 0) st8(reinterpret_cast<T*>(p.out) + (size_t)b * p.ldo + col, o);
    }
}

extern int g_da_cfg;      // tuning knob: chunk*100 + stages*10 + rows (0 = default); 2 / 3 / 4 = tensor-core form with that many
                          // ring stages, 12 / 13 / 14 = the same with one box per head; -1 = the bulk-copy kernels for every call

#include "decode_attn_mma.cuh"

template <typename T, int CH, int NS, int ROWS>
static int launch_decode_attn_cfg(const DecAttnParams& p, int B, cudaStream_t st) {
    const size_t smem = (size_t)NS * 2 * CH * p.H * 64 * sizeof(T);
    if (smem > 227 * 1024) GCT_FAIL(GCT_ERR_UNSUPPORTED, "decode attention config needs %zu B of shared memory", smem);
    auto kern = decode_attn_kernel<T, CH, NS, ROWS>;
    GCT_SMEM_LIMIT(kern, smem);
    GCT_CUDA(launch_k(kern, dim3((B + ROWS - 1) / ROWS), dim3((p.H + 1) * 32), smem, st, true, p, B));
    return GCT_OK;
}

template <typename T>
static int launch_decode_attn(const DecAttnParams& p, int B, cudaStream_t st) {
    GCT_REQUIRE(p.H >= 1 && p.H <= 8, "decode attention: H=%d outside [1,8]", p.H);
    GCT_REQUIRE(p.n_cached + 1 <= DEC_MAX_KEYS, "decode attention: %d keys > %d", p.n_cached + 1, DEC_MAX_KEYS);
    GCT_REQUIRE(p.pitch == p.H * 64, "decode attention: cache rows must be contiguous (pitch %d != %d)", p.pitch, p.H * 64);
    if constexpr (sizeof(T) == 4) {
        return launch_decode_attn_cfg<T, 8, 3, 1>(p, B, st);
    } else {
        // self-attention over the growing cache: tensor-core form (decode_attn_mma.cuh).  Cross-attention in K / V form keeps the
        // bulk-copy kernel: its rows end at different keys, which a box clipped per launch cannot express.
        if (p.knew != nullptr && g_da_cfg >= 0 && g_da_cfg < 100) {
            const long long rows = p.rows_phys > 0 ? p.rows_phys : B;
            const int ns = g_da_cfg % 10, ph = g_da_cfg >= 10;
            if (ns == 3) return damma::launch_cfg<3>(p, B, rows, ph, st);
            if (ns == 4) return damma::launch_cfg<4>(p, B, rows, ph, st);
            return damma::launch_cfg<2>(p, B, rows, ph, st);
        }
        switch (g_da_cfg) {
            case 1631: return launch_decode_attn_cfg<T, 16, 3, 1>(p, B, st);
            case 1632: return launch_decode_attn_cfg<T, 16, 3, 2>(p, B, st);
            case 1634: return launch_decode_attn_cfg<T, 16, 3, 4>(p, B, st);
            case 861: return launch_decode_attn_cfg<T, 8, 6, 1>(p, B, st);
            case 862: return launch_decode_attn_cfg<T, 8, 6, 2>(p, B, st);
            case 1661: return launch_decode_attn_cfg<T, 16, 6, 1>(p, B, st);
            case 1662: return launch_decode_attn_cfg<T, 16, 6, 2>(p, B, st);
            case 3231: return launch_decode_attn_cfg<T, 32, 3, 1>(p, B, st);
            case 3232: return launch_decode_attn_cfg<T, 32, 3, 2>(p, B, st);
            case 1641: return launch_decode_attn_cfg<T, 16, 4, 1>(p, B, st);
            case 831: return launch_decode_attn_cfg<T, 8, 3, 1>(p, B, st);
            case 832: return launch_decode_attn_cfg<T, 8, 3, 2>(p, B, st);
            case 841: return launch_decode_attn_cfg<T, 8, 4, 1>(p, B, st);
            case 821: return launch_decode_attn_cfg<T, 8, 2, 1>(p, B, st);
            case 1621: return launch_decode_attn_cfg<T, 16, 2, 1>(p, B, st);
            case 1622: return launch_decode_attn_cfg<T, 16, 2, 2>(p, B, st);
            case 431: return launch_decode_attn_cfg<T, 4, 3, 1>(p, B, st);
            case 441: return launch_decode_attn_cfg<T, 4, 4, 1>(p, B, st);
            default:
                // 48 KB ring -> 4 CTAs/SM: best of the sweep up to ~65 cached keys; longer rows stream better in 16-key chunks
                // with a 2-stage ring ([B200] B=30000, 90 keys: 6.86 vs 6.57 TB/s; 49 keys: 6.01 vs 6.28)
                if (p.n_cached >= 68) return launch_decode_attn_cfg<T, 16, 2, 1>(p, B, st);
                return launch_decode_attn_cfg<T, 8, 3, 1>(p, B, st);
        }
    }
}

// x[b,:] = table[ys[b,pos]]*sqrt(d) + pe[pos + pe_off]; key_valid[b,pos] = tok != pad
__global__ void decode_embed_kernel(const int64_t* __restrict__ ys, int ys_stride, int pos, const float* __restrict__ table,
                                    int vocab, const float* __restrict__ pe, int pe_off, int d, float scale, int pad_id,
                                    float* __restrict__ x, uint8_t* __restrict__ key_valid, int kv_stride, int B,
                                    const int* __restrict__ rowmap) {
    // one warp per batch row, four rows per CTA (a CTA per row left the 30 000-row launch bound by CTA issue, not by its 61 MB)
    // rowmap (active-row decode): x is written at the compact row b, ys / key_valid belong to the physical row rowmap[b]
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    if (b >= B) return;
    const int bp = rowmap ? rowmap[b] : b;
    long long t = ys[(size_t)bp * ys_stride + pos];
    if (lane == 0) key_valid[(size_t)bp * kv_stride + pos] = (t != pad_id);
    if (t < 0 || t >= vocab) t = 0;
    const float* e = table + (size_t)t * d;
    const float* per = pe + (size_t)(pos + pe_off) * d;
    for (int c = lane * 4; c < d; c += 128) {
        float4 ev = *reinterpret_cast<const float4*>(e + c);
        float4 pv = *reinterpret_cast<const float4*>(per + c);
        *reinterpret_cast<float4*>(x + (size_t)b * d + c) =
            make_float4(ev.x * scale + pv.x, ev.y * scale + pv.y, ev.z * scale + pv.z, ev.w * scale + pv.w);
    }
}

struct SampleParams {
    const float* logits; int ld; int V;
    int64_t* ys; int ys_stride; int pos;         // writes ys[b, pos+1]
    const int64_t* forced;                        // optional [B] (or broadcast if forced_stride==0) prefix token
    int forced_stride;
    const float* uniforms;                        // optional [B] U(0,1) for this step
    uint32_t seed; int step;
    int greedy; int eos_id;
    uint8_t* done; int* n_done; int* first_all_done; int B;
    float* probs_out;                             // optional [B, V]
    float* logits_out;                            // optional [B, V]
    // active-row decode: logits rows are COMPACT (B of them); ys / done / uniforms / the counter hash / the all-done count use
    // the physical row rowmap[b] of Bphys; with skip_done a finished row appends pad_id and nothing else happens to it
    const int* rowmap; int Bphys; int skip_done; int pad_id;
};

// one warp per row, V <= 128.  Matches the reference's order of operations: softmax over the
// vocabulary, then torch.max (first maximal index) or a categorical draw.
__global__ void decode_sample_kernel(SampleParams p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_wait();
    pdl_launch_dependents();
    if (b >= p.B) return;
    const int bp = p.rowmap ? p.rowmap[b] : b;
    if (p.skip_done && p.done[bp]) {
        if (lane == 0) p.ys[(size_t)bp * p.ys_stride + p.pos + 1] = p.pad_id;
        return;
    }
    const float* lr = p.logits + (size_t)b * p.ld;
    float v[4];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = i * 32 + lane;
        v[i] = (c < p.V) ? lr[c] : -INFINITY;
        if (p.logits_out && c < p.V) p.logits_out[(size_t)b * p.V + c] = v[i];
        mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = (i * 32 + lane < p.V) ? expf(v[i] - mx) : 0.f; s += v[i]; }
    s = warp_sum(s);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = v[i] / s;
        if (p.probs_out && i * 32 + lane < p.V) p.probs_out[(size_t)b * p.V + i * 32 + lane] = v[i];
    }
    int tok;
    if (p.forced) {
        tok = (int)p.forced[(size_t)bp * p.forced_stride];
    } else if (p.greedy) {
        float best = -1.f; int bi = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = i * 32 + lane;
            if (c < p.V && v[i] > best) { best = v[i]; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        tok = bi;
    } else {
        float u;
        if (p.uniforms) u = p.uniforms[bp];
        else u = (float)(mix32(p.seed ^ mix32((uint32_t)p.step * 0x9e3779b9U + (uint32_t)bp)) >> 8) * (1.0f / 16777216.0f);
        // sequential cumulative sum in index order (same order as a CPU cumsum); threshold u*total
        // (only the 32-column blocks the vocabulary reaches: V = 27-32 for MOSES is one block)
        const int nblk = (p.V + 31) >> 5;
        float total = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nblk)
                for (int l = 0; l < 32; ++l) {
                    const float pv = __shfl_sync(0xffffffffu, v[i], l);
                    if (i * 32 + l < p.V) total += pv;
                }
        const float thr = u * total;
        float run = 0.f;
        int pick = -1;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nblk)
                for (int l = 0; l < 32; ++l) {
                    const float pv = __shfl_sync(0xffffffffu, v[i], l);
                    const int c = i * 32 + l;
                    if (c < p.V) {
                        run += pv;
                        if (pick < 0 && run > thr) pick = c;
                    }
                }
        if (pick < 0) pick = p.V - 1;
        tok = pick;
    }
    if (lane == 0) {
        p.ys[(size_t)bp * p.ys_stride + p.pos + 1] = tok;
        if (!p.forced && tok == p.eos_id && !p.done[bp]) {
            p.done[bp] = 1;
            const int n = atomicAdd(p.n_done, 1) + 1;
            if (n == p.Bphys) *p.first_all_done = p.step;
        }
    }
}

// Active-row decode: rowmap[0 .. n_active) = the rows whose `done` flag is clear, ascending; rowmap[n_active .. n_out) = the
// first finished row (a finished row costs the attention kernels nothing and appends pad_id), so that the batch the step
// kernels see can be rounded up to a size the host chooses.  One CTA: B <= a few 10^5 flags, once per chunk of steps.
__global__ void __launch_bounds__(1024) decode_compact_kernel(const uint8_t* __restrict__ done, int B, int* __restrict__ rowmap, int n_out) {
    __shared__ int wsum[32];
    __shared__ int base_s, first_done_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { base_s = 0; first_done_s = B; }
    __syncthreads();
    for (int i0 = 0; i0 < B; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool in = i < B;
        const bool act = in && done[i] == 0;
        if (in && !act) atomicMin(&first_done_s, i);
        const uint32_t bal = __ballot_sync(0xffffffffu, act);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (act) {
            const int slot = off + __popc(bal & ((1u << lane) - 1u));
            if (slot < n_out) rowmap[slot] = i;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += wsum[w]; base_s += t; }
        __syncthreads();
    }
    const int n_act = base_s;
    const int filler = first_done_s < B ? first_done_s : 0;
    for (int i = n_act + threadIdx.x; i < n_out; i += 1024) rowmap[i] = filler;
}
