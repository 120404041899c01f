// KV-cached autoregressive decode kernels (replace the O(T^2) loop of
// Inference/sampling_tool.py:140-184): single-query attention over a cache, and the on-device
// next-token sampler (softmax -> argmax / inverse-CDF multinomial -> append -> <eos> bookkeeping).
// HBM-bound: each (batch, head) streams its K/V rows once with 16-byte loads.
#pragma once
#include "common.cuh"

struct DecAttnParams {
    const void* q; int ldq;                  // [B, ldq], head h at column h*64
    const void* knew; const void* vnew; int ldnew;   // this step's K/V rows (self-attention) or null
    void* kcache; void* vcache;              // [B, Lmax, pitch] ; pitch in elements
    long long cache_bstride; int pitch;
    int n_cached;                            // keys already in the cache (self: pos ; cross: Sm)
    const uint8_t* key_valid; int kv_stride; // [B, kv_stride] 1 = attend (0 -> score -1e9)
    void* out; int ldo;                      // [B, ldo]
    int H; float scale;
};

// one warp per (b, h); lane = (g = lane/8 : key slot, sub = lane%8 : 8-dim slice)
template <typename T>
__global__ void __launch_bounds__(256)
decode_attn_kernel(DecAttnParams p) {
    const int b = blockIdx.x;
    const int h = threadIdx.x >> 5;
    if (h >= p.H) return;
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    const int col = h * 64 + sub * 8;
    const f8 q = ld8(reinterpret_cast<const T*>(p.q) + (size_t)b * p.ldq + col);
    T* kc = reinterpret_cast<T*>(p.kcache) + (size_t)b * p.cache_bstride + col;
    T* vc = reinterpret_cast<T*>(p.vcache) + (size_t)b * p.cache_bstride + col;
    const uint8_t* valid = p.key_valid + (size_t)b * p.kv_stride;
    int nkeys = p.n_cached;
    f8 kn, vn;
    if (p.knew) {
        kn = ld8(reinterpret_cast<const T*>(p.knew) + (size_t)b * p.ldnew + col);
        vn = ld8(reinterpret_cast<const T*>(p.vnew) + (size_t)b * p.ldnew + col);
        if (g == 0) {       // append this step's K/V row to the cache
            st8(kc + (size_t)p.n_cached * p.pitch, kn);
            st8(vc + (size_t)p.n_cached * p.pitch, vn);
        }
        nkeys += 1;
    }
    float m = -INFINITY, l = 0.f;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const unsigned gmask = 0xFFu << (g * 8);      // the 8 lanes that share one key
#pragma unroll 2
    for (int j0 = 0; j0 < nkeys; j0 += 4) {
        const int j = j0 + g;
        if (j < nkeys) {     // uniform within each 8-lane group
            f8 kk, vv;
            if (p.knew && j == p.n_cached) { kk = kn; vv = vn; }
            else { kk = ld8(kc + (size_t)j * p.pitch); vv = ld8(vc + (size_t)j * p.pitch); }
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) s = fmaf(q.v[e], kk.v[e], s);
            s += __shfl_xor_sync(gmask, s, 1);
            s += __shfl_xor_sync(gmask, s, 2);
            s += __shfl_xor_sync(gmask, s, 4);
            s = valid[j] ? s * p.scale : -1e9f;
            const float mn = fmaxf(m, s);
            const float corr = __expf(m - mn), pj = __expf(s - mn);
            l = l * corr + pj;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vv.v[e], acc[e] * corr);
            m = mn;
        }
    }
    __syncwarp();
    // merge the four key slots
    float M = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 16));
    const float f = (m == -INFINITY) ? 0.f : __expf(m - M);
    l *= f;
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
    const float inv = 1.f / l;
    f8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float a = acc[e] * f;
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 16);
        o.v[e] = a * inv;
    }
    if (g == 0) st8(reinterpret_cast<T*>(p.out) + (size_t)b * p.ldo + col, o);
}

// x[b,:] = table[ys[b,pos]]*sqrt(d) + pe[pos + pe_off]; key_valid[b,pos] = tok != pad
__global__ void decode_embed_kernel(const int64_t* __restrict__ ys, int ys_stride, int pos, const float* __restrict__ table,
                                    int vocab, const float* __restrict__ pe, int pe_off, int d, float scale, int pad_id,
                                    float* __restrict__ x, uint8_t* __restrict__ key_valid, int kv_stride) {
    const int b = blockIdx.x;
    long long t = ys[(size_t)b * ys_stride + pos];
    if (threadIdx.x == 0) key_valid[(size_t)b * kv_stride + pos] = (t != pad_id);
    if (t < 0 || t >= vocab) t = 0;
    const float* e = table + (size_t)t * d;
    const float* per = pe + (size_t)(pos + pe_off) * d;
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
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
};

// one warp per row, V <= 128.  Matches the reference's order of operations: softmax over the
// vocabulary, then torch.max (first maximal index) or a categorical draw.
__global__ void decode_sample_kernel(SampleParams p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= p.B) return;
    const float* lr = p.logits + (size_t)b * p.ld;
    float v[4];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = i * 32 + lane;
        v[i] = (c < p.V) ? lr[c] : -INFINITY;
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
        tok = (int)p.forced[(size_t)b * p.forced_stride];
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
        if (p.uniforms) u = p.uniforms[b];
        else u = (float)(mix32(p.seed ^ mix32((uint32_t)p.step * 0x9e3779b9U + (uint32_t)b)) >> 8) * (1.0f / 16777216.0f);
        // sequential cumulative sum in index order (same order as a CPU cumsum); threshold u*total
        float total = 0.f;
        for (int i = 0; i < 4; ++i)
            for (int l = 0; l < 32; ++l) {
                const float pv = __shfl_sync(0xffffffffu, v[i], l);
                if (i * 32 + l < p.V) total += pv;
            }
        const float thr = u * total;
        float run = 0.f;
        int pick = -1;
        for (int i = 0; i < 4; ++i)
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
        p.ys[(size_t)b * p.ys_stride + p.pos + 1] = tok;
        if (!p.forced && tok == p.eos_id && !p.done[b]) {
            p.done[b] = 1;
            const int n = atomicAdd(p.n_done, 1) + 1;
            if (n == p.B) *p.first_all_done = p.step;
        }
    }
}
