// Memory-bound kernels: embedding + positional encoding, the reference's non-standard Norm,
// reparameterisation + KL, cross-entropy, mask builders, casts, column sums, Adam.
// All are coalesced / vectorised warp-shuffle kernels; HBM roofline applies.
#pragma once
#include "common.cuh"

// ==========================================================================================
// a2 + a3 (+ cond tokens of a11/a12): x[b,l,:] = value*sqrt(d) + pe[l,:], dropout
//   value = l < nc ? cond_W[l*d+c,:] . conds[b,:] + cond_B[l*d+c]  :  table[tok[b,l-nc], c]
// Reference: Model/modules.py:101-144, Model/cvaetf.py:36-43,98-112.
// ==========================================================================================
__global__ void embed_pe_kernel(const int64_t* __restrict__ tok, int Ltok, const float* __restrict__ table,
                                int vocab, const float* __restrict__ conds, const float* __restrict__ cond_W,
                                const float* __restrict__ cond_B, int nc, const float* __restrict__ pe,
                                float* __restrict__ out, int d, float scale, DropCtx drop) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int L = nc + Ltok;
    const int row = blockIdx.x;            // b*L + l
    const int b = row / L, l = row % L;
    float* o = out + (size_t)row * d;
    const float* per = pe + (size_t)l * d;
    if (l < nc) {
        float cv[8];
        for (int k = 0; k < nc; ++k) cv[k] = conds[(size_t)b * nc + k];
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            const float* w = cond_W + ((size_t)l * d + c) * nc;
            float v = cond_B[(size_t)l * d + c];
            for (int k = 0; k < nc; ++k) v = fmaf(w[k], cv[k], v);
            o[c] = drop_apply(drop, (uint64_t)row * d + c, v * scale + per[c]);
        }
    } else {
        long long t = tok[(size_t)b * Ltok + (l - nc)];
        if (t < 0 || t >= vocab) t = 0;    // out-of-range ids are a caller bug; stay in bounds
        const float* e = table + (size_t)t * d;
        for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
            float4 ev = *reinterpret_cast<const float4*>(e + c);
            float4 pv = *reinterpret_cast<const float4*>(per + c);
            float4 r;
            const uint64_t base = (uint64_t)row * d + c;
            r.x = drop_apply(drop, base + 0, ev.x * scale + pv.x);
            r.y = drop_apply(drop, base + 1, ev.y * scale + pv.y);
            r.z = drop_apply(drop, base + 2, ev.z * scale + pv.z);
            r.w = drop_apply(drop, base + 3, ev.w * scale + pv.w);
            *reinterpret_cast<float4*>(o + c) = r;
        }
    }
}

// backward of the above.  The table is tiny (<= 128 ids) and the rows are many (~40k): each block reduces a
// chunk of rows into a [vocab][128-column] tile in shared memory (thread = column, so no intra-block
// conflicts), then flushes the tile with one atomic per (id, column).  dx is the gradient w.r.t. `out`.
constexpr int EMB_BWD_ROWS = 128;
__global__ void embed_bwd_kernel(const int64_t* __restrict__ tok, int B, int Ltok, int nc, const float* __restrict__ dx,
                                 int d, float scale, DropCtx drop, float* __restrict__ dtable, int vocab) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    extern __shared__ float acc[];           // [vocab][128]
    const int c = blockIdx.y * 128 + threadIdx.x;
    for (int i = threadIdx.x; i < vocab * 128; i += 128) acc[i] = 0.f;
    __syncthreads();
    const int L = nc + Ltok;
    const long long p0 = (long long)blockIdx.x * EMB_BWD_ROWS, p1 = min((long long)B * Ltok, p0 + EMB_BWD_ROWS);
    if (c < d) {
        constexpr int U = 8;                 // rows per batch of independent loads
        for (long long pb = p0; pb < p1; pb += U) {
            int vv[U];
            float gg[U];
            size_t rr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long p = min(pb + u, p1 - 1);
                long long v = tok[p];
                if (v < 0 || v >= vocab) v = 0;
                vv[u] = (int)v;
                rr[u] = (size_t)(p / Ltok) * L + nc + (size_t)(p % Ltok);
                gg[u] = dx[rr[u] * d + c];
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (pb + u < p1) acc[vv[u] * 128 + threadIdx.x] += drop_apply(drop, (uint64_t)rr[u] * d + c, gg[u]);
        }
        for (int v = 0; v < vocab; ++v) {
            const float a = acc[v * 128 + threadIdx.x];
            if (a != 0.f) atomicAdd(dtable + (size_t)v * d + c, a * scale);
        }
    }
}

// cond-token linear backward: dW[l*d+c, k] += scale * sum_b dx[b,l,c]*conds[b,k]; dB[l*d+c] += scale*sum_b dx
// grid (nc, d/128, batch slices): each CTA reduces COND_BWD_BATCH batch rows and adds its partial atomically.
constexpr int COND_BWD_BATCH = 16;
__global__ void cond_embed_bwd_kernel(const float* __restrict__ dx, int B, int L, int nc, int d,
                                      const float* __restrict__ conds, float scale, DropCtx drop, int use_drop_index,
                                      float* __restrict__ dW, float* __restrict__ dB) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int l = blockIdx.x;
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= d) return;
    float aw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float ab = 0.f;
    const int b0 = blockIdx.z * COND_BWD_BATCH, b1 = min(B, b0 + COND_BWD_BATCH);
    float g[COND_BWD_BATCH];
#pragma unroll
    for (int i = 0; i < COND_BWD_BATCH; ++i) {          // all loads in flight before the first use
        const size_t row = (size_t)min(b0 + i, B - 1) * L + l;
        g[i] = dx[row * d + c];
    }
#pragma unroll
    for (int i = 0; i < COND_BWD_BATCH; ++i) {
        const int b = b0 + i;
        if (b < b1) {
            const size_t row = (size_t)b * L + l;
            float gv = g[i];
            if (use_drop_index) gv = drop_apply(drop, (uint64_t)row * d + c, gv);
            ab += gv;
            for (int k = 0; k < nc; ++k) aw[k] = fmaf(gv, conds[(size_t)b * nc + k], aw[k]);
        }
    }
    atomicAdd(dB + (size_t)l * d + c, ab * scale);
    for (int k = 0; k < nc; ++k) atomicAdd(dW + ((size_t)l * d + c) * nc + k, aw[k] * scale);
}

// cond2lat tokens written into the decoder memory: mem[b, j, :] = W[j*d+c,:].conds[b] + B[j*d+c]
template <typename T>
__global__ void cond_tokens_kernel(const float* __restrict__ conds, const float* __restrict__ W,
                                   const float* __restrict__ Bv, int nc, int d, T* __restrict__ mem, int Lmem) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int b = blockIdx.x / nc, j = blockIdx.x % nc;
    T* o = mem + ((size_t)b * Lmem + j) * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float* w = W + ((size_t)j * d + c) * nc;
        float v = Bv[(size_t)j * d + c];
        for (int k = 0; k < nc; ++k) v = fmaf(w[k], conds[(size_t)b * nc + k], v);
        o[c] = from_f<T>(v);
    }
}

// ==========================================================================================
// a1 Norm: y = alpha*(x-mean)/(std_unbiased+eps)+bias     (Model/modules.py:80-95)
// one warp per row; d % 128 == 0, d <= 1024.  Optionally also writes an fp32 copy (the encoder
// keeps the *normalised* value as its residual stream, Model/layers.py:23,28).
// ==========================================================================================
template <typename T, int NV /* float4 per lane */>
__global__ void norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha,
                                const float* __restrict__ bias, T* __restrict__ y, float* __restrict__ y32,
                                int rows, float eps) {
    const int d = NV * 128;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_wait();
    pdl_launch_dependents();
    if (row >= rows) return;
    const float* xr = x + (size_t)row * d;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float stdv = sqrtf(warp_sum(q) / (d - 1));
    const float r = 1.f / (stdv + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 a = *reinterpret_cast<const float4*>(alpha + c);
        float4 b = *reinterpret_cast<const float4*>(bias + c);
        float4 o;
        o.x = a.x * v[i].x * r + b.x; o.y = a.y * v[i].y * r + b.y;
        o.z = a.z * v[i].z * r + b.z; o.w = a.w * v[i].w * r + b.w;
        if (y32) *reinterpret_cast<float4*>(y32 + (size_t)row * d + c) = o;
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)row * d + c) = o;
        } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + (size_t)row * d + c) = u;
        }
    }
}

// Norm backward.  dx = r*(g - mean(g)) - c*xc,  g = dy*alpha, c = r^2*sum(g*xc)/((d-1)*std)
//   (+ `add` : gradient arriving on the residual path).  dalpha/dbias accumulated with atomics.
// Optional fused consumer prologue: dropT = T(dropmask(dc) * dx) and its column sums (the bias gradient of the
// projection whose output carried that dropout) -- the operand the next dgrad / wgrad GEMMs read, so the fp32 dx is
// not re-read by a separate cast kernel.
// The per-column partial sums (dalpha, dbias, dropsum) live in a per-warp slab of shared memory, not in registers: each lane
// owns its 4*NV columns of the slab (float4 accesses, no conflicts, no atomics), which keeps the kernel under 80 registers
// so that 3-4 CTAs are resident per SM (ncu, round 1: 125 registers, 25 % occupancy, 4.9 TB/s).
template <typename T, int NV>
__global__ void __launch_bounds__(256, (NV <= 4) ? 3 : 1)
norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha,
                                const float* __restrict__ dy, const float* __restrict__ add,
                                float* __restrict__ dx, float* __restrict__ dalpha, float* __restrict__ dbias,
                                int rows, float eps, T* __restrict__ dropT, DropCtx dc, float* __restrict__ dropsum) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int d = NV * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    extern __shared__ float sm[];              // [nwarp][3][d] per-warp partials
    float* mine = sm + (size_t)warp * 3 * d;
#pragma unroll
    for (int i = 0; i < 3 * NV; ++i) *reinterpret_cast<float4*>(mine + (i * 32 + lane) * 4) = make_float4(0, 0, 0, 0);
    for (int row = blockIdx.x * nwarp + warp; row < rows; row += gridDim.x * nwarp) {
        const float* xr = x + (size_t)row * d;
        const float* gr = dy + (size_t)row * d;
        float4 v[NV], g[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
            g[i] = *reinterpret_cast<const float4*>(gr + (i * 32 + lane) * 4);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) / d;
        float q = 0.f, sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        const float stdv = sqrtf(warp_sum(q) / (d - 1));
        const float r = 1.f / (stdv + eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            // parameter grads use the raw dy
            float4* pa = reinterpret_cast<float4*>(mine + (i * 32 + lane) * 4);
            float4* pb = reinterpret_cast<float4*>(mine + d + (i * 32 + lane) * 4);
            float4 a4 = *pa, b4 = *pb;
            b4.x += g[i].x; b4.y += g[i].y; b4.z += g[i].z; b4.w += g[i].w;
            a4.x += g[i].x * v[i].x * r; a4.y += g[i].y * v[i].y * r;
            a4.z += g[i].z * v[i].z * r; a4.w += g[i].w * v[i].w * r;
            *pa = a4; *pb = b4;
            const float4 al = __ldg(reinterpret_cast<const float4*>(alpha + (i * 32 + lane) * 4));      // 2 KB, L1-resident
            g[i].x *= al.x; g[i].y *= al.y; g[i].z *= al.z; g[i].w *= al.w;
            sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
        }
        const float mg = warp_sum(sg) / d;
        const float c = (stdv > 0.f) ? r * r * warp_sum(sgx) / ((d - 1) * stdv) : 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const size_t off = (size_t)row * d + (i * 32 + lane) * 4;
            float4 o;
            o.x = r * (g[i].x - mg) - c * v[i].x; o.y = r * (g[i].y - mg) - c * v[i].y;
            o.z = r * (g[i].z - mg) - c * v[i].z; o.w = r * (g[i].w - mg) - c * v[i].w;
            if (add) {
                float4 a = *reinterpret_cast<const float4*>(add + off);
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            *reinterpret_cast<float4*>(dx + off) = o;
            if (dropT) {
                drop_pair(dc, (uint32_t)(off >> 1), o.x, o.y);           // identical decisions to drop_apply(off + k)
                drop_pair(dc, (uint32_t)(off >> 1) + 1, o.z, o.w);
                if constexpr (sizeof(T) == 4) {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(dropT) + off) = o;
                } else {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
                    uint2 u;
                    u.x = *reinterpret_cast<uint32_t*>(&p0);
                    u.y = *reinterpret_cast<uint32_t*>(&p1);
                    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(dropT) + off) = u;
                }
                float4* ps = reinterpret_cast<float4*>(mine + 2 * d + (i * 32 + lane) * 4);
                float4 s4 = *ps;
                s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w;
                *ps = s4;
            }
        }
    }
    // reduce the per-warp slabs, then one atomic per column
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float a = 0.f, b = 0.f, s2 = 0.f;
        for (int w = 0; w < nwarp; ++w) {
            a += sm[(size_t)w * 3 * d + c];
            b += sm[(size_t)w * 3 * d + d + c];
            s2 += sm[(size_t)w * 3 * d + 2 * d + c];
        }
        atomicAdd(dalpha + c, a);
        atomicAdd(dbias + c, b);
        if (dropsum) atomicAdd(dropsum + c, s2);
    }
}

// ==========================================================================================
// a8 reparameterisation (Model/sublayers.py:11-18, Model/cvaetf.py:63-69) on the fused
// [rows, 2*lat] head output:  mu | log_var -> z = mu + eps*exp(0.5*log_var).
// Also emits the padded decoder-memory operand zpad[b, off + s, :] (type T) used by fc_z.
// ==========================================================================================
template <typename T>
__global__ void reparam_fwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps, int rows, int lat,
                                   int Se, int Sm, float* __restrict__ mu, float* __restrict__ lv,
                                   float* __restrict__ z, T* __restrict__ zpad) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * lat) return;
    const size_t row = i / lat;
    const int c = (int)(i % lat);
    const float m = mulv[row * 2 * lat + c], l = mulv[row * 2 * lat + lat + c];
    const float zz = eps ? fmaf(eps[i], __expf(0.5f * l), m) : m;
    mu[i] = m; lv[i] = l; z[i] = zz;
    if (zpad) {
        const size_t b = row / Se, s = row % Se;
        zpad[((b * Sm) + (Sm - Se) + s) * lat + c] = from_f<T>(zz);
    }
}

// d(mulv) from dz (decoder path) + external dmu/dlv (loss path, may be null) :
//   dmu_tot = dz + dmu ;  dlv_tot = dz*eps*0.5*exp(0.5 lv) + dlv
template <typename T>
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ dmu_ext,
                                   const float* __restrict__ dlv_ext, const float* __restrict__ eps,
                                   const float* __restrict__ lv, int rows, int lat, T* __restrict__ dmulv) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * lat) return;
    const size_t row = i / lat;
    const int c = (int)(i % lat);
    const float g = dz ? dz[i] : 0.f;
    float dm = g + (dmu_ext ? dmu_ext[i] : 0.f);
    float dl = (dlv_ext ? dlv_ext[i] : 0.f);
    if (eps) dl += g * eps[i] * 0.5f * __expf(0.5f * lv[i]);
    dmulv[row * 2 * lat + c] = from_f<T>(dm);
    dmulv[row * 2 * lat + lat + c] = from_f<T>(dl);
}

// ==========================================================================================
// a15 loss (Train/trainer1.py:19-30): per-row CE with ignore_index, KL partials, final reduce.
// ==========================================================================================
// one warp per row, V <= 128.  row_loss[r] = ignored ? 0 : logsumexp - logit[target]
// dlogits (optional) = gscale * (softmax - onehot) or 0 for ignored rows.
__global__ void ce_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int rows, int V,
                               int ld, int pad_id, float* __restrict__ row_loss, float* __restrict__ dlogits,
                               float gscale) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* lr = logits + (size_t)row * ld;
    float v[4];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = i * 32 + lane;
        v[i] = (c < V) ? lr[c] : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (i * 32 + lane < V) ? expf(v[i] - mx) : 0.f;
    s = warp_sum(s);
    const long long t = target[row];
    const bool ign = (t == pad_id) || t < 0 || t >= V;
    const float lse = mx + logf(s);
    if (lane == 0) row_loss[row] = ign ? 0.f : (lse - lr[t]);
    if (dlogits) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = i * 32 + lane;
            if (c < V) {
                float p = expf(v[i] - lse);
                dlogits[(size_t)row * ld + c] = ign ? 0.f : gscale * (p - (c == t ? 1.f : 0.f));
            }
        }
    }
}

// KL partial sums: part[block] = sum over its slice of -0.5*(1 + lv - mu^2 - exp(lv))
__global__ void kl_partial_kernel(const float* __restrict__ mu, const float* __restrict__ lv, size_t n,
                                  float* __restrict__ part) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    __shared__ float sm[32];
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mu[i], l = lv[i];
        acc += -0.5f * (1.f + l - m * m - expf(l));
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// deterministic single-block sum of n floats -> out[0] (double accumulation)
__global__ void final_sum_kernel(const float* __restrict__ in, size_t n, float* __restrict__ out) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    __shared__ double sm[32];
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)in[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double r = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (threadIdx.x == 0) out[0] = (float)r;
    }
}

// Property head of use_cond2dec models (Model/cvaetf.py:184-186, Train/trainer1.py:24-26): prop[b,i] = logits[b,i,:] . w + b0 on the
// first nc rows of every sample, RCE_prop = sum (prop - y)^2, and its backward: dlogits[b,i,:] = 2 g (prop - y) w (those rows carry
// no cross-entropy term), dw += 2 g (prop - y) logits[b,i,:], db += 2 g (prop - y).  One warp per (b, i) row, V <= 128.
__global__ void prop_head_kernel(const float* __restrict__ logits, int B, int Ld, int nc, int V, const float* __restrict__ w,
                                 const float* __restrict__ b0, const float* __restrict__ target, float gscale, float* __restrict__ prop_out,
                                 float* __restrict__ out4, float* __restrict__ dlogits, float* __restrict__ dw, float* __restrict__ db) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= B * nc) return;
    const int b = r / nc, i = r % nc;
    const float* lr = logits + ((size_t)b * Ld + i) * V;
    float v[4], wv[4], dot = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = k * 32 + lane;
        v[k] = (c < V) ? lr[c] : 0.f;
        wv[k] = (c < V) ? w[c] : 0.f;
        dot = fmaf(v[k], wv[k], dot);
    }
    const float p = warp_sum(dot) + b0[0];
    const float diff = p - target[r];
    if (lane == 0) {
        if (prop_out) prop_out[r] = p;
        if (out4) { atomicAdd(out4 + 0, diff * diff); atomicAdd(out4 + 2, diff * diff); }
    }
    if (!dlogits) return;
    const float g = 2.f * gscale * diff;
    float* dr = dlogits + ((size_t)b * Ld + i) * V;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = k * 32 + lane;
        if (c < V) {
            dr[c] = g * wv[k];
            atomicAdd(dw + c, g * v[k]);
        }
    }
    if (lane == 0) atomicAdd(db, g);
}

// d(KL)/dmu = beta*mu ; d(KL)/dlv = beta*0.5*(exp(lv)-1)   (scaled by upstream gscale)
__global__ void kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, size_t n, float g,
                              float* __restrict__ dmu, float* __restrict__ dlv) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dmu[i] = g * mu[i];
    dlv[i] = g * 0.5f * (expf(lv[i]) - 1.f);
}

// ==========================================================================================
// a4 masks (Model/modules.py:10-66) as byte arrays.
// ==========================================================================================
// key mask [B, nc+L]: first nc entries 1, then tok != pad
__global__ void src_mask_kernel(const int64_t* __restrict__ tok, int B, int L, int nc, int pad, uint8_t* __restrict__ m) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = nc + L;
    if (i >= B * W) return;
    const int b = i / W, j = i % W;
    m[i] = (j < nc) ? 1 : (tok[(size_t)b * L + (j - nc)] != pad);
}
// dense target mask [B, nc+T, nc+T] (nc = 0 unless cond2dec): keypad(j) & nopeak(i,j)
__global__ void trg_mask_kernel(const int64_t* __restrict__ tok, int B, int T, int nc, int pad, uint8_t* __restrict__ m) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    const int W = nc + T;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * W * W) return;
    const int j = (int)(i % W), q = (int)((i / W) % W), b = (int)(i / ((size_t)W * W));
    const bool key_ok = (j < nc) ? true : (tok[(size_t)b * T + (j - nc)] != pad);
    bool peek;
    if (q < nc) peek = (j < nc) || (j == nc);            // cond rows: all cond cols + first target col
    else peek = (j < nc) || ((j - nc) <= (q - nc));      // target rows: cond cols + causal
    m[i] = key_ok && peek;
}
// generic "nonzero" cast of a caller-supplied mask (bool/uint8: esize 1, int32: 4, int64: 8)
__global__ void mask_cast_kernel(const void* __restrict__ in, int esize, size_t n, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool nz;
    if (esize == 1) nz = reinterpret_cast<const uint8_t*>(in)[i] != 0;
    else if (esize == 4) nz = reinterpret_cast<const int32_t*>(in)[i] != 0;
    else nz = reinterpret_cast<const int64_t*>(in)[i] != 0;
    out[i] = nz;
}

// ==========================================================================================
// casts / column sums / optimiser
// ==========================================================================================
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
#pragma unroll
        for (int k = 0; k < 4; ++k) out[i + k] = from_f<TO>(to_f(in[i + k]));
    } else {
        for (size_t k = i; k < n; ++k) out[k] = from_f<TO>(to_f(in[k]));
    }
}

// out[r,c] = T(dropmask(site, r*cols + c) * in[r,c]); optionally accumulates column sums (bias grad).
// Block = 64 column-threads (4 columns each -> 256 columns) x 4 row groups; the row groups are reduced through
// shared memory so each block issues one atomic per column.
template <typename T>
__global__ void __launch_bounds__(256)
cast_drop_colsum_kernel(const float* __restrict__ in, T* __restrict__ out, int rows, int cols, DropCtx drop,
                        float* __restrict__ colsum) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    __shared__ float red[4][256];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int c = (blockIdx.y * 64 + tx) * 4;
    float4 acc = make_float4(0, 0, 0, 0);
    if (c < cols) {
#pragma unroll 2
        for (int r = blockIdx.x * 4 + ty; r < rows; r += gridDim.x * 4) {
            const size_t off = (size_t)r * cols + c;
            float4 v = *reinterpret_cast<const float4*>(in + off);
            v.x = drop_apply(drop, off + 0, v.x); v.y = drop_apply(drop, off + 1, v.y);
            v.z = drop_apply(drop, off + 2, v.z); v.w = drop_apply(drop, off + 3, v.w);
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off) = v;
            } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
                uint2 u;
                u.x = *reinterpret_cast<uint32_t*>(&p0);
                u.y = *reinterpret_cast<uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out) + off) = u;
            }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (!colsum) return;
    red[ty][tx * 4 + 0] = acc.x; red[ty][tx * 4 + 1] = acc.y; red[ty][tx * 4 + 2] = acc.z; red[ty][tx * 4 + 3] = acc.w;
    __syncthreads();
    const int cc = blockIdx.y * 256 + threadIdx.x;
    if (cc < cols) atomicAdd(colsum + cc, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// colsum[c] += sum_r in[r, c]  (leading dimension ld).  Block = 32 column-threads (8 columns each -> 256 columns)
// x 8 row groups, reduced through shared memory: one atomic per column per block.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ in, int rows, int cols, int ld, float* __restrict__ colsum) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    __shared__ float red[8][256];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.y * 32 + tx) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c + 8 <= cols && (ld & 7) == 0) {
#pragma unroll 4
        for (int r = blockIdx.x * 8 + ty; r < rows; r += gridDim.x * 8) {
            f8 v = ld8(in + (size_t)r * ld + c);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v.v[e];
        }
    } else if (c < cols) {
        for (int r = blockIdx.x * 8 + ty; r < rows; r += gridDim.x * 8)
            for (int e = 0; e < 8 && c + e < cols; ++e) acc[e] += to_f(in[(size_t)r * ld + c + e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[ty][tx * 8 + e] = acc[e];
    __syncthreads();
    const int cc = blockIdx.y * 256 + threadIdx.x;
    if (cc < cols) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += red[g][threadIdx.x];
        atomicAdd(colsum + cc, s);
    }
}

// Fused Adam over the flat parameter buffer (torch.optim.Adam semantics, train1.py:116-119),
// with the 1/world_size gradient averaging folded in and the bf16 shadow refreshed in place.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, bf16* __restrict__ shadow, size_t n, float lr, float b1, float b2,
                            float eps, float bc1, float bc2_sqrt, float gscale, int vec16) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    // four consecutive elements per thread: 16-byte loads / stores when the buffers allow it (vec16), the same arithmetic per element
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    float pv[4], gv[4], mv[4], vv[4];
    const bool full = vec16 && i0 + 3 < n;
    if (full) {
        const float4 a = *reinterpret_cast<const float4*>(p + i0), b = *reinterpret_cast<const float4*>(g + i0);
        const float4 c = *reinterpret_cast<const float4*>(m + i0), e = *reinterpret_cast<const float4*>(v + i0);
        pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
        mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w; vv[0] = e.x; vv[1] = e.y; vv[2] = e.z; vv[3] = e.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < n) { pv[k] = p[i0 + k]; gv[k] = g[i0 + k]; mv[k] = m[i0 + k]; vv[k] = v[i0 + k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float gr = gv[k] * gscale;
        const float mi = b1 * mv[k] + (1.f - b1) * gr;
        const float vi = b2 * vv[k] + (1.f - b2) * gr * gr;
        mv[k] = mi; vv[k] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pv[k] = pv[k] - (lr / bc1) * (mi / denom);
    }
    if (full) {
        *reinterpret_cast<float4*>(m + i0) = make_float4(mv[0], mv[1], mv[2], mv[3]);
        *reinterpret_cast<float4*>(v + i0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        *reinterpret_cast<float4*>(p + i0) = make_float4(pv[0], pv[1], pv[2], pv[3]);
        if (shadow) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pv[0], pv[1]), hi = __floats2bfloat162_rn(pv[2], pv[3]);
            *reinterpret_cast<uint2*>(shadow + i0) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < n) {
                m[i0 + k] = mv[k]; v[i0 + k] = vv[k]; p[i0 + k] = pv[k];
                if (shadow) shadow[i0 + k] = __float2bfloat16_rn(pv[k]);
            }
    }
}

// ==========================================================================================
// input pipeline: batch assembly from the pre-tokenised corpus (Model/collate_fn.py + Field.process)
// one CTA per batch row; int16 ids in, int64 padded rows out (the layout forward_propagation expects)
// ==========================================================================================
struct CollateParams {
    const int16_t* src_ids; const int16_t* trg_ids; const int64_t* tok_off;
    const int16_t* sca_src_ids; const int16_t* sca_trg_ids; const int64_t* sca_off;
    const float* econds; const float* dconds; int nconds; long long n_rows;
    const int64_t* rows; int B, S, T;
    int pad_src, pad_trg, sos, eos, sep_src, sep_trg;
    int64_t* src; int64_t* trg; float* econds_out; float* dconds_out;
};
__global__ void collate_kernel(CollateParams p) {
    const int b = blockIdx.x;
    long long r = p.rows[b];
    if (r < 0 || r >= p.n_rows) r = 0;                    // out-of-range row ids are a caller bug; stay in bounds
    const long long t0 = p.tok_off[r];
    const int nt = (int)(p.tok_off[r + 1] - t0);
    const bool sca = p.sep_src >= 0 && p.sca_off != nullptr;
    const long long s0 = sca ? p.sca_off[r] : 0;
    const int ns = sca ? (int)(p.sca_off[r + 1] - s0) : 0;
    const int pre = sca ? ns + 1 : 0;                     // scaffold tokens + <sep>
    if (p.src) {
        int64_t* o = p.src + (size_t)b * p.S;
        for (int j = threadIdx.x; j < p.S; j += blockDim.x) {
            int64_t v = p.pad_src;
            if (j < ns) v = p.sca_src_ids[s0 + j];
            else if (sca && j == ns) v = p.sep_src;
            else if (j - pre < nt) v = p.src_ids[t0 + (j - pre)];
            o[j] = v;
        }
    }
    if (p.trg) {
        int64_t* o = p.trg + (size_t)b * p.T;
        for (int j = threadIdx.x; j < p.T; j += blockDim.x) {
            const int k = j - 1;                          // position after <sos>
            int64_t v = p.pad_trg;
            if (j == 0) v = p.sos;
            else if (k < ns) v = p.sca_trg_ids[s0 + k];
            else if (sca && k == ns) v = p.sep_trg;
            else if (k - pre < nt) v = p.trg_ids[t0 + (k - pre)];
            else if (k - pre == nt) v = p.eos;
            o[j] = v;
        }
    }
    if (threadIdx.x < p.nconds) {
        if (p.econds_out) p.econds_out[(size_t)b * p.nconds + threadIdx.x] = p.econds[(size_t)r * p.nconds + threadIdx.x];
        if (p.dconds_out) p.dconds_out[(size_t)b * p.nconds + threadIdx.x] = p.dconds[(size_t)r * p.nconds + threadIdx.x];
    }
}
