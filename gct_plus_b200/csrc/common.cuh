// Shared device/host helpers for the gct_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#ifndef __CUDA_ARCH__
#define GCT_HOST_SIDE 1
#endif

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns 0 / negative, message in a thread-local buffer
// ------------------------------------------------------------------------------------------
#define GCT_OK 0
#define GCT_ERR_ARG (-1)
#define GCT_ERR_CUDA (-2)
#define GCT_ERR_UNSUPPORTED (-3)

extern thread_local char g_gct_err[512];

#define GCT_FAIL(code, ...)                                   \
    do {                                                      \
        snprintf(g_gct_err, sizeof(g_gct_err), __VA_ARGS__);  \
        return (code);                                        \
    } while (0)

#define GCT_REQUIRE(cond, ...)                                \
    do {                                                      \
        if (!(cond)) GCT_FAIL(GCT_ERR_ARG, __VA_ARGS__);      \
    } while (0)

#define GCT_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            GCT_FAIL(GCT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,                 \
                     cudaGetErrorString(_e));                                                   \
    } while (0)

#define GCT_LAUNCH_CHECK() GCT_CUDA(cudaGetLastError())

// Function attributes (dynamic shared-memory limit, carve-out) and the SM count belong to a DEVICE, not to the process:
// the once-flags below are kept per device ordinal so that a process driving several GPUs sets them on each.
constexpr int GCT_MAX_DEVICES = 64;
static inline int gct_cur_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < GCT_MAX_DEVICES) ? d : 0;
}
struct PerDeviceSize {
    size_t v[GCT_MAX_DEVICES] = {};
    size_t& cur() { return v[gct_cur_device()]; }
};
// raises the kernel's dynamic shared-memory limit to `bytes` on the current device if it is below that
#define GCT_SMEM_LIMIT(kern, bytes)                                                                              \
    do {                                                                                                         \
        static PerDeviceSize lim_;                                                                               \
        size_t& cur_ = lim_.cur();                                                                               \
        if ((size_t)(bytes) > cur_) {                                                                            \
            GCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
            cur_ = (size_t)(bytes);                                                                              \
        }                                                                                                        \
    } while (0)

#define GCT_TRY(expr)                 \
    do {                              \
        int _r = (expr);              \
        if (_r != GCT_OK) return _r;  \
    } while (0)

// ------------------------------------------------------------------------------------------
// type helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

// 8-wide vector load/store (bf16: 16 B, float: 2 x 16 B); pointers must be 16 B aligned
struct f8 { float v[8]; };

__device__ __forceinline__ f8 ld8(const float* p) {
    f8 r;
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ f8 ld8(const bf16* p) {
    f8 r;
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ void st8(float* p, const f8& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const f8& r) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum for blockDim.x <= 1024 (result valid in every thread)
__device__ __forceinline__ float block_sum(float v, float* smem32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    float r = (lane < nw) ? smem32[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// exact-erf GELU (Model/sublayers.py:86 uses F.gelu default) and its derivative
__device__ __noinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __noinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// bf16 tier: Phi(x) = 0.5 erfc(-x/sqrt2) by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below bf16 resolution) with
// the 1/sqrt2 and the 0.5 folded into the constants:  t = 1/(1 + p'|x|), e = exp(-x^2/2), Phi(-|x|) = poly(t) t e.
// The same two MUFU results (t, e) give both gelu(x) = x Phi(x) and gelu'(x) = Phi(x) + x e / sqrt(2 pi).
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define GCT_AS_P  0.23164189f           /* 0.3275911 / sqrt(2) */
#define GCT_AS_A1 0.127414796f          /* 0.5 * 0.254829592 */
#define GCT_AS_A2 (-0.142248368f)
#define GCT_AS_A3 0.7107068705f
#define GCT_AS_A4 (-0.7265760135f)
#define GCT_AS_A5 0.5307027145f
#define GCT_NEG_HALF_LOG2E (-0.72134752044f)
#define GCT_INV_SQRT_2PI 0.39894228040143268f
__device__ __forceinline__ void gelu_phi_e(float x, float& phi, float& e) {
    const float t = rcp_approx(fmaf(GCT_AS_P, fabsf(x), 1.f));
    e = ex2_approx(x * x * GCT_NEG_HALF_LOG2E);
    float p = fmaf(GCT_AS_A5, t, GCT_AS_A4);
    p = fmaf(p, t, GCT_AS_A3);
    p = fmaf(p, t, GCT_AS_A2);
    p = fmaf(p, t, GCT_AS_A1);
    const float h = p * (t * e);                      // Phi(-|x|)
    phi = 0.5f + copysignf(0.5f - h, x);
}
__device__ __forceinline__ float gelu_fast(float x) { float phi, e; gelu_phi_e(x, phi, e); return x * phi; }
__device__ __forceinline__ float gelu_fast_grad(float x) {
    float phi, e; gelu_phi_e(x, phi, e);
    return fmaf(x * e, GCT_INV_SQRT_2PI, phi);
}
// Packed fp32x2 evaluation (FFMA2 / FMUL2 / FADD2: one instruction per element PAIR on sm_100).  m = per-element
// multipliers applied to both results (dropout keep * 1/(1-p), or 1):  g = m * gelu(x),  dg = m * gelu'(x).
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
// NP independent pairs, written stage by stage so that the dependent FFMA2 chains of different pairs can interleave in
// the instruction stream (a lone Horner chain stalls ~4 cycles per step on its own result).
// WITH_GRAD = false skips dg.
template <int NP, bool WITH_GRAD>
__device__ __forceinline__ void gelu_pairs(const float2* x, const float2* m, float2* g, float2* dg) {
    float2 t[NP], e[NP], p[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = __ffma2_rn(make_float2(fabsf(x[i].x), fabsf(x[i].y)), f2(GCT_AS_P), f2(1.f));
#pragma unroll
    for (int i = 0; i < NP; ++i) e[i] = __fmul2_rn(__fmul2_rn(x[i], x[i]), f2(GCT_NEG_HALF_LOG2E));
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = make_float2(rcp_approx(t[i].x), rcp_approx(t[i].y));
#pragma unroll
    for (int i = 0; i < NP; ++i) e[i] = make_float2(ex2_approx(e[i].x), ex2_approx(e[i].y));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = __ffma2_rn(t[i], f2(GCT_AS_A5), f2(GCT_AS_A4));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = __ffma2_rn(p[i], t[i], f2(GCT_AS_A3));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = __ffma2_rn(p[i], t[i], f2(GCT_AS_A2));
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = __ffma2_rn(p[i], t[i], f2(GCT_AS_A1));
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = __fmul2_rn(t[i], e[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = __ffma2_rn(__fmul2_rn(p[i], t[i]), f2(-1.f), f2(0.5f));      // 0.5 - Phi(-|x|)
#pragma unroll
    for (int i = 0; i < NP; ++i) { p[i].x = copysignf(p[i].x, x[i].x); p[i].y = copysignf(p[i].y, x[i].y); }
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = __fadd2_rn(p[i], f2(0.5f));                                  // Phi(x)
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = __fmul2_rn(x[i], m[i]);                                      // keep * x
#pragma unroll
    for (int i = 0; i < NP; ++i) g[i] = __fmul2_rn(t[i], p[i]);
    if (WITH_GRAD) {
#pragma unroll
        for (int i = 0; i < NP; ++i) p[i] = __fmul2_rn(p[i], m[i]);
#pragma unroll
        for (int i = 0; i < NP; ++i) dg[i] = __ffma2_rn(__fmul2_rn(t[i], e[i]), f2(GCT_INV_SQRT_2PI), p[i]);
    }
}

// ------------------------------------------------------------------------------------------
// counter-based dropout RNG: keep(seed, site, idx) is a pure function, so backward regenerates
// the same mask without storing it.  (Parity with torch's Philox stream is statistical only.)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// dropout decision bits for element pair `x = seed + pair_index * 0x9e3779b9`: one multiply-xorshift round on top of the
// multiplicative counter (the 16-bit halves are compared against the threshold; tests/test_cpu_host.py checks keep rate
// and lag / cross-half correlations of a NumPy mirror).  Five instructions cheaper per pair than mix32, which matters in
// the GELU / attention epilogues where the hash was a third of the per-element work.
__host__ __device__ __forceinline__ uint32_t drop_bits(uint32_t x) {
    x ^= x >> 15; x *= 0x2c1b3c6dU;
    return x;
}
struct DropCtx {
    uint32_t seed;       // per-step seed
    uint32_t thresh;     // drop iff hash < thresh  (thresh = p * 2^32); 0 disables dropout
    float scale;         // 1/(1-p)
};
__host__ __device__ __forceinline__ DropCtx drop_site(DropCtx c, uint32_t site) {
    c.seed = mix32(c.seed ^ (0x9e3779b9U * (site + 1)));
    return c;
}
// one 32-bit hash serves the element pair (idx, idx^1): 16 bits each against thresh>>16
__device__ __forceinline__ float drop_apply(const DropCtx& c, uint64_t idx, float v) {
    if (c.thresh == 0) return v;
    const uint64_t pair = idx >> 1;
    const uint32_t h = drop_bits(c.seed + (uint32_t)pair * 0x9e3779b9U + (uint32_t)(pair >> 32) * 0x85ebca6bU);
    const uint32_t bits = (idx & 1) ? (h >> 16) : (h & 0xffffU);
    return (bits < (c.thresh >> 16)) ? 0.f : v * c.scale;
}

// both elements of the pair (2*pair_idx, 2*pair_idx+1) from one hash; identical to drop_apply for indices < 2^33
__device__ __forceinline__ void drop_pair(const DropCtx& c, uint32_t pair_idx, float& a, float& b) {
    const uint32_t h = drop_bits(c.seed + pair_idx * 0x9e3779b9U);
    const uint32_t t16 = c.thresh >> 16;
    a = ((h & 0xffffU) < t16) ? 0.f : a * c.scale;
    b = ((h >> 16) < t16) ? 0.f : b * c.scale;
}
// the same decisions as multipliers {0, 1/(1-p)} (1 when dropout is off): lets the caller fold the mask into packed math
template <bool BRANCH = true>
__device__ __forceinline__ float2 drop_mult_pair(const DropCtx& c, uint32_t pair_idx) {
    if (BRANCH && c.thresh == 0) return make_float2(1.f, 1.f);      // without the branch: thresh 0 keeps everything, scale is 1
    const uint32_t h = drop_bits(c.seed + pair_idx * 0x9e3779b9U);
    const uint32_t thi = c.thresh & 0xffff0000U;          // (h >> 16) < t16  <=>  h < (t16 << 16)
    return make_float2(((h << 16) < thi) ? 0.f : c.scale, (h < thi) ? 0.f : c.scale);
}
__device__ __forceinline__ float drop_mult(const DropCtx& c, uint64_t idx) {
    if (c.thresh == 0) return 1.f;
    const uint64_t pair = idx >> 1;
    const uint32_t h = drop_bits(c.seed + (uint32_t)pair * 0x9e3779b9U + (uint32_t)(pair >> 32) * 0x85ebca6bU);
    const uint32_t bits = (idx & 1) ? (h >> 16) : (h & 0xffffU);
    return (bits < (c.thresh >> 16)) ? 0.f : c.scale;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with the attribute may start (and run its
// prologue) while its predecessor drains; pdl_wait() blocks until the predecessor grid has fully
// completed and its writes are visible.  Both are no-ops for ordinary launches.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_gct_pdl;

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                   Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && g_gct_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
