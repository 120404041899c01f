// Decode self-attention over the KV cache, tensor-core form (bf16 tier; replaces the FMA / convert loop of decode_attn_kernel for
// the self-attention calls of a decode step -- Inference/sampling_tool.py:150-160 re-runs Model/sublayers.py:8-30 `attention` on the
// whole prefix every step; here one query row per (batch row, head) meets the cached keys).
//
// Why: ncu of decode_attn_kernel<bf16, 8, 3, 1> at 30 000 rows x 49 keys (profiles/r02_decode_attn_b30000_t49_v1_ncu_details.txt)
// shows 6.28 TB/s with the issue slots 78 % busy at 1.89 GHz: every cached element costs a bf16 -> fp32 convert and an FMA, so
// inside a power-capped decode (SM clock ~1.75 GHz) the kernel is as much issue-bound as HBM-bound.  Here the two contractions of a
// 16-key chunk are eight mma.sync.m16n8k16 per head, fed by ldmatrix from a swizzled tile -- about a third of the instructions.
//
// Data movement: one CTA per batch row, producer warp + one consumer warp per head as before, but a chunk arrives as ONE tiled
// TMA box per tensor: the cache [B][Lmax][H*64] is described as a rank-4 tensor (64 dims, key, head, row) -- the head dimension
// AFTER the key dimension -- so a (64, 16, H, 1) box lands as [head][key][128 B] with the 16-byte pieces of every 128-byte row
// XOR-swizzled by (key & 7) (SWIZZLE_128B): the 8 x 8 matrices ldmatrix fetches are conflict-free, and the DRAM side still sees
// whole 1 KB key rows.  The key extent of the tensor map is the number of cached keys of THIS step, so the box is clipped there:
// keys past it are neither fetched nor left to chance (zero fill; a stale NaN in the cache would survive p = 0 in the P V MMA).
//     S^T[key][n]  = K[key][dim] . q[dim]            A = K tile (ldmatrix), B = the query replicated over the 8 columns
//     O^T[dim][n] += V^T[dim][key] . p[key]          A = V tile (ldmatrix.trans), B = the probabilities replicated likewise
// All eight columns of an accumulator are equal, which is what makes the softmax shuffle-free: lane (g, t) reads the scores of keys
// g and g + 8 straight from its own registers.  This step's own key / value never touch shared memory (registers, as before).
#pragma once

namespace damma {

constexpr int CH = 16;                       // keys per chunk = the M of one MMA
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

// NS ring stages of (K box | V box), each box H x 16 keys x 128 B.  per_head = 1: the maps are rank 3 (dims, key, row) and a chunk
// is H boxes of one head each (same shared-memory image); used when the driver refuses the permuted strides of the rank-4 map.
// Three CTAs per SM (NS = 2: 3 x 70 KB of shared memory) need <= 75 registers per thread: without the bound ptxas took 93 for
// the version with the early producer start and the kernel dropped to two CTAs per SM (ncu: 5.64 instead of 6.6 TB/s).
template <int NS>
__global__ void __launch_bounds__(288, NS <= 2 ? 3 : (NS == 3 ? 2 : 1))
decode_attn_mma_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, DecAttnParams p, int B, int per_head) {
    extern __shared__ uint8_t dam_smem[];
    __shared__ uint8_t valid_s[DEC_MAX_KEYS];
    __shared__ __align__(8) uint64_t bars[2 * NS];
    __shared__ __align__(16) float osc[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = p.H;
    const uint32_t ring = (tc::smem_u32(dam_smem) + 1023u) & ~1023u;
    const uint32_t box_bytes = (uint32_t)H * CH * 128u, stage_bytes = 2u * box_bytes;
    const uint32_t bar0 = tc::smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { tc::mbar_init(bar0 + 8 * s, 1); tc::mbar_init(bar0 + 8 * (NS + s), H); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    }
    __syncthreads();
    // Programmatic dependent launch: the kernel before this one (the QKV projection) produces only q and this step's K / V row.
    // The cached keys, the validity bytes, `done` and the row map were written at least two kernels earlier, and a kernel is
    // launched only after its predecessor has passed ITS dependency wait, i.e. after everything older has completed -- so the
    // producer warp streams the cache without waiting, and the boxes of a CTA's first stages arrive while the projection drains.
    const int b = blockIdx.x;
    const int bp = p.rowmap ? p.rowmap[b] : b;
    if (p.done != nullptr && p.done[bp] != 0) {                // uniform over the CTA
        pdl_wait();                                            // keeps "launched => everything older has completed" true for the successors
        pdl_launch_dependents();
        return;
    }
    const int nc = p.n_cached;
    const int nchunks = (nc + CH - 1) / CH;

    if (warp == H) {
        // ---------------- producer (no dependency wait, see above; it does not release the dependents either) ----------------
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % NS;
                tc::mbar_wait(bar0 + 8 * (NS + s), ((c / NS) & 1) ^ 1);
                const uint32_t dst = ring + s * stage_bytes;
                tc::mbar_expect_tx(bar0 + 8 * s, stage_bytes);
                if (!per_head) {
                    tma_load_4d(dst, &tmK, 0, c * CH, 0, bp, bar0 + 8 * s);
                    tma_load_4d(dst + box_bytes, &tmV, 0, c * CH, 0, bp, bar0 + 8 * s);
                } else {
                    for (int h = 0; h < H; ++h) {
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                     ::"r"(dst + h * (CH * 128)), "l"(reinterpret_cast<uint64_t>(&tmK)), "r"(h * 64), "r"(c * CH), "r"(bp), "r"(bar0 + 8 * s) : "memory");
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                     ::"r"(dst + box_bytes + h * (CH * 128)), "l"(reinterpret_cast<uint64_t>(&tmV)), "r"(h * 64), "r"(c * CH), "r"(bp), "r"(bar0 + 8 * s) : "memory");
                    }
                }
            }
        }
        return;
    }
    // ---------------- consumers: warp = head ----------------
    const int h = warp, g = lane >> 2, t = lane & 3, sub = lane & 7;
    const uint8_t* valid = p.key_valid + (size_t)bp * p.kv_stride;
    // validity bytes of the cached keys and of this step's key: every consumer warp copies its share; they are read only by the
    // warp's own lanes for keys 16c + g (+8), so a warp-wide copy + __syncwarp would do -- but the slots are shared, hence the
    // named barrier over the consumer warps below
    for (int j = threadIdx.x; j <= nc && j < DEC_MAX_KEYS; j += H * 32) valid_s[j] = valid[j];
    pdl_wait();
    pdl_launch_dependents();
    const bf16* qrow = reinterpret_cast<const bf16*>(p.q) + (size_t)b * p.ldq + h * 64;
    uint32_t qf[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        qf[ks][0] = *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * t);
        qf[ks][1] = *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * t + 8);
    }
    const f8 q8 = ld8(qrow + sub * 8);
    f8 kn, vn;
    const bool has_new = p.knew != nullptr;
    if (has_new) {
        const int col = h * 64 + sub * 8;
        kn = ld8(reinterpret_cast<const bf16*>(p.knew) + (size_t)b * p.ldnew + col);
        vn = ld8(reinterpret_cast<const bf16*>(p.vnew) + (size_t)b * p.ldnew + col);
        if (lane < 8) {          // append this step's K / V row to the cache
            st8(reinterpret_cast<bf16*>(p.kcache) + (size_t)bp * p.cache_bstride + (size_t)nc * p.pitch + col, kn);
            st8(reinterpret_cast<bf16*>(p.vcache) + (size_t)bp * p.cache_bstride + (size_t)nc * p.pitch + col, vn);
        }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(H * 32) : "memory");          // valid_s complete (consumer warps only)

    const float sl2 = p.scale * kLog2e, masked = -1e9f * kLog2e;
    float m = -INFINITY, l = 0.f;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    // ldmatrix row addresses of this lane inside a head's [16 keys][128 B] tile (16-byte piece index XOR key & 7)
    const int kkey = ((lane >> 3) & 1) * 8 + (lane & 7), kpart = lane >> 4;          // K: matrices (keys lo|hi) x (dims lo|hi)
    const int vkey = ((lane >> 4) & 1) * 8 + (lane & 7), vpart = (lane >> 3) & 1;    // V^T: matrices (dims lo|hi) x (keys lo|hi)
    for (int c = 0; c < nchunks; ++c) {
        const int s = c % NS;
        tc::mbar_wait(bar0 + 8 * s, (c / NS) & 1);
        const uint32_t Kt = ring + s * stage_bytes + h * (CH * 128), Vt = Kt + box_bytes;
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(Kt + kkey * 128 + (((2 * ks + kpart) ^ (kkey & 7)) << 4), a0, a1, a2, a3);
            mma16816(sc, a0, a1, a2, a3, qf[ks][0], qf[ks][1]);
        }
        const int klo = c * CH + g, khi = klo + 8;
        const float slo = klo < nc ? (valid_s[klo] ? sc[0] * sl2 : masked) : -INFINITY;
        const float shi = khi < nc ? (valid_s[khi] ? sc[2] * sl2 : masked) : -INFINITY;
        float cm = fmaxf(slo, shi);
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 4));
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 8));
        cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 16));
        const float mn = fmaxf(m, cm);                      // finite: key 16c is in range
        const float corr = ex2_approx(m - mn);              // 0 on the first chunk
        const float plo = ex2_approx(slo - mn), phi = ex2_approx(shi - mn);
        l = l * corr + plo + phi;
        if (corr != 1.f) {                                  // uniform over the warp
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[i][0] *= corr; acc[i][1] *= corr; acc[i][2] *= corr; acc[i][3] *= corr; }
        }
        m = mn;
        // B fragment: rows (keys) 2t, 2t+1 and 2t+8, 2t+9, the same in every column
        const uint32_t pk = tc::pack_bf16x2(plo, phi);                      // (p[g], p[g+8])
        const uint32_t X = __shfl_sync(0xffffffffu, pk, 8 * t), Y = __shfl_sync(0xffffffffu, pk, 8 * t + 4);
        const uint32_t pb0 = __byte_perm(X, Y, 0x5410), pb1 = __byte_perm(X, Y, 0x7632);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4_t(Vt + vkey * 128 + (((2 * mt + vpart) ^ (vkey & 7)) << 4), a0, a1, a2, a3);
            mma16816(acc[mt], a0, a1, a2, a3, pb0, pb1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (NS + s));
    }
    // l holds this lane's keys g, g + 8 of every chunk: total over g (lanes that differ only in t hold copies)
    l += __shfl_xor_sync(0xffffffffu, l, 4);
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
    float corr = 1.f, pn = 0.f;
    if (has_new) {
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d = fmaf(q8.v[e], kn.v[e], d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        const float sn = valid_s[nc] ? d * sl2 : masked;
        const float mn = fmaxf(m, sn);
        corr = ex2_approx(m - mn);
        pn = ex2_approx(sn - mn);
        l = l * corr + pn;
    }
    const float inv = 1.f / l;
    if (t == 0) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) { osc[h][16 * mt + g] = acc[mt][0]; osc[h][16 * mt + g + 8] = acc[mt][2]; }
    }
    __syncwarp();
    if (lane < 8) {
        f8 o;
        const float4 x0 = *reinterpret_cast<const float4*>(&osc[h][sub * 8]), x1 = *reinterpret_cast<const float4*>(&osc[h][sub * 8 + 4]);
        const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = (xs[e] * corr + (has_new ? pn * vn.v[e] : 0.f)) * inv;
        st8(reinterpret_cast<bf16*>(p.out) + (size_t)b * p.ldo + h * 64 + sub * 8, o);
    }
}

// (dims 64 | key | head | row) view of a [rows][Lmax][H*64] bf16 cache clipped to `keys` keys; falls back to (H*64 | key | row).
struct CacheMapKey {
    const void* p; uint64_t keys, rows, bstride; int H, rank;
    bool operator==(const CacheMapKey& o) const { return p == o.p && keys == o.keys && rows == o.rows && bstride == o.bstride && H == o.H && rank == o.rank; }
};
struct CacheMapHash {
    size_t operator()(const CacheMapKey& k) const {
        size_t h = (size_t)k.p; h = h * 1000003u ^ k.keys; h = h * 1000003u ^ k.rows; h = h * 1000003u ^ k.bstride; h = h * 1000003u ^ (size_t)(k.H * 8 + k.rank);
        return h;
    }
};
static int g_rank4_ok = -1;       // -1 unknown, 1 the driver takes the permuted rank-4 strides, 0 it does not

static int cache_tensor_map(const void* ptr, int keys, long long rows, long long bstride_elems, int H, int rank, CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<CacheMapKey, CUtensorMap, CacheMapHash> cache;
    static tc::PFN_encodeTiled encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GCT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) GCT_FAIL(GCT_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<tc::PFN_encodeTiled>(fn);
    }
    CacheMapKey key{ptr, (uint64_t)keys, (uint64_t)rows, (uint64_t)bstride_elems, H, rank};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return GCT_OK; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((bstride_elems * 2) & 15))
        GCT_FAIL(GCT_ERR_ARG, "decode attention: the cache must be 16-byte aligned (ptr=%p row stride %lld elements)", ptr, bstride_elems);
    CUtensorMap m;
    CUresult r;
    if (rank == 4) {
        cuuint64_t dims[4] = {64, (cuuint64_t)keys, (cuuint64_t)H, (cuuint64_t)rows};
        cuuint64_t strides[3] = {(cuuint64_t)H * 128, 128, (cuuint64_t)bstride_elems * 2};
        cuuint32_t box[4] = {64, CH, (cuuint32_t)H, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[3] = {(cuuint64_t)H * 64, (cuuint64_t)keys, (cuuint64_t)rows};
        cuuint64_t strides[2] = {(cuuint64_t)H * 128, (cuuint64_t)bstride_elems * 2};
        cuuint32_t box[3] = {64, CH, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        if (rank == 4) return GCT_ERR_UNSUPPORTED;          // caller retries with rank 3 (no message: not an error yet)
        GCT_FAIL(GCT_ERR_CUDA, "cuTensorMapEncodeTiled (cache, rank %d) failed: %d", rank, (int)r);
    }
    if (cache.size() > 16384) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return GCT_OK;
}

template <int NS>
static int launch_cfg(const DecAttnParams& p, int B, long long rows, int force_per_head, cudaStream_t st) {
    CUtensorMap tk, tv;
    const int keys = p.n_cached > 0 ? p.n_cached : 1;        // step 0: no box is issued, any valid map will do
    int per_head = force_per_head || g_rank4_ok == 0;
    if (!per_head) {
        int r = cache_tensor_map(p.kcache, keys, rows, p.cache_bstride, p.H, 4, &tk);
        if (r == GCT_OK) r = cache_tensor_map(p.vcache, keys, rows, p.cache_bstride, p.H, 4, &tv);
        if (r == GCT_ERR_UNSUPPORTED) { g_rank4_ok = 0; per_head = 1; }
        else if (r != GCT_OK) return r;
        else g_rank4_ok = 1;
    }
    if (per_head) {
        GCT_TRY(cache_tensor_map(p.kcache, keys, rows, p.cache_bstride, p.H, 3, &tk));
        GCT_TRY(cache_tensor_map(p.vcache, keys, rows, p.cache_bstride, p.H, 3, &tv));
    }
    const size_t smem = (size_t)NS * 2 * p.H * CH * 128 + 1024;
    auto kern = decode_attn_mma_kernel<NS>;
    GCT_SMEM_LIMIT(kern, smem);
    GCT_CUDA(launch_k(kern, dim3(B), dim3((p.H + 1) * 32), smem, st, true, tk, tv, p, B, per_head));
    return GCT_OK;
}

}  // namespace damma
