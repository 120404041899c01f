// fp32-accumulate SIMT GEMM: the parity tier (1e-4 vs the fp32 reference, bit-identical greedy
// tokens) and the cross-check for the tcgen05 kernel.  C[M,N] = sum_k A(m,k)*B(n,k) with arbitrary
// element strides, so the same kernel serves fprop (A K-contig, B K-contig), dgrad (B N-contig)
// and wgrad (both M-contig).  64x64x16 tile, 256 threads, 4x4 micro-tile, optional split-K.
#pragma once
#include "epilogue.cuh"

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, long long sam, long long sak, const TB* __restrict__ B, long long sbn,
                 long long sbk, int M, int N, int K, int k_per_split, Epilogue epi) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const bool a_kc = (sak == 1), b_kc = (sbk == 1);
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;
            int mm, kk;
            if (a_kc) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < kend) ? to_f(A[(long long)gm * sam + (long long)gk * sak]) : 0.f;
            int nn, kb;
            if (b_kc) { kb = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kb = idx >> 6; }
            const int gn = n0 + nn, gkb = k0 + kb;
            Bs[kb][nn] = (gn < N && gkb < kend) ? to_f(B[(long long)gn * sbn + (long long)gkb * sbk]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool atomic = gridDim.z > 1;
    if (blockIdx.z != 0) epi.bias = nullptr;      // split-K: the bias is added once
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + ty * 4 + i;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c < N) epilogue_apply<TC>(epi, r, c, acc[i][j], atomic);
        }
    }
}

template <typename TA, typename TB, typename TC>
static int launch_gemm_simt(const TA* A, long long sam, long long sak, const TB* B, long long sbn, long long sbk,
                            int M, int N, int K, int split_k, const Epilogue& epi, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return GCT_OK;
    if (split_k < 1) split_k = 1;
    if (split_k > 1 && !(epi.flags & EPI_ACCUM)) GCT_FAIL(GCT_ERR_ARG, "split-K needs an accumulating epilogue");
    int kps = ((K + split_k - 1) / split_k + 15) / 16 * 16;
    split_k = (K + kps - 1) / kps;
    dim3 grid(cdiv(N, 64), cdiv(M, 64), split_k);
    gemm_simt_kernel<TA, TB, TC><<<grid, 256, 0, st>>>(A, sam, sak, B, sbn, sbk, M, N, K, kps, epi);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
