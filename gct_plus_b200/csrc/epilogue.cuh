// GEMM epilogue shared by the SIMT (fp32 parity tier) and tcgen05 (bf16) kernels.
//   v = acc (+bias[col]);  [GELU: save pre-activation or keep*gelu', v = gelu(v)]  [DGELU: v = drop'(v)*gelu'(aux)]  [MUL_AUX]
//   v = dropout(v);  v += residual;  store fp32 and/or T, or accumulate (wgrad).
#pragma once
#include "common.cuh"

enum : int {
    EPI_GELU = 1,        // v = gelu(v); pre-activation stored to aux_out if non-null
    EPI_DGELU = 2,       // v = v * gelu'(aux_in)
    EPI_ACCUM = 4,       // out32 += v (atomic when split_k > 1)
    EPI_BIAS_ROW = 8,    // bias indexed by row instead of column (unused by the model; kept for tests)
    EPI_NOSTORE = 16,    // measurement hook: run the epilogue math but skip the global stores
    // 32 / 64 are measurement hooks of the persistent kernel
    EPI_GELU_GRAD = 128, // with EPI_GELU: aux_out receives keep*gelu'(pre) (the factor that turns the gradient of the
                         // dropped-out activation into the gradient of the pre-activation) instead of the pre-activation
    EPI_MUL_AUX = 256,   // v *= aux_in  (backward of an EPI_GELU|EPI_GELU_GRAD forward: no erf, no mask regeneration)
};

struct Epilogue {
    const float* bias;      // [N] or null
    const float* res32;     // [M, ldc] fp32 addend (residual / incoming gradient) or null
    const void* aux_in;     // [M, ldc] T: pre-activation for DGELU
    void* aux_out;          // [M, ldc] T: pre-activation saved by GELU
    float* out32;           // fp32 output or null
    void* outT;             // T output or null
    int ldc;                // leading dimension of every [M, .] tensor above
    int flags;
    float alpha;            // scales acc before anything else
    DropCtx drop;           // dropout applied to the value (index = row*ldc+col)
};

template <typename T>
__device__ __forceinline__ void epilogue_apply(const Epilogue& e, int row, int col, float acc, bool atomic) {
    const size_t idx = (size_t)row * e.ldc + col;
    float v = acc * e.alpha;
    if (e.bias) v += e.bias[(e.flags & EPI_BIAS_ROW) ? row : col];
    const float keep = drop_mult(e.drop, idx);
    if (e.flags & EPI_GELU) {
        if (e.aux_out) reinterpret_cast<T*>(e.aux_out)[idx] = from_f<T>((e.flags & EPI_GELU_GRAD) ? keep * gelu_erf_grad(v) : v);
        v = gelu_erf(v);
    }
    v *= keep;
    if (e.flags & EPI_DGELU) v *= gelu_erf_grad(to_f(reinterpret_cast<const T*>(e.aux_in)[idx]));
    if (e.flags & EPI_MUL_AUX) v *= to_f(reinterpret_cast<const T*>(e.aux_in)[idx]);
    if (e.res32) v += e.res32[idx];
    if (e.flags & EPI_ACCUM) {
        if (atomic) atomicAdd(e.out32 + idx, v);
        else e.out32[idx] += v;
        return;
    }
    if (e.out32) e.out32[idx] = v;
    if (e.outT) reinterpret_cast<T*>(e.outT)[idx] = from_f<T>(v);
}
