// C-ABI entry points of libgct_b200.so (declared in include/gct_b200.h).
#include "model.cuh"
#include <dlfcn.h>

thread_local char g_gct_err[512] = {0};
int g_gct_simt_only = 0;
int g_gct_pdl = 1;
int g_da_cfg = 0;
int g_gct_simt_attn = 0;
int g_gct_zattn = 1;
int g_gct_ffn_classic = 0;
int g_gct_persist = 1;
int g_gct_tma_store = 1;
int g_gct_ew4 = 1;
int g_gct_pair = 2;
int g_gct_attn_bias_separate = 0;
int g_za_cfg = 0;
int g_gct_sm_budget = 0;
int g_gct_res_box = 1;
int g_gct_attn_box = 1;
int g_gct_attn_persist = 2;      // backward only: the persistent forward is no faster than four one-tile CTAs per SM (profiles/r02_ab_train_attention.txt)
unsigned long long* g_gct_attn_trace = nullptr;
int g_gct_rownorm = 0;
int g_gct_rownorm_res_tma = 1;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* gct_last_error(void) { return g_gct_err; }
int gct_version(void) { return 100; }
int gct_sm(void) {
#ifdef GCT_SM_TARGET
    return GCT_SM_TARGET;
#else
    return 0;
#endif
}
int gct_num_slots(int n_layers) { return GCT_NUM_GLOBAL_SLOTS + n_layers * (GCT_ENC_LAYER_SLOTS + GCT_DEC_LAYER_SLOTS); }
int gct_set_gemm_backend(int simt_only) { g_gct_simt_only = simt_only; return GCT_OK; }
int gct_set_pdl(int enabled) { g_gct_pdl = enabled; return GCT_OK; }
int gct_set_decode_attn_config(int cfg) { g_da_cfg = cfg; return GCT_OK; }
int gct_set_attention_backend(int simt_only) { g_gct_simt_attn = simt_only; return GCT_OK; }
int gct_set_latent_cross_attention(int enabled) { g_gct_zattn = enabled; return GCT_OK; }
int gct_set_ffn_saved_activation(int preact) { g_gct_ffn_classic = preact; return GCT_OK; }
int gct_set_persistent_gemm(int enabled) { g_gct_persist = enabled; return GCT_OK; }
int gct_set_tma_store(int enabled) { g_gct_tma_store = enabled; return GCT_OK; }
int gct_set_epilogue_warps16(int enabled) { g_gct_ew4 = enabled; return GCT_OK; }
int gct_set_cta_pair_gemm(int enabled) { g_gct_pair = enabled; return GCT_OK; }
int gct_set_rownorm_fusion(int mode) { g_gct_rownorm = mode & 3; g_gct_rownorm_res_tma = (mode & 4) ? 0 : 1; return GCT_OK; }
int gct_set_attention_persistent(int mode) { g_gct_attn_persist = mode & 3; return GCT_OK; }
int gct_set_attention_trace(void* dev_buf) { g_gct_attn_trace = static_cast<unsigned long long*>(dev_buf); return GCT_OK; }
int gct_set_residual_box(int enabled) { g_gct_res_box = enabled & 1; g_gct_attn_box = (enabled & 2) ? 0 : 1; return GCT_OK; }
int gct_set_sm_budget(int sms) { g_gct_sm_budget = sms; return GCT_OK; }
int gct_set_zattn_config(int ctas_per_sm) { g_za_cfg = ctas_per_sm; return GCT_OK; }
int gct_set_attention_bias_grad_fused(int enabled) { g_gct_attn_bias_separate = !enabled; return GCT_OK; }

int gct_norm_fwd(const float* x, const float* alpha, const float* bias, void* y, float* y32, int rows, int d, int dtype,
                 void* stream) {
    GCT_REQUIRE(d % 128 == 0 && d <= 1024, "norm: d=%d must be a multiple of 128, <= 1024", d);
    if (rows <= 0) return GCT_OK;
    dim3 grid(cdiv(rows, 8));
    const int nv = d / 128;
#define CASE(NV)                                                                                                       \
    case NV:                                                                                                           \
        if (dtype == GCT_DTYPE_F32) norm_fwd_kernel<float, NV><<<grid, 256, 0, ST(stream)>>>(x, alpha, bias, (float*)y, y32, rows, 1e-6f); \
        else norm_fwd_kernel<bf16, NV><<<grid, 256, 0, ST(stream)>>>(x, alpha, bias, (bf16*)y, y32, rows, 1e-6f);      \
        break;
    switch (nv) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) }
#undef CASE
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

int gct_norm_bwd(const float* x, const float* alpha, const float* dy, const float* add, float* dx, float* dalpha,
                 float* dbias, int rows, int d, void* stream) {
    GCT_REQUIRE(d % 128 == 0 && d <= 1024, "norm: d=%d must be a multiple of 128, <= 1024", d);
    if (rows <= 0) return GCT_OK;
    dim3 grid(min(cdiv(rows, 8), 148 * 3));
    const size_t sm = (size_t)8 * 3 * d * sizeof(float);
    DropCtx nodrop; nodrop.seed = 0; nodrop.thresh = 0; nodrop.scale = 1.f;
#define CASE(NV) case NV: if (sm > 48 * 1024) GCT_SMEM_LIMIT((norm_bwd_kernel<float, NV>), sm); norm_bwd_kernel<float, NV><<<grid, 256, sm, ST(stream)>>>(x, alpha, dy, add, dx, dalpha, dbias, rows, 1e-6f, (float*)nullptr, nodrop, (float*)nullptr); break;
    switch (d / 128) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) }
#undef CASE
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

int gct_gemm(const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N, int K,
             const float* bias, const float* res32, const void* aux_in, void* aux_out, float* out32, void* outT, int ldc,
             int flags, int split_k, int bn_hint, int dtype, void* stream) {
    Epilogue e;
    memset(&e, 0, sizeof(e));
    e.bias = bias; e.res32 = res32; e.aux_in = aux_in; e.aux_out = aux_out; e.out32 = out32; e.outT = outT; e.ldc = ldc;
    e.flags = flags; e.alpha = 1.f; e.drop.thresh = 0; e.drop.scale = 1.f;
    GCT_REQUIRE(out32 || outT, "gemm: no output");
    if (dtype == GCT_DTYPE_F32) {
        return launch_gemm_simt<float, float, float>((const float*)A, a_mn ? 1 : lda, a_mn ? lda : 1, (const float*)B,
                                                     b_mn ? 1 : ldb, b_mn ? ldb : 1, M, N, K, split_k, e, ST(stream));
    }
    if (g_gct_simt_only)
        return launch_gemm_simt<bf16, bf16, bf16>((const bf16*)A, a_mn ? 1 : lda, a_mn ? lda : 1, (const bf16*)B, b_mn ? 1 : ldb,
                                                  b_mn ? ldb : 1, M, N, K, split_k, e, ST(stream));
    return tc::launch_gemm_tc((const bf16*)A, a_mn != 0, lda, (const bf16*)B, b_mn != 0, ldb, M, N, K, split_k, bn_hint, e,
                              ST(stream));
}

int gct_gemm_rownorm(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, const float* bias, const float* res32,
                     float* out32, const float* alpha, const float* beta, void* normT, float* norm32, float eps, void* stream) {
    GCT_REQUIRE(A && W && alpha && beta && normT, "gemm_rownorm: null argument");
    tc::RowNormParams rp;
    rp.bias = bias; rp.res32 = res32; rp.out32 = out32; rp.alpha = alpha; rp.beta = beta; rp.norm32 = norm32;
    rp.drop.seed = 0; rp.drop.thresh = 0; rp.drop.scale = 1.f; rp.eps = eps; rp.M = M; rp.K = K;
    return tc::launch_gemm_rownorm((const bf16*)A, lda, (const bf16*)W, ldw, (bf16*)normT, rp, ST(stream));
}

static AttnParams make_attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* mask,
                            int64_t mb, int mr, void* out, int ldo, float* lse, float* probs, int B, int H, int Lq, int Lk) {
    AttnParams p;
    p.Q = q; p.K = k; p.V = v; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.mask = mask; p.mask_bstride = mb; p.mask_rstride = mr;
    p.O = out; p.ldo = ldo; p.lse = lse; p.probs = probs; p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.scale = 0.125f;
    p.drop.seed = 0; p.drop.thresh = 0; p.drop.scale = 1.f;
    return p;
}

int gct_attention_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* mask,
                      int64_t mask_bstride, int mask_rstride, void* out, int ldo, float* lse, float* probs, int B, int H,
                      int Lq, int Lk, int dtype, void* stream) {
    AttnParams p = make_attn(q, ldq, k, ldk, v, ldv, mask, mask_bstride, mask_rstride, out, ldo, lse, probs, B, H, Lq, Lk);
    return dtype == GCT_DTYPE_F32 ? attn_fwd_dispatch<float>(p, ST(stream)) : attn_fwd_dispatch<bf16>(p, ST(stream));
}

int gct_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* mask,
                      int64_t mask_bstride, int mask_rstride, const float* lse, const void* out, int ldo, const void* dout, int lddo,
                      void* dq, int lddq,
                      void* dk, int lddk, void* dv, int lddv, int B, int H, int Lq, int Lk, int dtype, void* stream) {
    AttnBwdParams bp;
    bp.f = make_attn(q, ldq, k, ldk, v, ldv, mask, mask_bstride, mask_rstride, const_cast<void*>(out), ldo, const_cast<float*>(lse),
                     nullptr, B, H, Lq, Lk);
    bp.dO = dout; bp.lddo = lddo; bp.dQ = dq; bp.dK = dk; bp.dV = dv; bp.lddq = lddq; bp.lddk = lddk; bp.lddv = lddv;
    return dtype == GCT_DTYPE_F32 ? attn_bwd_dispatch<float>(bp, ST(stream)) : attn_bwd_dispatch<bf16>(bp, ST(stream));
}

int gct_src_mask(const int64_t* tok, int B, int L, int nc, int pad, uint8_t* out, void* stream) {
    const int n = B * (nc + L);
    if (n <= 0) return GCT_OK;
    GCT_CUDA(launch_k(src_mask_kernel, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), ST(stream), true, tok, B, L, nc, pad, out));
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
int gct_trg_mask(const int64_t* tok, int B, int T, int nc, int pad, uint8_t* out, void* stream) {
    const size_t n = (size_t)B * (nc + T) * (nc + T);
    if (n == 0) return GCT_OK;
    GCT_CUDA(launch_k(trg_mask_kernel, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), ST(stream), true, tok, B, T, nc, pad, out));
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
int gct_mask_cast(const void* in, int elem_size, int64_t n, uint8_t* out, void* stream) {
    GCT_REQUIRE(elem_size == 1 || elem_size == 4 || elem_size == 8, "mask_cast: element size %d", elem_size);
    if (n <= 0) return GCT_OK;
    mask_cast_kernel<<<cdiv(n, 256), 256, 0, ST(stream)>>>(in, elem_size, (size_t)n, out);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
int gct_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
    if (n <= 0) return GCT_OK;
    cast_kernel<float, bf16><<<cdiv(n, 1024), 256, 0, ST(stream)>>>(in, (bf16*)out, (size_t)n);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
int gct_cast_bf16_to_f32(const void* in, float* out, int64_t n, void* stream) {
    if (n <= 0) return GCT_OK;
    cast_kernel<bf16, float><<<cdiv(n, 1024), 256, 0, ST(stream)>>>((const bf16*)in, out, (size_t)n);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

// ---------------- loss ----------------
static const int KL_BLOCKS = 592;
size_t gct_loss_scratch_bytes(int64_t rows, int64_t n_latent) { (void)n_latent; return (size_t)(rows + KL_BLOCKS + 64) * sizeof(float); }

__global__ void loss_combine_kernel(const float* rce, const float* kld, float beta, float* out4) {
    pdl_wait();                  // programmatic dependent launch: everything below may read / write what earlier kernels touch
    pdl_launch_dependents();
    out4[0] = rce[0] + beta * kld[0]; out4[1] = rce[0]; out4[2] = 0.f; out4[3] = kld[0];
}

int gct_loss_fwd_bwd(const float* logits, int ld, int V, const int64_t* target, int64_t rows, int pad_id, const float* mu,
                     const float* log_var, int64_t n_latent, float beta, float gscale, float* out4, float* dlogits, float* dmu,
                     float* dlv, void* scratch, void* stream) {
    GCT_REQUIRE(V <= 128, "loss: vocabulary %d > 128", V);
    GCT_REQUIRE(scratch && out4, "loss: scratch / out missing");
    float* row_loss = reinterpret_cast<float*>(scratch);
    float* part = row_loss + rows;
    float* sums = part + KL_BLOCKS;       // [0] = RCE, [1] = KLD
    cudaStream_t st = ST(stream);
    GCT_CUDA(launch_k(ce_rows_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)(0), st, true, logits, target, (int)rows, V, ld, pad_id, row_loss, dlogits, gscale));
    GCT_LAUNCH_CHECK();
    GCT_CUDA(launch_k(final_sum_kernel, dim3(1), dim3(1024), (size_t)(0), st, true, row_loss, (size_t)rows, sums));
    GCT_LAUNCH_CHECK();
    GCT_CUDA(launch_k(kl_partial_kernel, dim3(KL_BLOCKS), dim3(256), (size_t)(0), st, true, mu, log_var, (size_t)n_latent, part));
    GCT_LAUNCH_CHECK();
    GCT_CUDA(launch_k(final_sum_kernel, dim3(1), dim3(1024), (size_t)(0), st, true, part, (size_t)KL_BLOCKS, sums + 1));
    GCT_LAUNCH_CHECK();
    GCT_CUDA(launch_k(loss_combine_kernel, dim3(1), dim3(1), (size_t)(0), st, true, sums, sums + 1, beta, out4));
    GCT_LAUNCH_CHECK();
    if (dmu && dlv) {
        GCT_CUDA(launch_k(kl_bwd_kernel, dim3(cdiv(n_latent, 256)), dim3(256), (size_t)(0), st, true, mu, log_var, (size_t)n_latent, beta * gscale, dmu, dlv));
        GCT_LAUNCH_CHECK();
    }
    return GCT_OK;
}

int gct_prop_head_fwd_bwd(const float* logits, int B, int Ld, int nc, int V, const float* w, const float* b0, const float* target,
                          float gscale, float* prop_out, float* out4, float* dlogits, float* dw, float* db, void* stream) {
    GCT_REQUIRE(logits && w && b0 && target && B >= 0 && nc >= 1 && nc <= Ld && V >= 1 && V <= 128, "prop_head: bad arguments");
    GCT_REQUIRE(!dlogits || (dw && db), "prop_head: gradient outputs missing");
    if (B == 0) return GCT_OK;
    GCT_CUDA(launch_k(prop_head_kernel, dim3(cdiv((long long)B * nc, 8)), dim3(256), (size_t)(0), ST(stream), true, logits, B, Ld, nc, V, w, b0, target, gscale, prop_out, out4, dlogits, dw, db));
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

// ---------------- model ----------------
}  // extern "C"
template <typename T>
static int forward_impl(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, void* ws, size_t ws_bytes,
                        void* stream) {
    Model<T> m;
    GCT_TRY(m.init(cfg, w, io->seed, io->train, stream));
    Acts<T> A;
    A.carve(*cfg, io->B, io->S, io->T, ws);
    GCT_REQUIRE(A.bytes <= ws_bytes, "forward: workspace too small (%zu < %zu)", ws_bytes, A.bytes);
    return model_forward<T>(m, *io, A);
}
template <typename T>
static int backward_impl(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, const float* dlogits,
                         const float* dmu, const float* dlv, const float* dz, void* ws, size_t ws_bytes, void* scratch,
                         size_t scratch_bytes, void* stream, StageHook* hook = nullptr) {
    Model<T> m;
    GCT_TRY(m.init(cfg, w, io->seed, io->train, stream));
    Acts<T> A;
    A.carve(*cfg, io->B, io->S, io->T, ws);
    GCT_REQUIRE(A.bytes <= ws_bytes, "backward: workspace too small");
    BwdScratch<T> S;
    S.carve(*cfg, A, scratch);
    GCT_REQUIRE(S.bytes <= scratch_bytes, "backward: scratch too small (%zu < %zu)", scratch_bytes, S.bytes);
    GCT_TRY(model_backward<T>(m, *io, A, S, dlogits, dmu, dlv, dz, hook));
    if (hook) GCT_TRY(hook->done(2 * cfg->n_layers));
    return GCT_OK;
}

extern "C" {
size_t gct_forward_workspace_bytes(const gct_config_t* cfg, int B, int S, int T) {
    if (cfg->dtype == GCT_DTYPE_F32) { Acts<float> A; A.carve(*cfg, B, S, T, nullptr); return A.bytes; }
    Acts<bf16> A; A.carve(*cfg, B, S, T, nullptr); return A.bytes;
}
size_t gct_backward_scratch_bytes(const gct_config_t* cfg, int B, int S, int T) {
    if (cfg->dtype == GCT_DTYPE_F32) { Acts<float> A; A.carve(*cfg, B, S, T, nullptr); BwdScratch<float> X; X.carve(*cfg, A, nullptr); return X.bytes; }
    Acts<bf16> A; A.carve(*cfg, B, S, T, nullptr); BwdScratch<bf16> X; X.carve(*cfg, A, nullptr); return X.bytes;
}
int gct_forward(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, void* workspace, size_t workspace_bytes,
                void* stream) {
    GCT_REQUIRE(cfg && w && io && workspace, "forward: null argument");
    return cfg->dtype == GCT_DTYPE_F32 ? forward_impl<float>(cfg, w, io, workspace, workspace_bytes, stream)
                                       : forward_impl<bf16>(cfg, w, io, workspace, workspace_bytes, stream);
}
int gct_backward(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, const float* dlogits, const float* dmu,
                 const float* dlog_var, const float* dz, void* workspace, size_t workspace_bytes, void* scratch,
                 size_t scratch_bytes, void* stream) {
    GCT_REQUIRE(cfg && w && io && workspace && scratch, "backward: null argument");
    return cfg->dtype == GCT_DTYPE_F32
               ? backward_impl<float>(cfg, w, io, dlogits, dmu, dlog_var, dz, workspace, workspace_bytes, scratch, scratch_bytes, stream)
               : backward_impl<bf16>(cfg, w, io, dlogits, dmu, dlog_var, dz, workspace, workspace_bytes, scratch, scratch_bytes, stream);
}

// ---------------- optimiser ----------------
int gct_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow, int64_t n, int step,
                  float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
    GCT_REQUIRE(step >= 1, "adam: step must start at 1");
    if (n <= 0) return GCT_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const int vec16 = ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                        reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(bf16_shadow) & 7) == 0;
    GCT_CUDA(launch_k(adam_kernel, dim3(cdiv(n, 1024)), dim3(256), (size_t)(0), ST(stream), true, params, grads, exp_avg, exp_avg_sq, (bf16*)bf16_shadow, (size_t)n, lr, beta1,
                                                      beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale, vec16));
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
double gct_noam_lr(int64_t step, int d_model, int64_t warmup) {
    const double head = pow((double)step, -0.5), tail = (double)step * pow((double)warmup, -1.5);
    return pow((double)d_model, -0.5) * (head < tail ? head : tail);
}

// Host-side detokeniser (Inference/sampling_tool.py:54-61 for a whole batch): each row of ids is cut at its first
// <eos>, <sos> ids are dropped, the token strings (UTF-8, concatenated in `vocab` with offsets `voff[V+1]`) are joined
// and the row is terminated by '\n'.  `out` must hold n*(width*max_token_bytes + 1) bytes; returns the bytes written
// or a negative code.  Plain C loop on the calling thread: 3 M tokens take a few milliseconds.
int64_t gct_detokenize(const int16_t* ids, int64_t n, int width, const char* vocab, const int32_t* voff, int V, int eos_id,
                       int sos_id, char* out, int64_t out_bytes) {
    if (!ids || !vocab || !voff || !out || n < 0 || width < 0 || V <= 0) {
        snprintf(g_gct_err, sizeof(g_gct_err), "detokenize: bad arguments");
        return GCT_ERR_ARG;
    }
    int maxlen = 0;
    for (int v = 0; v < V; ++v) maxlen = voff[v + 1] - voff[v] > maxlen ? voff[v + 1] - voff[v] : maxlen;
    if (out_bytes < n * ((int64_t)width * maxlen + 1)) {
        snprintf(g_gct_err, sizeof(g_gct_err), "detokenize: output buffer too small");
        return GCT_ERR_ARG;
    }
    char* o = out;
    for (int64_t r = 0; r < n; ++r) {
        const int16_t* row = ids + r * width;
        for (int j = 0; j < width; ++j) {
            const int t = row[j];
            if (t == eos_id) break;
            if (t == sos_id || t < 0 || t >= V) continue;
            const int len = voff[t + 1] - voff[t];
            memcpy(o, vocab + voff[t], (size_t)len);
            o += len;
        }
        *o++ = '\n';
    }
    return (int64_t)(o - out);
}

// Host-side target-length sampler (Inference/toklen_sampling.py:4-16 `run_sampling`, one Python iteration per draw in the
// reference): per draw one np.random.uniform(0,1) -> first CDF edge >= a, then one np.random.normal() for the half-bin jitter.
// The draws continue NumPy's GLOBAL legacy generator bit for bit: the caller passes np.random.get_state() (MT19937 key / pos /
// has_gauss / cached_gaussian), this loop advances it exactly like RandomState.uniform / .normal would (random_sample =
// 53-bit double from two outputs; legacy polar Gaussian with its cached second value) and hands the state back.
namespace {
struct Mt { uint32_t* key; int pos; };
inline void mt_refill(Mt& m) {
    uint32_t* k = m.key;
    int i = 0;
    for (; i < 624 - 397; ++i) { uint32_t y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu); k[i] = k[i + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
    for (; i < 623; ++i) { uint32_t y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu); k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
    uint32_t y = (k[623] & 0x80000000u) | (k[0] & 0x7fffffffu);
    k[623] = k[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    m.pos = 0;
}
inline uint32_t mt_next(Mt& m) {
    if (m.pos >= 624) mt_refill(m);
    uint32_t y = m.key[m.pos++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
inline double mt_double(Mt& m) {
    const uint32_t a = mt_next(m) >> 5, b = mt_next(m) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
}  // namespace
int gct_toklen_draw(uint32_t* mt_key, int32_t* mt_pos, int32_t* has_gauss, double* cached_gauss, const double* cdf, int n_edges,
                    const double* centres, double width, int64_t n, double* out) {
    if (!mt_key || !mt_pos || !has_gauss || !cached_gauss || !cdf || !centres || !out || n_edges < 2 || n < 0) {
        snprintf(g_gct_err, sizeof(g_gct_err), "toklen_draw: bad arguments");
        return GCT_ERR_ARG;
    }
    Mt m{mt_key, *mt_pos};
    int hg = *has_gauss;
    double cg = *cached_gauss;
    const int nb = n_edges - 1;
    for (int64_t k = 0; k < n; ++k) {
        const double a = 0.0 + (1.0 - 0.0) * mt_double(m);          // RandomState.uniform(0, 1)
        int first = 0;                                              // np.argmax(cdf >= a): 0 when no edge qualifies
        for (int i = 0; i < n_edges; ++i) if (cdf[i] >= a) { first = i; break; }
        int idx = first - 1;
        if (idx < 0) idx += nb;                                     // Python's negative index into the bin centres
        double g;
        if (hg) { g = cg; hg = 0; cg = 0.0; }
        else {
            double x1, x2, r2;
            do { x1 = 2.0 * mt_double(m) - 1.0; x2 = 2.0 * mt_double(m) - 1.0; r2 = x1 * x1 + x2 * x2; } while (r2 >= 1.0 || r2 == 0.0);
            const double f = sqrt(-2.0 * log(r2) / r2);
            cg = f * x1; hg = 1; g = f * x2;
        }
        out[k] = centres[idx] + width * g / 2;
    }
    *mt_pos = m.pos; *has_gauss = hg; *cached_gauss = cg;
    return GCT_OK;
}

// Host half of the seed-faithful latent draw (Inference/sampling_tool.py:93-97: torch.normal on the CPU generator): advances
// PyTorch's CPU MT19937 engine exactly like at::mt19937::operator() (--left == 0 -> regenerate; tempering) and writes the raw
// 32-bit outputs, one per element -- at::normal_fill draws one uniform per element this way before its Box-Muller pass.
// `state` is the engine's 624-word array as stored in torch.get_rng_state() (uint64 per word), left / next its counters.
}  // extern "C"
namespace {
// whole 624-word blocks: regenerate, then temper all 624 in one vectorisable pass (AVX2 build of the same code when the CPU has it)
#define GCT_MT_BLOCK_BODY                                                                                             \
    for (int64_t b = 0; b < nblocks; ++b) {                                                                           \
        int i = 0;                                                                                                    \
        for (; i < 624 - 397; ++i) { const uint32_t y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu); k[i] = k[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu); } \
        for (; i < 623; ++i) { const uint32_t y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu); k[i] = k[i - 227] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu); }       \
        { const uint32_t y = (k[623] & 0x80000000u) | (k[0] & 0x7fffffffu); k[623] = k[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu); }                              \
        uint32_t* o = out + b * 624;                                                                                  \
        for (int j = 0; j < 624; ++j) { uint32_t y = k[j]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18; o[j] = y; }              \
    }
void mt_blocks_generic(uint32_t* k, uint32_t* out, int64_t nblocks) { GCT_MT_BLOCK_BODY }
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) void mt_blocks_avx2(uint32_t* k, uint32_t* out, int64_t nblocks) { GCT_MT_BLOCK_BODY }
#endif
#undef GCT_MT_BLOCK_BODY
}  // namespace
extern "C" {
int gct_mt19937_fill(uint64_t* state, int32_t* left, uint64_t* next, uint32_t* out, int64_t n) {
    if (!state || !left || !next || (!out && n > 0) || n < 0) {
        snprintf(g_gct_err, sizeof(g_gct_err), "mt19937_fill: bad arguments");
        return GCT_ERR_ARG;
    }
    uint32_t k[624];
    for (int i = 0; i < 624; ++i) k[i] = (uint32_t)state[i];
    Mt m{k, (int)*next};
    int l = *left;
    auto one = [&]() {
        if (--l == 0) { mt_refill(m); l = 624; }
        uint32_t y = k[m.pos++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    };
    int64_t i = 0;
    // drain the current block (left counts the outputs remaining in it, + 1), then whole blocks, then the head of the last one
    while (i < n && l > 1) out[i++] = one();
    const int64_t nblocks = (n - i) / 624;
    if (nblocks > 0) {
#if defined(__x86_64__) && defined(__GNUC__)
        if (__builtin_cpu_supports("avx2")) mt_blocks_avx2(k, out + i, nblocks); else
#endif
        mt_blocks_generic(k, out + i, nblocks);
        i += nblocks * 624;
        l = 1; m.pos = 624;          // the state every full block leaves behind: all 624 outputs consumed
    }
    while (i < n) out[i++] = one();
    for (int j = 0; j < 624; ++j) state[j] = k[j];
    *left = l; *next = (uint64_t)m.pos;
    return GCT_OK;
}
}  // extern "C"
// Device half: at::normal_fill's transform (aten/src/ATen/native/cpu/DistributionTemplates.h) on the raw outputs: uniform =
// (raw & (2^24 - 1)) * 2^-24; per block of 16, (u1 = 1 - x[j], u2 = x[j+8]) -> radius = sqrt(-2 log u1), theta = 2 pi u2,
// x[j] = radius cos(theta), x[j+8] = radius sin(theta); a tail shorter than 16 is recomputed from 16 fresh draws (raw[n..n+16))
// over the LAST 16 elements.  Equal to torch.normal(0, 1) on the CPU to the rounding of logf / cosf / sinf (1-2 ulp).
__global__ void normal_from_mt_kernel(const uint32_t* __restrict__ raw, float* __restrict__ z, long long nblocks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one (j, j+8) pair per thread
    const long long blk = i >> 3;
    const int j = (int)(i & 7);
    if (blk >= nblocks) return;
    const uint32_t* src = raw + blk * 16;
    float* dst = z + blk * 16;
    const float u1 = 1.f - (float)(src[j] & 0xFFFFFFu) * 5.9604644775390625e-08f;
    const float u2 = (float)(src[j + 8] & 0xFFFFFFu) * 5.9604644775390625e-08f;
    const float radius = sqrtf(-2.f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    dst[j] = radius * cs;
    dst[j + 8] = radius * sn;
}
extern "C" {
int gct_normal_from_mt(const uint32_t* raw, float* z, int64_t n, void* stream) {
    GCT_REQUIRE(raw && z && n >= 16, "normal_from_mt: needs n >= 16 elements (smaller draws take torch's scalar path)");
    const long long nblocks = n / 16;
    normal_from_mt_kernel<<<(unsigned)cdiv(nblocks * 8, 256), 256, 0, ST(stream)>>>(raw, z, nblocks);
    GCT_LAUNCH_CHECK();
    if (n % 16) {        // the last 16 elements are recomputed from 16 fresh draws (stream order: after the full blocks)
        normal_from_mt_kernel<<<1, 8, 0, ST(stream)>>>(raw + n, z + n - 16, 1);
        GCT_LAUNCH_CHECK();
    }
    return GCT_OK;
}

// ---------------- input pipeline ----------------
int gct_collate(const gct_corpus_t* c, const int64_t* rows, int B, int S, int T, int pad_src, int pad_trg, int sos_id, int eos_id,
                int sep_src, int sep_trg, int64_t* src, int64_t* trg, float* econds_out, float* dconds_out, void* stream) {
    GCT_REQUIRE(c && rows && B >= 0 && S >= 0 && T >= 0, "collate: bad arguments");
    GCT_REQUIRE(c->tok_off && (!src || c->src_ids) && (!trg || c->trg_ids), "collate: corpus arrays missing");
    GCT_REQUIRE(sep_src < 0 || (c->sca_off && c->sca_src_ids && c->sca_trg_ids), "collate: scaffold arrays missing");
    GCT_REQUIRE(c->nconds >= 0 && c->nconds <= 32, "collate: nconds=%d outside [0,32]", c->nconds);
    GCT_REQUIRE(!(econds_out && !c->econds) && !(dconds_out && !c->dconds), "collate: condition arrays missing");
    if (B == 0) return GCT_OK;
    CollateParams p;
    p.src_ids = c->src_ids; p.trg_ids = c->trg_ids; p.tok_off = c->tok_off; p.sca_src_ids = c->sca_src_ids;
    p.sca_trg_ids = c->sca_trg_ids; p.sca_off = c->sca_off; p.econds = c->econds; p.dconds = c->dconds; p.nconds = c->nconds;
    p.n_rows = c->n_rows; p.rows = rows; p.B = B; p.S = S; p.T = T; p.pad_src = pad_src; p.pad_trg = pad_trg; p.sos = sos_id;
    p.eos = eos_id; p.sep_src = sep_src; p.sep_trg = sep_trg; p.src = src; p.trg = trg; p.econds_out = econds_out;
    p.dconds_out = dconds_out;
    collate_kernel<<<B, 128, 0, ST(stream)>>>(p);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}

// ---------------- decode ----------------
size_t gct_decode_workspace_bytes(const gct_config_t* cfg, int B, int Lz, int max_len) {
    if (cfg->dtype == GCT_DTYPE_F32) { DecodeWs<float> W; W.carve(*cfg, B, Lz, max_len, nullptr); return W.bytes; }
    DecodeWs<bf16> W; W.carve(*cfg, B, Lz, max_len, nullptr); return W.bytes;
}
}  // extern "C"
template <typename T>
static int decode_begin_impl(const gct_config_t* cfg, const gct_weights_t* w, const gct_decode_t* d, void* ws, size_t ws_bytes,
                             void* stream) {
    Model<T> m;
    GCT_TRY(m.init(cfg, w, 0, 0, stream));
    DecodeWs<T> W;
    W.carve(*cfg, d->B, d->Lz, d->max_len, ws);
    GCT_REQUIRE(W.bytes <= ws_bytes, "decode: workspace too small (%zu < %zu)", ws_bytes, W.bytes);
    GCT_TRY(decode_begin<T>(m, *d, W));
    for (int pos = 0; pos + 1 < d->prefix_len; ++pos) GCT_TRY(decode_one<T>(m, *d, W, pos, 0, false));
    return GCT_OK;
}
template <typename T>
static int decode_steps_impl(const gct_config_t* cfg, const gct_weights_t* w, const gct_decode_t* d, int s0, int s1, void* ws,
                             size_t ws_bytes, void* stream) {
    Model<T> m;
    GCT_TRY(m.init(cfg, w, 0, 0, stream));
    DecodeWs<T> W;
    W.carve(*cfg, d->B, d->Lz, d->max_len, ws);
    GCT_REQUIRE(W.bytes <= ws_bytes, "decode: workspace too small");
    GCT_REQUIRE(d->n_active >= 0 && d->n_active <= d->B && (d->n_active == 0 || d->rowmap || d->n_active == d->B),
                "decode: n_active %d needs a rowmap and must not exceed B = %d", d->n_active, d->B);
    GCT_REQUIRE(!(d->skip_done || d->rowmap) || !(d->forced || d->probs_out || d->logits_out),
                "decode: active-row decode cannot be combined with forced tokens / per-step probes");
    for (int s = s0; s < s1; ++s) {
        const int pos = d->prefix_len - 1 + s;
        GCT_REQUIRE(pos + 1 < d->max_len, "decode: step %d overflows ys (max_len %d)", s, d->max_len);
        GCT_TRY(decode_one<T>(m, *d, W, pos, s, true));
    }
    return GCT_OK;
}
extern "C" {
int gct_decode_begin(const gct_config_t* cfg, const gct_weights_t* w, const gct_decode_t* d, void* workspace,
                     size_t workspace_bytes, void* stream) {
    GCT_REQUIRE(cfg && w && d && workspace, "decode_begin: null argument");
    return cfg->dtype == GCT_DTYPE_F32 ? decode_begin_impl<float>(cfg, w, d, workspace, workspace_bytes, stream)
                                       : decode_begin_impl<bf16>(cfg, w, d, workspace, workspace_bytes, stream);
}
int gct_decode_steps(const gct_config_t* cfg, const gct_weights_t* w, const gct_decode_t* d, int step_begin, int step_end,
                     void* workspace, size_t workspace_bytes, void* stream) {
    GCT_REQUIRE(cfg && w && d && workspace, "decode_steps: null argument");
    return cfg->dtype == GCT_DTYPE_F32 ? decode_steps_impl<float>(cfg, w, d, step_begin, step_end, workspace, workspace_bytes, stream)
                                       : decode_steps_impl<bf16>(cfg, w, d, step_begin, step_end, workspace, workspace_bytes, stream);
}
int gct_decode_compact(const gct_config_t* cfg, const gct_decode_t* d, int n_out, int32_t* rowmap_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
    GCT_REQUIRE(cfg && d && rowmap_out && workspace, "decode_compact: null argument");
    GCT_REQUIRE(n_out >= 1 && n_out <= d->B, "decode_compact: n_out %d outside [1, %d]", n_out, d->B);
    uint8_t* done;
    size_t bytes;
    if (cfg->dtype == GCT_DTYPE_F32) { DecodeWs<float> W; W.carve(*cfg, d->B, d->Lz, d->max_len, workspace); done = W.done; bytes = W.bytes; }
    else { DecodeWs<bf16> W; W.carve(*cfg, d->B, d->Lz, d->max_len, workspace); done = W.done; bytes = W.bytes; }
    GCT_REQUIRE(bytes <= workspace_bytes, "decode_compact: workspace too small");
    decode_compact_kernel<<<1, 1024, 0, ST(stream)>>>(done, d->B, rowmap_out, n_out);
    GCT_LAUNCH_CHECK();
    return GCT_OK;
}
int gct_decode_launches_per_step(const gct_config_t* cfg) { return 1 + cfg->n_layers * 11 + 3; }
int gct_decode_launches_per_step_at(const gct_config_t* cfg, int B) {
    // the two residual projections of a layer absorb the Norm that follows them (gemm_rownorm.cuh) at large batch
    const bool fused = cfg->dtype != GCT_DTYPE_F32 && cfg->d_model == tc::RN_N && !g_gct_simt_only &&
                       (g_gct_rownorm == 2 || (g_gct_rownorm == 1 && cdiv(B, 128) >= 96));
    return 1 + cfg->n_layers * (fused ? 9 : 11) + 3;
}
int gct_decode_begin_launches(const gct_config_t* cfg, int Lz) {
    const int nck = (cfg->use_cond2lat && cfg->nconds > 0 && !cfg->use_cond2dec) ? 1 : 0;
    if (cfg->dtype != GCT_DTYPE_F32 && DecodeWs<bf16>::want_zmode(*cfg, Lz)) return 2 + nck + cfg->n_layers * (5 + 3 * cfg->heads + 4 * nck);
    return 3 + 2 * cfg->n_layers + nck;
}

int gct_decode_attention(const void* q, int ldq, const void* knew, const void* vnew, int ldnew, void* kcache, void* vcache,
                         int64_t cache_bstride, int pitch, int n_cached, const uint8_t* key_valid, int kv_stride, void* out,
                         int ldo, int B, int H, int dtype, void* stream) {
    GCT_REQUIRE(H >= 1 && H <= 8, "decode attention: H=%d outside [1,8]", H);
    DecAttnParams p;
    p.q = q; p.ldq = ldq; p.knew = knew; p.vnew = vnew; p.ldnew = ldnew; p.kcache = kcache; p.vcache = vcache;
    p.cache_bstride = cache_bstride; p.pitch = pitch; p.n_cached = n_cached; p.key_valid = key_valid; p.kv_stride = kv_stride;
    p.out = out; p.ldo = ldo; p.H = H; p.scale = 0.125f;
    return dtype == GCT_DTYPE_F32 ? launch_decode_attn<float>(p, B, ST(stream)) : launch_decode_attn<bf16>(p, B, ST(stream));
}

// ---------------- data-parallel gradient exchange (train1.py:111-112 DDP) ----------------
// NCCL is resolved at run time from the libnccl.so.2 the process already holds (PyTorch's), so the library has no link-time
// dependency and never brings a second NCCL into the process.
}  // extern "C"
namespace {
typedef struct { char internal[128]; } nccl_uid_t;
struct Nccl {
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
Nccl& nccl() {
    static Nccl n;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (n.ok) return n;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return n;
#define GCT_SYM(name) *(void**)(&n.name) = dlsym(h, "nccl" #name)
    GCT_SYM(GetUniqueId); GCT_SYM(CommInitRank); GCT_SYM(CommDestroy); GCT_SYM(AllReduce); GCT_SYM(GroupStart); GCT_SYM(GroupEnd);
    GCT_SYM(GetErrorString);
#undef GCT_SYM
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.GroupStart && n.GroupEnd && n.GetErrorString;
    return n;
}
#define GCT_NCCL(expr)                                                                                                  \
    do {                                                                                                                \
        int r_ = (expr);                                                                                                \
        if (r_ != 0) GCT_FAIL(GCT_ERR_CUDA, "%s -> %s", #expr, nccl().GetErrorString(r_));                              \
    } while (0)
constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0;

// Gradient buckets keyed by the backward stage after which they are final.
struct DpHook : StageHook {
    void* comm; cudaStream_t st, cs; float* grads; const gct_bucket_t* buckets; int n_buckets;
    cudaEvent_t ev[2 * 16 + 2];
    int done(int stage) override {
        Nccl& n = nccl();
        bool any = false;
        for (int i = 0; i < n_buckets; ++i) any |= buckets[i].stage == stage;
        if (!any) return GCT_OK;
        GCT_CUDA(cudaEventRecord(ev[stage], st));
        GCT_CUDA(cudaStreamWaitEvent(cs, ev[stage], 0));
        GCT_NCCL(n.GroupStart());
        for (int i = 0; i < n_buckets; ++i)
            if (buckets[i].stage == stage && buckets[i].count > 0)
                GCT_NCCL(n.AllReduce(grads + buckets[i].offset, grads + buckets[i].offset, (size_t)buckets[i].count, NCCL_FLOAT32, NCCL_SUM,
                                     comm, cs));
        GCT_NCCL(n.GroupEnd());
        return GCT_OK;
    }
};
// events are created once per (device) and reused: cudaEventDisableTiming events are cheap to record / wait on
int dp_events(cudaEvent_t* out, int n) {
    static cudaEvent_t ev[GCT_MAX_DEVICES][2 * 16 + 2];
    static bool made[GCT_MAX_DEVICES] = {};
    const int dev = gct_cur_device();
    if (!made[dev]) {
        for (int i = 0; i < 2 * 16 + 2; ++i) GCT_CUDA(cudaEventCreateWithFlags(&ev[dev][i], cudaEventDisableTiming));
        made[dev] = true;
    }
    for (int i = 0; i < n; ++i) out[i] = ev[dev][i];
    return GCT_OK;
}
}  // namespace
extern "C" {

int gct_nccl_unique_id(void* out128) {
    GCT_REQUIRE(out128, "nccl_unique_id: null argument");
    if (!nccl().ok) GCT_FAIL(GCT_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
    GCT_NCCL(nccl().GetUniqueId(reinterpret_cast<nccl_uid_t*>(out128)));
    return GCT_OK;
}
int gct_nccl_comm_init(void** comm, int nranks, int rank, const void* id128) {
    GCT_REQUIRE(comm && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "nccl_comm_init: bad arguments");
    if (!nccl().ok) GCT_FAIL(GCT_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
    nccl_uid_t id;
    memcpy(&id, id128, sizeof(id));
    GCT_NCCL(nccl().CommInitRank(comm, nranks, id, rank));
    return GCT_OK;
}
int gct_nccl_comm_destroy(void* comm) {
    if (!comm) return GCT_OK;
    if (!nccl().ok) GCT_FAIL(GCT_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
    GCT_NCCL(nccl().CommDestroy(comm));
    return GCT_OK;
}
int gct_allreduce_grads(void* nccl_comm, float* grads, int64_t n, void* stream) {
    GCT_REQUIRE(nccl_comm && grads && n >= 0, "allreduce_grads: bad arguments");
    if (!nccl().ok) GCT_FAIL(GCT_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
    GCT_NCCL(nccl().AllReduce(grads, grads, (size_t)n, NCCL_FLOAT32, NCCL_SUM, nccl_comm, ST(stream)));
    return GCT_OK;
}
int gct_backward_dp(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, const float* dlogits, const float* dmu,
                    const float* dlog_var, const float* dz, void* workspace, size_t workspace_bytes, void* scratch,
                    size_t scratch_bytes, void* nccl_comm, const gct_bucket_t* buckets_host, int n_buckets, void* comm_stream,
                    void* stream) {
    GCT_REQUIRE(cfg && w && io && workspace && scratch && nccl_comm && buckets_host && n_buckets > 0 && comm_stream,
                "backward_dp: null argument");
    GCT_REQUIRE(w->grads_f32, "backward_dp: gradient buffer missing");
    if (!nccl().ok) GCT_FAIL(GCT_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
    const int nstages = 2 * cfg->n_layers + 1;
    for (int i = 0; i < n_buckets; ++i)
        GCT_REQUIRE(buckets_host[i].stage >= 0 && buckets_host[i].stage < nstages, "backward_dp: bucket %d has stage %d outside [0,%d)", i,
                    buckets_host[i].stage, nstages);
    DpHook hook;
    hook.comm = nccl_comm; hook.st = ST(stream); hook.cs = ST(comm_stream); hook.grads = w->grads_f32; hook.buckets = buckets_host;
    hook.n_buckets = n_buckets;
    GCT_TRY(dp_events(hook.ev, nstages + 1));
    int rc = cfg->dtype == GCT_DTYPE_F32
                 ? backward_impl<float>(cfg, w, io, dlogits, dmu, dlog_var, dz, workspace, workspace_bytes, scratch, scratch_bytes, stream, &hook)
                 : backward_impl<bf16>(cfg, w, io, dlogits, dmu, dlog_var, dz, workspace, workspace_bytes, scratch, scratch_bytes, stream, &hook);
    if (rc != GCT_OK) return rc;
    // the compute stream continues (optimiser step) only after every bucket has been reduced
    GCT_CUDA(cudaEventRecord(hook.ev[nstages], hook.cs));
    GCT_CUDA(cudaStreamWaitEvent(hook.st, hook.ev[nstages], 0));
    return GCT_OK;
}

}  // extern "C"
