/* gct_b200 -- C ABI of the B200-native GCT-Plus Transformer-VAE hot path.
 *
 * The reference (chaoting-sun/GCT-Plus) has no FFI: its boundary is the Python call surface
 *   Model/build_model.py:79-116, Model/forward_propagation1.py:4-48, Model/{vaetf,cvaetf}.py
 *   (forward / encode / decode), Train/trainer1.py:19-30,71-157, Inference/sampling_tool.py:140-184.
 * This header is the C-ABI those Python entry points bind to in gct_plus_b200 (ctypes); each
 * function names the reference code it replaces.  See INTEGRATION.md for the binding stubs.
 *
 * Conventions
 *   - every pointer is a raw CUDA device pointer unless the name ends in _host;
 *   - the library never allocates or frees device memory: parameters, activations, KV cache and
 *     scratch all live in caller-owned buffers (PyTorch tensors) sized by the *_bytes queries;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no hidden syncs, so the
 *     calls can be captured into CUDA graphs;
 *   - return 0 on success, negative on error; gct_last_error() returns a thread-local message;
 *   - dtype 0 = fp32 parity tier (SIMT FFMA GEMMs), 1 = bf16 operands / fp32 accumulate (tcgen05).
 */
#ifndef GCT_B200_H
#define GCT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCT_DTYPE_F32 0
#define GCT_DTYPE_BF16 1

/* parameter slots of the flat parameter buffer (element offsets are supplied by the host) */
#define GCT_SLOT_ENC_EMB 0      /* encoder.embed_sentence.embed.weight [Vs,d]      modules.py:101-110 */
#define GCT_SLOT_ENC_C2E_W 1    /* encoder.embed_cond2enc.weight [nc*d,nc]         cvaetf.py:25-26   */
#define GCT_SLOT_ENC_C2E_B 2
#define GCT_SLOT_ENC_NORM_A 3   /* encoder.norm.alpha                               modules.py:80-95  */
#define GCT_SLOT_ENC_NORM_B 4
#define GCT_SLOT_MULV_W 5       /* fc_mu.weight ; fc_log_var.weight  [2*lat,d]      sublayers.py:7-26 */
#define GCT_SLOT_MULV_B 6
#define GCT_SLOT_DEC_EMB 7      /* decoder.embed.embed.weight [Vt,d] */
#define GCT_SLOT_DEC_C2D_W 8    /* decoder.embed_cond2dec.*                         cvaetf.py:86-87   */
#define GCT_SLOT_DEC_C2D_B 9
#define GCT_SLOT_DEC_C2L_W 10   /* decoder.embed_cond2lat.*                         cvaetf.py:88-89   */
#define GCT_SLOT_DEC_C2L_B 11
#define GCT_SLOT_FCZ_W 12       /* decoder.fc_z.weight [d,lat]                      cvaetf.py:91      */
#define GCT_SLOT_FCZ_B 13
#define GCT_SLOT_DEC_NORM_A 14
#define GCT_SLOT_DEC_NORM_B 15
#define GCT_SLOT_OUT_W 16       /* out.weight [Vt,d]                                cvaetf.py:156     */
#define GCT_SLOT_OUT_B 17
#define GCT_SLOT_PROP_W 18      /* prop_fc.weight [1,Vt]                            cvaetf.py:155     */
#define GCT_SLOT_PROP_B 19
#define GCT_SLOT_ENC_PE 20      /* encoder.pe.pe [200,d] (buffer)                   modules.py:116-131*/
#define GCT_SLOT_DEC_PE 21
#define GCT_NUM_GLOBAL_SLOTS 22
/* encoder layer l: GCT_NUM_GLOBAL_SLOTS + 12*l + {N1A,N1B,QKV_W,QKV_B,O_W,O_B,N2A,N2B,F1_W,F1_B,F2_W,F2_B}
 * decoder layer l: GCT_NUM_GLOBAL_SLOTS + 12*N + 20*l +
 *   {N1A,N1B,QKV_W,QKV_B,O1_W,O1_B,N2A,N2B,Q2_W,Q2_B,KV2_W,KV2_B,O2_W,O2_B,N3A,N3B,F1_W,F1_B,F2_W,F2_B}
 * QKV_W is [3d,d] = q_linear ; k_linear ; v_linear rows, KV2_W is [2d,d] = k_linear ; v_linear. */
#define GCT_ENC_LAYER_SLOTS 12
#define GCT_DEC_LAYER_SLOTS 20

typedef struct {
    int32_t src_vocab, trg_vocab;
    int32_t n_layers, d_model, d_ff, heads, latent_dim;
    int32_t nconds;            /* number of property tokens (0 for vaetf / scavaetf)            */
    int32_t use_cond2dec;      /* cond tokens prepended to the decoder input  (cvaetf.py:103-105) */
    int32_t use_cond2lat;      /* cond tokens prepended to the decoder memory (cvaetf.py:107-116) */
    int32_t dtype;             /* GCT_DTYPE_*                                                   */
    int32_t pad_id;
    float dropout;             /* p of every nn.Dropout on the path; applied only when train!=0 */
} gct_config_t;

typedef struct {
    const float* params_f32;       /* flat fp32 master parameters                                 */
    const void* params_bf16;       /* flat bf16 shadow with identical offsets (dtype bf16) or NULL */
    float* grads_f32;              /* flat fp32 gradient buffer (backward accumulates into it)     */
    const int64_t* slot_offsets_host;   /* HOST array [gct_num_slots()] of element offsets, -1 = absent */
} gct_weights_t;

/* ---- library queries ------------------------------------------------------------------- */
const char* gct_last_error(void);
int gct_version(void);
int gct_sm(void);                                   /* compiled SM target: must be 100 */
int gct_num_slots(int n_layers);
int gct_set_gemm_backend(int simt_only);            /* test hook: 1 routes bf16 GEMMs through the SIMT kernel */
int gct_set_persistent_gemm(int enabled);          /* persistent, TMEM double-buffered GEMM for > 148 tiles (default on) */
int gct_set_attention_backend(int simt_only);       /* test hook: 1 keeps bf16 attention on the SIMT kernel */
int gct_set_ffn_saved_activation(int preact);        /* training FFN: 0 (default) save keep*gelu'(pre) in the forward, 1 save the
                                                       pre-activation and evaluate gelu' + the mask in the backward epilogue */
int gct_set_latent_cross_attention(int enabled);    /* bf16 decode: 1 (default) evaluates cross-attention in latent space when the
                                                       memory has no condition rows, 0 keeps the per-layer K/V form */
int gct_set_decode_attn_config(int cfg);           /* tuning: chunk*100 + ring stages*10 + rows per CTA (0 = default) */
int gct_set_tma_store(int enabled);                 /* TMA tensor stores in the persistent GEMM epilogue: 0 off, 1 (default) on, the wait for a
                                                        store's shared-memory read deferred to the staging tile's next write, 2 on, wait at the store */
int gct_set_cta_pair_gemm(int level);               /* persistent GEMM over CTA pairs (tcgen05.mma.cta_group::2): 0 off, 1 K-major
                                                       operands only, 2 (default) also dgrad / wgrad operand layouts */
int gct_set_residual_box(int mode);                /* bit 0 (default 1): persistent GEMM epilogues fetch the fp32 residual / the multiply-by-aux factor as
                                                       TMA boxes into the staging tile; bit 1 set: the tcgen05 attention kernels store O / dQ / dK / dV (and
                                                       load the saved O) per lane instead of as TMA boxes */
int gct_set_attention_persistent(int mode);        /* tcgen05 attention with persistent CTAs that prefetch the next (batch, head) tile's operands:
                                                       bit 0 the forward (L <= 96), bit 1 the backward (L <= 112); default 2 (the persistent forward measured
                                                       no faster than four one-tile CTAs per SM), 0 = one tile per CTA */
int gct_set_attention_trace(void* dev_buf);         /* debugging: when non-null, the tcgen05 attention kernels write per-CTA phase timestamps
                                                       ([B*H][16] u64: globaltimer ns of phases 0..6, %smid, SM clock of phases 0..6) */
int gct_set_sm_budget(int sms);                     /* persistent GEMMs use at most this many SMs (0 = all): room for a concurrent NCCL kernel */
int gct_set_zattn_config(int ctas_per_sm);          /* tuning: latent-space cross-attention: 0 (default) rows staged in shared memory (CTAs of 4 warps, or of 3
                                                        warps where that puts more rows on an SM), 5 the same with 4-warp CTAs only, 3 / 4 rows straight from
                                                        global memory into registers at 3 / 4 CTAs per SM */
int gct_set_attention_bias_grad_fused(int enabled); /* bias gradients of the q/k/v projections from the tcgen05 attention backward's
                                                       write-out (default on) instead of separate column-sum launches */
int gct_set_epilogue_warps16(int enabled);          /* persistent GEMM: 16 (default) or 8 epilogue warps for the specialised modes */
int gct_set_pdl(int enabled);                       /* programmatic dependent launch on the decode path (default on) */

/* ---- operator level (used by the unit tests and by the Python autograd wrappers) -------- */
/* Norm: Model/modules.py:80-95.  y (dtype T) [, y32] = alpha*(x-mean)/(std+eps)+bias ; x fp32 [rows,d] */
int gct_norm_fwd(const float* x, const float* alpha, const float* bias, void* y, float* y32, int rows, int d,
                 int dtype, void* stream);
int gct_norm_bwd(const float* x, const float* alpha, const float* dy, const float* add, float* dx, float* dalpha,
                 float* dbias, int rows, int d, void* stream);
/* Generic GEMM C[M,N] = A(m,k)*B(n,k) (+bias, GELU, residual): nn.Linear of Model/sublayers.py:54-88.
 * a_mn/b_mn: 0 = K-major storage ([M|N rows, K cols]), 1 = MN-major ([K rows, M|N cols]).
 * flags: 1 GELU (aux_out <- pre-activation), 2 dGELU(aux_in), 4 accumulate into out32, 128 (with 1) aux_out <-
 * keep*gelu'(pre) instead of the pre-activation, 256 multiply by aux_in.  dtype 1 -> tcgen05, 0 -> SIMT fp32. */
int gct_gemm(const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N, int K,
             const float* bias, const float* res32, const void* aux_in, void* aux_out, float* out32, void* outT,
             int ldc, int flags, int split_k, int bn_hint, int dtype, void* stream);
/* Residual projection + the Norm after it in one kernel (Model/layers.py:26-29,62-72 + Model/modules.py:80-95), d_model = 512, bf16:
 * out32 [M,512] = A[M,K] W[512,K]^T + bias + res32 ; normT (bf16) [, norm32] = alpha (out32 - mean) / (std_unbiased + eps) + beta.
 * out32 may alias res32.  gct_set_rownorm_fusion: bits 0-1: 0 (default) = the model keeps the residual GEMM + norm_fwd pair, 1 = fused
 * when >= 96 row tiles, 2 = always; bit 2 set = the fused kernel reads the residual per lane instead of as TMA boxes.
 * [B200] correct (unit + model + decode tests with it forced on) but not faster than the pair it replaces: M=41472 K=512 79 us
 * (96.8 with the per-lane residual) vs 71 us for the pair (single-buffered 512-column accumulator: the three-pass epilogue cannot
 * overlap the next tile's main loop), so it is off by default. */
int gct_gemm_rownorm(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, const float* bias, const float* res32,
                     float* out32, const float* alpha, const float* beta, void* normT, float* norm32, float eps, void* stream);
int gct_set_rownorm_fusion(int mode);
/* attention: Model/sublayers.py:29-41.  q/k/v [B,L,ld] with head h at column 64*h; mask bytes. */
int gct_attention_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* mask,
                      int64_t mask_bstride, int mask_rstride, void* out, int ldo, float* lse, float* probs, int B,
                      int H, int Lq, int Lk, int dtype, void* stream);
int gct_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* mask,
                      int64_t mask_bstride, int mask_rstride, const float* lse, const void* out, int ldo, const void* dout,
                      int lddo, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int B, int H, int Lq, int Lk,
                      int dtype, void* stream);      /* out = the forward result (softmax-backward row term D = rowsum(dout*out)) */
/* masks: Model/modules.py:33-58 as byte arrays */
int gct_src_mask(const int64_t* tok, int B, int L, int nc, int pad, uint8_t* out, void* stream);
int gct_trg_mask(const int64_t* tok, int B, int T, int nc_cond2dec, int pad, uint8_t* out, void* stream);
int gct_mask_cast(const void* in, int elem_size, int64_t n, uint8_t* out, void* stream);
int gct_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);
int gct_cast_bf16_to_f32(const void* in, float* out, int64_t n, void* stream);

/* ---- loss: Train/trainer1.py:19-30 -------------------------------------------------------- */
/* out4 = {loss, RCE_mol, RCE_prop(=0), KLD}.  When dlogits/dmu/dlv are non-null they receive
 * d loss / d{logits,mu,log_var} (scaled by gscale).  scratch: gct_loss_scratch_bytes(). */
size_t gct_loss_scratch_bytes(int64_t rows, int64_t n_latent);
int gct_loss_fwd_bwd(const float* logits, int ld, int V, const int64_t* target, int64_t rows, int pad_id,
                     const float* mu, const float* log_var, int64_t n_latent, float beta, float gscale, float* out4,
                     float* dlogits, float* dmu, float* dlv, void* scratch, void* stream);

/* Property head of use_cond2dec models (Model/cvaetf.py:184-186 `prop_fc` on the first nconds logit rows; Train/trainer1.py:24-26
 * RCE_prop = mse_loss(sum)): prop_out[B,nc] (optional), out4[0] += RCE_prop, out4[2] += RCE_prop (after gct_loss_fwd_bwd filled
 * out4), and with dlogits != NULL the backward: rows [0,nc) of every sample of dlogits [B,Ld,V] are SET to 2 g (prop - y) w,
 * dw[V] / db[1] accumulate the head's parameter gradients. */
int gct_prop_head_fwd_bwd(const float* logits, int B, int Ld, int nc, int V, const float* w, const float* b0, const float* target,
                          float gscale, float* prop_out, float* out4, float* dlogits, float* dw, float* db, void* stream);

/* ---- whole-model forward / backward: Vaetf.forward / Cvaetf.forward (+ .encode/.decode) ------ */
typedef struct {
    const int64_t* src;        /* [B,S]                                  */
    const int64_t* trg;        /* [B,T]  decoder input (already [:, :-1]) */
    const uint8_t* src_mask;   /* [B,Se] key mask, Se = nconds+S          */
    const uint8_t* trg_mask;   /* [B,Ld,Ld] dense mask, Ld = T (+nconds if cond2dec) */
    const float* econds;       /* [B,nconds] or NULL                      */
    const float* dconds;       /* [B,nconds] or NULL                      */
    const float* eps;          /* [B,Se,lat] N(0,1) draw for z, or NULL for z = mu */
    const float* z_in;         /* decode-only: latent [B,Se,lat] supplied by the caller */
    int32_t B, S, T;
    int32_t train;             /* 1: dropout active                        */
    uint32_t seed;             /* dropout seed of this step                */
    int32_t run_encoder, run_decoder;
    float* logits;             /* [B,Ld,Vt] fp32                           */
    float* mu; float* log_var; float* z;      /* [B,Se,lat] fp32           */
    float* enc_attn; float* dec_attn1; float* dec_attn2;   /* optional get_attn outputs [N,B,H,Lq,Lk] */
} gct_io_t;

size_t gct_forward_workspace_bytes(const gct_config_t* cfg, int B, int S, int T);
int gct_forward(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, void* workspace,
                size_t workspace_bytes, void* stream);
/* backward of the forward that last filled `workspace`; gradients of the four outputs may be NULL */
size_t gct_backward_scratch_bytes(const gct_config_t* cfg, int B, int S, int T);
int gct_backward(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, const float* dlogits,
                 const float* dmu, const float* dlog_var, const float* dz, void* workspace, size_t workspace_bytes,
                 void* scratch, size_t scratch_bytes, void* stream);

/* ---- optimiser: Adam + the trainer's LR rule (Train/trainer1.py:112-127, train1.py:116-119) --- */
int gct_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow, int64_t n,
                  int step, float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
double gct_noam_lr(int64_t step, int d_model, int64_t warmup);

/* ---- KV-cached sampler: Sampling.decode, Inference/sampling_tool.py:140-184 ------------------ */
typedef struct {
    int32_t B;                 /* sequences in this batch                                  */
    int32_t Lz;                /* latent length (before cond2lat tokens)                   */
    int32_t max_len;           /* capacity of ys per row: prefix + generated tokens        */
    int32_t prefix_len;        /* tokens already in ys (1 = <sos>, or 2+len(scaffold))      */
    int32_t greedy;            /* 1 greedy (torch.max), 0 multinomial                       */
    int32_t eos_id;
    uint32_t seed;             /* multinomial RNG seed when `uniforms` is NULL              */
    const float* zs;           /* [B,Lz,lat] fp32                                           */
    const uint8_t* src_mask;   /* [B,Lz] latent key mask                                    */
    const float* dconds;       /* [B,nconds] or NULL                                        */
    const float* uniforms;     /* optional [max_steps,B] U(0,1) numbers for multinomial     */
    int64_t* ys;               /* [B,max_len] int64, prefix filled by the caller            */
    int32_t* status;           /* device int[2]: {#rows that emitted <eos>, first step at which all had} */
    const int64_t* forced;     /* optional [B,max_len]: teacher forcing -- step i appends forced[:, prefix_len+i] instead of
                                  the drawn token (scoring given sequences; per-step parity tests); no <eos> bookkeeping */
    float* probs_out;          /* optional [max_steps,B,Vt]: softmax(logits) of every step                       */
    float* logits_out;         /* optional [max_steps,B,Vt]: the logits of every step (model.decode(...)[:, -1])  */
    /* Active-row decode (not in the reference, whose loop re-runs every row until the LAST one has emitted <eos>,
       Inference/sampling_tool.py:144-183).  skip_done = 1: a row that has emitted <eos> appends pad_id from then on and
       its attention work is skipped (the strings id_to_smi builds, :54-61, are unchanged: they end at the first <eos>).
       rowmap / n_active: the step kernels run on n_active COMPACT rows, row i standing for physical row rowmap[i] of
       the caches / ys / uniforms (built by gct_decode_compact between chunks of steps); NULL / 0 = all B rows.        */
    int32_t skip_done;
    int32_t n_active;
    const int32_t* rowmap;
} gct_decode_t;

size_t gct_decode_workspace_bytes(const gct_config_t* cfg, int B, int Lz, int max_len);
/* projects the latent to the cross-attention K/V of every layer, resets the caches, consumes the prefix */
int gct_decode_begin(const gct_config_t* cfg, const gct_weights_t* w, const gct_decode_t* d, void* workspace,
                     size_t workspace_bytes, void* stream);
/* runs decode steps [step_begin, step_end): step i reads ys[:, prefix_len-1+i] and writes ys[:, prefix_len+i] */
int gct_decode_steps(const gct_config_t* cfg, const gct_weights_t* w, const gct_decode_t* d, int step_begin,
                     int step_end, void* workspace, size_t workspace_bytes, void* stream);
/* Active-row decode: writes the physical indices of the rows that have NOT emitted <eos> yet (ascending) into
 * rowmap[0 .. n_active) and pads rowmap[n_active .. n_out) with a finished row; n_out >= n_active is the caller's
 * rounded-up batch (n_active = B - status[0], read by the caller after the previous chunk of steps).  One launch. */
int gct_decode_compact(const gct_config_t* cfg, const gct_decode_t* d, int n_out, int32_t* rowmap_out, void* workspace,
                       size_t workspace_bytes, void* stream);
/* number of kernels one decode step launches (for bench.py's gpu_launches claim) */
int gct_decode_launches_per_step(const gct_config_t* cfg);
int gct_decode_launches_per_step_at(const gct_config_t* cfg, int B);   /* at batch B (large batches fuse two Norms per layer away) */
int gct_decode_begin_launches(const gct_config_t* cfg, int Lz);
/* Host-side batch detokeniser (replaces the per-row Python loop of Inference/sampling_tool.py:54-61 `id_to_smi`): rows of
 * int16 ids [n, width] are cut at the first eos_id, sos_id is dropped, token strings (UTF-8 blob `vocab`, offsets
 * voff[V+1]) are joined, each row ends with '\n'.  out_bytes >= n*(width*max_token_bytes+1).  Returns bytes written. */
int64_t gct_detokenize(const int16_t* ids, int64_t n, int width, const char* vocab, const int32_t* voff, int V, int eos_id,
                       int sos_id, char* out, int64_t out_bytes);
/* Host-side target-length sampler (replaces the per-draw Python loop of Inference/toklen_sampling.py:4-16): n draws from the
 * histogram CDF (`cdf[n_edges]`, bin `centres[n_edges-1]`, bin `width`) with the reference's half-bin Gaussian jitter, consuming
 * NumPy's global legacy MT19937 stream exactly like np.random.uniform / np.random.normal would (state passed in and returned:
 * key[624], pos, has_gauss, cached_gaussian = np.random.get_state()[1:5]).  out[n] doubles. */
int gct_toklen_draw(uint32_t* mt_key, int32_t* mt_pos, int32_t* has_gauss, double* cached_gauss, const double* cdf, int n_edges,
                    const double* centres, double width, int64_t n, double* out);
/* Seed-faithful latent draw (Inference/sampling_tool.py:93-97 `sample_z`: torch.normal on the CPU generator, ~7 ns per element
 * in PyTorch): gct_mt19937_fill advances PyTorch's CPU MT19937 engine (state words / left / next as stored in
 * torch.get_rng_state()) and writes the raw 32-bit outputs, one per element, into a (pinned) host buffer; gct_normal_from_mt
 * applies at::normal_fill's uniform conversion and 16-element Box-Muller blocks on the device (raw must hold n + 16 words when
 * n % 16 != 0).  Same stream, same values to the rounding of logf / sincosf. */
int gct_mt19937_fill(uint64_t* state624, int32_t* left, uint64_t* next, uint32_t* out_host, int64_t n);
int gct_normal_from_mt(const uint32_t* raw, float* z, int64_t n, void* stream);
/* the step's attention kernel on its own (unit tests, roofline timing): one query per (batch, head) over
 * n_cached cached keys (+ this step's knew/vnew row, which is appended to the cache when non-NULL) */
int gct_decode_attention(const void* q, int ldq, const void* knew, const void* vnew, int ldnew, void* kcache, void* vcache,
                         int64_t cache_bstride, int pitch, int n_cached, const uint8_t* key_valid, int kv_stride, void* out,
                         int ldo, int B, int H, int dtype, void* stream);

/* ---- input pipeline: Model/collate_fn.py:4-124 + torchtext-0.6 Field.process, on a corpus that was
 * tokenised once (Utils/dataset.py:251-289 tokenises every row again on every access) and lives in HBM as CSR arrays.
 * One launch assembles a whole batch: src[b] = [scaffold <sep>] smiles <pad>..., trg[b] = <sos> [scaffold <sep>] smiles
 * <eos> <pad>..., econds/dconds rows gathered.  S / T are the batch maxima the host computed from the row lengths. */
typedef struct {
    const int16_t* src_ids;     /* smiles tokens in the SRC vocabulary, all rows back to back      */
    const int16_t* trg_ids;     /* the same tokens in the TRG vocabulary                           */
    const int64_t* tok_off;     /* [n_rows + 1]                                                    */
    const int16_t* sca_src_ids; /* scaffold tokens (SRC / TRG vocabulary) or NULL                  */
    const int16_t* sca_trg_ids;
    const int64_t* sca_off;     /* [n_rows + 1] or NULL                                            */
    const float* econds;        /* [n_rows, nconds] or NULL                                        */
    const float* dconds;
    int32_t nconds;
    int64_t n_rows;
} gct_corpus_t;
int gct_collate(const gct_corpus_t* corpus, const int64_t* rows, int B, int S, int T, int pad_src, int pad_trg, int sos_id,
                int eos_id, int sep_src, int sep_trg, int64_t* src, int64_t* trg, float* econds_out, float* dconds_out,
                void* stream);      /* sep_src < 0: no scaffold prefix */

/* ---- data-parallel gradient exchange (train1.py:111-112: DistributedDataParallel over NCCL) ------------------------ */
/* The host owns the rendezvous (torch.distributed carries the 128-byte id from rank 0 to the others) and the communicator
 * handle; NCCL itself is resolved at run time from the libnccl.so.2 already loaded in the process (PyTorch's), so the
 * library has no link-time NCCL dependency.  All return GCT_ERR_UNSUPPORTED when no libnccl.so.2 can be found. */
int gct_nccl_unique_id(void* out128);                                        /* rank 0: ncclGetUniqueId                  */
int gct_nccl_comm_init(void** comm, int nranks, int rank, const void* id128); /* every rank: ncclCommInitRank (collective) */
int gct_nccl_comm_destroy(void* comm);
/* in-place sum over ranks of a flat fp32 buffer on `stream` */
int gct_allreduce_grads(void* nccl_comm, float* grads, int64_t n, void* stream);
/* gct_backward with the gradient exchange overlapped: the flat gradient buffer is cut into buckets, each tagged with the
 * backward stage after which it is final (stage k < N: decoder layer N-1-k; N+k: encoder layer N-1-k; 2N: the rest --
 * embeddings, heads, final norms).  After each stage an event is recorded on `stream`, `comm_stream` waits for it and the
 * stage's buckets are all-reduced (sum) there while the backward continues; on return `stream` has been made to wait for the
 * last bucket, so the optimiser step can be enqueued directly.  DDP's bucketed overlap (train1.py:111) in one call. */
typedef struct { int32_t stage; int32_t reserved; int64_t offset; int64_t count; } gct_bucket_t;
int gct_backward_dp(const gct_config_t* cfg, const gct_weights_t* w, const gct_io_t* io, const float* dlogits, const float* dmu,
                    const float* dlog_var, const float* dz, void* workspace, size_t workspace_bytes, void* scratch,
                    size_t scratch_bytes, void* nccl_comm, const gct_bucket_t* buckets_host, int n_buckets, void* comm_stream,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCT_B200_H */
