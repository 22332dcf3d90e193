/*
 * libb200enc — C ABI of the B200-native (sm_100a) transformer-encoder hot path.
 *
 * The reference (alibaba-damo-academy/SpokenNLP) has no FFI: its encoder arithmetic is PyTorch library calls
 * made by HuggingFace `BertModel`.  Each entry point below replaces the library call(s) named
 * in its comment; the Python host package `spokennlp_b200` binds them with ctypes behind a
 * drop-in `BertModel`.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - the caller owns all memory (activations, workspaces, outputs); the library never allocates device memory
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 *   - return 0 on success, <0 on error (B200_ERR_*); b200_last_error() gives a thread-local message
 *   - fp16 activations / compute copies of weights ("half"), fp32 master parameters, biases, LN params,
 *     statistics, logits and gradients; all row pitches ("ld") are in ELEMENTS and must be multiples of 8
 */
#ifndef B200ENC_H_
#define B200ENC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_SHAPE (-1)       /* bad shape / alignment / unsupported combination */
#define B200_ERR_DTYPE (-2)
#define B200_ERR_CUDA (-3)        /* CUDA launch / driver error, see b200_last_error() */

/* epilogues of b200_gemm_f16 */
#define B200_EPI_STORE 0          /* out = acc */
#define B200_EPI_BIAS 1           /* out = acc + bias[n]                      (nn.Linear: QKV projection, bert_model.py:271,293-296) */
#define B200_EPI_BIAS_GELU 2      /* z = acc + bias[n]; out = gelu_erf(z); out2 = gelu'(z) saved for the backward (BertIntermediate, bert_model.py:436-439) */
#define B200_EPI_BIAS_RES 3       /* out = acc + bias[n] + aux[m,n]           (dense + residual of BertSelfOutput / BertOutput, :371-375, :449-453) */
#define B200_EPI_DGELU 4          /* out = acc * aux[m,n], aux = gelu'(z) from the forward (autograd of :436-439) */
#define B200_EPI_ADD 5            /* out = acc + aux[m,n]                     (dgrad + residual-branch gradient) */
#define B200_EPI_ATOMIC 6         /* out(fp32) += alpha * acc                 (wgrad, split-K) */
#define B200_EPI_BIAS_RES32 7     /* out(fp32) = acc + bias[n] + aux32[m,n]   (as BIAS_RES with the residual stream kept in fp32) */

#define B200_DT_F16 0
#define B200_DT_F32 1
#define B200_DT_BF16 2 /* b200_gemm_bf16 only */

const char* b200_last_error(void);
int b200_version(void);
/* sha1 over the sources this library was built from (csrc/ + this header), or "unknown" for a hand-run make; the
 * Python binding compares it with the tree it runs in, so a stale library is rebuilt or refused instead of half-working */
const char* b200_source_hash(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
long long b200_launch_count(void);

/*
 * C[M,N] = epilogue(A * B^T) on tcgen05 tensor cores (fp16 x fp16 -> fp32 in TMEM).
 *   a_layout 0: A stored [M,K] row-major (K contiguous)      1: A stored [K,M] row-major (M contiguous)
 *   b_layout 0: B stored [N,K] row-major (torch Linear.weight) 1: B stored [K,N] row-major
 * Replaces torch.nn.functional.linear and its autograd (dgrad: b_layout=1 on the same weight; wgrad:
 * a_layout=b_layout=1 on dY and X) — SURVEY.md K2, K4, K5, K6, K8.
 * Supported (a_layout,b_layout,epilogue,out_dtype): (0,0,{STORE,BIAS,BIAS_GELU,BIAS_RES,BIAS_RES32},{F16,F32*}),
 * (0,1,{STORE,ADD,DGELU},F16), (1,1,ATOMIC,F32).  (*F32 for STORE, BIAS and BIAS_RES32; BIAS_RES32 is F32-only.)
 * alpha: optional device scalar multiplied into the accumulator.  k_splits > 1 only with EPI_ATOMIC.
 */
/* The library keeps no result-affecting global state (no kernel selectors, no debug switches: the round-1 A/B knobs were
 * removed once their measurements were in profiles/).  The one process-wide setting is a RESOURCE reservation:
 * persistent kernels (GEMM, attention: one CTA or CTA pair per SM) size their grids to min(device SMs, sms); 0 = all SMs.
 * A data-parallel caller reserves the SMs its concurrently running NCCL kernels occupy, so that no compute CTA has to
 * wait for a communication kernel to leave before it can start.  Results do not depend on it. */
void b200_set_sm_limit(int sms);
int b200_gemm_f16(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, int M, int N, int K,
                  int epilogue, const float* bias, const void* aux, int ld_aux, void* out, int ld_out, int out_dtype,
                  void* out2, int ld_out2, const float* alpha, int k_splits, void* stream);

/*
 * Fused attention forward, head_dim 64: ctx = softmax(Q K^T / 8 + key_bias) V  (BertSelfAttention.forward,
 * bert_model.py:309-350; HF eager_attention_forward).  Q: [B*Sq, ldq] with head h at columns q_col0 + 64h;
 * K, V: [B*Sk, ldkv] at k_col0 / v_col0 + 64h (self-attention: all three inside the packed [tokens,3H] QKV buffer).
 * key_bias: optional [B,Sk] additive fp32 (0 keep / -inf drop, or any finite additive mask); kv_len: optional [B] int32
 * as produced by b200_mask_to_bias (positive: prefix mask, the bias array is then not read; negative or absent: general bias).
 * ctx: [B*Sq, ld_out] fp16; lse2: optional [B,heads,Sq] fp32 log2-domain log-sum-exp saved for the backward.
 */
int b200_attn_fwd(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const float* key_bias,
                  const int32_t* kv_len, void* ctx, int ld_out, float* lse2, int B, int heads, int Sq, int Sk, void* stream);

/* P[b,h,i,j] for output_attentions=True (ditto/evaluation_ditto.py:121-127); needs lse2 from b200_attn_fwd. */
int b200_attn_probs(const void* q, int ldq, int q_col0, const void* k, int ldk, int k_col0, const float* key_bias,
                    const float* lse2, float* probs, int B, int heads, int Sq, int Sk, void* stream);

/* key_bias[b,s] = mask01 ? 0 : -inf ; kv_len[b] = +(1 + last kept key) for a right-padded row, -(1 + last kept key) when the kept
 * range has holes (attention kernels then read the per-key bias)  (HF create_bidirectional_mask / mmvts -1e6 masks) */
int b200_mask_to_bias(const void* mask, int mask_dtype /*0=int64,1=f32,2=int32*/, float* key_bias, int32_t* kv_len, int B, int S,
                      void* stream);

/* y = LayerNorm(x) * gamma + beta (eps inside sqrt, biased variance; torch.nn.LayerNorm in bert_model.py:368,446).
 * x: [rows,H] fp32 (x_dtype=1) or fp16 (0) pre-LN sum; y fp16; y32 optional fp32 copy; mean/rstd optional [rows]. */
int b200_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, float* y32, float* mean,
                       float* rstd, int rows, int H, float eps, void* stream);

/* LayerNorm backward.  dy (+ optional dy2) fp16, x pre-LN sum, saved mean/rstd.  dx fp16 = grad wrt the pre-LN sum;
 * dgamma/dbeta (and optional dbias = column sums of dx) are ACCUMULATED in fp32, scaled by *alpha (device, optional). */
int b200_layernorm_bwd(const void* dy, const void* dy2, const void* x, int x_dtype, const float* mean, const float* rstd,
                       const float* gamma, void* dx, float* dgamma, float* dbeta, float* dbias, const float* alpha, int rows, int H,
                       void* stream);

/* y = LN(word[ids] + pos[pos_ids] + type[tt])  (BertEmbeddings.forward, bert_model.py:184-210).  Tables fp32.
 * ids/tt/pos: int64 [rows]; tt/pos may be null (0 / row % S); inputs_embeds optional fp32 [rows,H] instead of the gather. */
int b200_embed_ln_fwd(const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* inputs_embeds, const float* word,
                      const float* pos_tab, const float* type_tab, const float* gamma, const float* beta, void* y, float* y32,
                      int rows, int S, int H, float eps, void* stream);
/* Backward of the above (recomputes the pre-LN sum).  Gradients are ACCUMULATED into the fp32 tables, scaled by *alpha.
 * pad_id: nn.Embedding(padding_idx=config.pad_token_id) semantics (bert_model.py:171) — rows gathered with that id take part
 * in the forward but add nothing to dword; -1 = no padding index.  When the forward ran on inputs_embeds (ids == null) pass
 * the same inputs_embeds; d_inputs_embeds (optional fp32 [rows,H]) then receives the gradient instead of dword. */
int b200_embed_ln_bwd(const void* dy, const void* dy2, const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* word,
                      const float* pos_tab, const float* type_tab, const float* gamma, float* dword, float* dpos, float* dtype_tab,
                      float* dgamma, float* dbeta, const float* alpha, int rows, int S, int H, float eps, long long pad_id,
                      const float* inputs_embeds, float* d_inputs_embeds, void* stream);

/* Token-classification head: logits[rows,C] = h . W^T + b, C in {2,3} (LossCalculator.classifier, loss_calculator.py:17,42;
 * modeling_ponet.py:43,83-84; TSSP tssp.py:26-34).  argmax optional int32 [rows] (np.argmax, ts_sentence_seq_labeling.py:1143). */
int b200_cls_head_fwd(const void* h, const float* W, const float* b, float* logits, int32_t* argmax, int rows, int H, int C,
                      void* stream);
/* stats[0] += sum_i w_i * nll_i, stats[1] += sum_i w_i over rows with label != -100 (CrossEntropyLoss, utils.py:173-182). */
int b200_ce_stats(const float* logits, const int64_t* labels, const float* class_weight, float* stats, int rows, int C, void* stream);
/* backward of head + mean CE: dh = *scale * dlogits . W (fp16), dW += dlogits^T h, db += sum dlogits. */
int b200_cls_head_bwd(const void* h, const float* logits, const int64_t* labels, const float* class_weight, const float* stats,
                      const float* W, const float* scale, void* dh, float* dW, float* db, int rows, int H, int C, void* stream);

/*
 * PoNet pooling mixer forward (PoNetSelfAttention of modelscope 1.1.0, called at alimeeting4mug/src/models/modeling_ponet.py:68-79;
 * algorithm restated in oracle/ponet_oracle.py — source absent, parity unpinned).  proj: fp16 [B*S, ld] holding the five
 * projections [Q | K | O | Sg | Lc] (5H columns); key_bias: optional [B,S] (non-zero = padding); segment_ids int64 [B,S],
 * monotone, values in [0, nseg).  out: fp16 [B*S, H] = (global + segment-max) * O + local-max3, zero on padding.
 * workspace: caller-owned, b200_ponet_workspace() bytes.
 */
size_t b200_ponet_workspace(int B, int S, int H, int heads, int nseg);
int b200_ponet_mix_fwd(const void* proj, int ld, const float* key_bias, const int64_t* segment_ids, void* workspace, void* out, int B, int S,
                       int H, int heads, int nseg, void* stream);

/* Backward of the mixer: dproj [B*S, ld_d] receives the gradients wrt [Q | K | O | Sg | Lc] (fp16, same scaling as dout).
 * fwd_workspace: the workspace b200_ponet_mix_fwd filled for this layer (kept by the caller); bwd_workspace: scratch of
 * b200_ponet_bwd_workspace() bytes. */
size_t b200_ponet_bwd_workspace(int B, int S, int H, int heads, int nseg);
int b200_ponet_mix_bwd(const void* proj, int ld, const void* dout, const float* key_bias, const int64_t* segment_ids, const void* fwd_workspace,
                       void* bwd_workspace, void* dproj, int ld_d, int B, int S, int H, int heads, int nseg, void* stream);

/* db[n] += *alpha * sum_m dy[m,n]   (bias gradients) */
int b200_colsum(const void* dy, int ld, float* db, const float* alpha, int rows, int cols, void* stream);

/* casts: n must be a multiple of 8 */
int b200_cast_f32_to_f16(const float* src, void* dst, size_t n, void* stream);
int b200_cast_f16_to_f32(const void* src, float* dst, size_t n, void* stream);
/* scale[0] = power of two bringing amax|src| to ~target, scale[1] = 1/scale[0]; dst = fp16(src * scale[0]).
 * `amax_slot` is a 4-byte device scratch.  Used to carry an fp32 upstream gradient into the fp16 backward. */
int b200_scale_cast_grad(const float* src, void* dst, size_t n, float target, float* scale, void* amax_slot, void* stream);
/* dst(fp32) = src(fp16) * scale[1]: hands a scaled fp16 gradient back to an fp32 caller */
int b200_unscale_cast_grad(const void* src, float* dst, size_t n, const float* scale, void* stream);

/*
 * Fused attention backward (autograd of bert_model.py:309-350), head_dim 64.  Inputs as b200_attn_fwd plus
 * dctx/ctx [B*Sq, heads*64] fp16 and the saved lse2.  Writes dq into dq[:, dq_col0 + 64h ...] (fp16, pitch ld_dq) and
 * dk / dv into dkv[:, dk_col0 / dv_col0 + 64h ...] (pitch ld_dkv) — for self-attention all three are column ranges of
 * one packed [tokens, 3H] gradient buffer, ready to be the A operand of the QKV dgrad/wgrad GEMMs.
 * workspace: caller-owned, b200_attn_bwd_workspace(B, heads, Sq) bytes (delta + fp32 dQ accumulator).
 */
size_t b200_attn_bwd_workspace(int B, int heads, int Sq);
int b200_attn_bwd(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const void* dctx, int ld_dctx,
                  const void* ctx, int ld_ctx, const float* key_bias, const int32_t* kv_len, const float* lse2, void* workspace,
                  void* dq, int ld_dq, int dq_col0, void* dkv, int ld_dkv, int dk_col0, int dv_col0, int B, int heads, int Sq, int Sk,
                  void* stream);

/* Optimizer over the flat fp32 parameter / gradient buffers (HF Trainer defaults: AdamW, clip_grad_norm_ 1.0).
 * sumsq[0] += sum g^2 (caller zeroes); coef = {grad multiplier incl. clipping, finite flag, norm};
 * adamw_step skips the update when coef[1] == 0 and always refreshes the fp16 compute copy p16 (optional). */
int b200_grad_sumsq(const float* g, size_t n, float* sumsq, void* stream);
int b200_clip_coef(const float* sumsq, float max_norm, float grad_mult, float* coef, void* stream);
/* as b200_clip_coef, plus dynamic loss scaling on the device (GradScaler semantics; replays inside a CUDA graph): a step with a
 * non-finite norm is skipped, loss_scale = {scale, 1/scale} is multiplied by `backoff` and state[1] (skipped steps) counts it;
 * after `growth_interval` consecutive finite steps the scale is multiplied by `growth`.  state = {good steps, skipped steps}. */
int b200_clip_coef_scaled(const float* sumsq, float max_norm, float grad_mult, float* coef, float* loss_scale, float* state,
                          int growth_interval, float backoff, float growth, float min_scale, float max_scale, void* stream);
int b200_adamw_step(float* p, const float* g, float* m, float* v, void* p16, size_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, float bias_corr1, float bias_corr2, const float* coef, void* stream);
/* Same with the hyper-parameters {lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, -} read from device memory, so a
 * captured CUDA graph of the whole step can be replayed while the schedule advances (b200_set_hyper runs outside the graph). */
int b200_set_hyper(float* hyper, float lr, float beta1, float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2,
                   void* stream);
int b200_adamw_step_dev(float* p, const float* g, float* m, float* v, void* p16, size_t n, const float* hyper, const float* coef,
                        void* stream);
/* Same, and the gradient buffer is zeroed on the way out: the next step's accumulating gradient reductions need no
 * separate fill pass (the caller must not read `g` afterwards). */
int b200_adamw_step_dev_zero(float* p, float* g, float* m, float* v, void* p16, size_t n, const float* hyper, const float* coef,
                             void* stream);

/*
 * Dropout-enabled variants (reference: nn.Dropout at bert_model.py:209 (embeddings), :338 (attention probabilities), :373 and
 * :451 (hidden, before the residual), bert_for_ts.py:66-67 (classifier input)).  Masks are a stateless hash of
 * (*seed, site, element index): nothing is stored, the backward regenerates the forward's mask from the same (seed, site).
 * `seed` is a DEVICE pointer to the per-step base seed (so a replayed CUDA graph draws fresh masks); seed == NULL or p == 0
 * turns dropout off and makes each call identical to its plain counterpart.  Kept elements are scaled by 1/(1-p).
 */
int b200_gemm_f16_drop(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, const float* bias, const void* aux,
                       int ld_aux, void* out, int ld_out, int out_dtype, const uint32_t* seed, unsigned site, float p, void* stream);
/* as b200_layernorm_bwd; additionally dx_drop = dx * mask/(1-p) (what the dense layer's dgrad/wgrad consume; dbias sums dx_drop) */
int b200_layernorm_bwd_drop(const void* dy, const void* dy2, const void* x, int x_dtype, const float* mean, const float* rstd, const float* gamma,
                            void* dx, void* dx_drop, float* dgamma, float* dbeta, float* dbias, const float* alpha, int rows, int H,
                            const uint32_t* seed, unsigned site, float p, void* stream);
int b200_embed_ln_fwd_drop(const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* inputs_embeds, const float* word,
                           const float* pos_tab, const float* type_tab, const float* gamma, const float* beta, void* y, float* y32, int rows, int S,
                           int H, float eps, const uint32_t* seed, unsigned site, float p, void* stream);
int b200_embed_ln_bwd_drop(const void* dy, const void* dy2, const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* word,
                           const float* pos_tab, const float* type_tab, const float* gamma, float* dword, float* dpos, float* dtype_tab,
                           float* dgamma, float* dbeta, const float* alpha, int rows, int S, int H, float eps, long long pad_id,
                           const float* inputs_embeds, float* d_inputs_embeds, const uint32_t* seed, unsigned site, float p, void* stream);
int b200_attn_fwd_drop(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const float* key_bias,
                       const int32_t* kv_len, void* ctx, int ld_out, float* lse2, int B, int heads, int Sq, int Sk, const uint32_t* seed,
                       unsigned site, float p, void* stream);
int b200_attn_bwd_drop(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const void* dctx, int ld_dctx,
                       const void* ctx, int ld_ctx, const float* key_bias, const int32_t* kv_len, const float* lse2, void* workspace,
                       void* dq, int ld_dq, int dq_col0, void* dkv, int ld_dkv, int dk_col0, int dv_col0, int B, int heads, int Sq, int Sk,
                       const uint32_t* seed, unsigned site, float p, void* stream);
int b200_cls_head_fwd_drop(const void* h, const float* W, const float* b, float* logits, int32_t* argmax, int rows, int H, int C,
                           const uint32_t* seed, unsigned site, float p, void* stream);
int b200_cls_head_bwd_drop(const void* h, const float* logits, const int64_t* labels, const float* class_weight, const float* stats, const float* W,
                           const float* scale, void* dh, float* dW, float* db, int rows, int H, int C, const uint32_t* seed, unsigned site,
                           float p, void* stream);

/* ---- fused epilogue variants (written after round 1, validated and measured on a B200 in round 2 — profiles/r02a_* —
 * and used by the default layer schedule, spokennlp_b200/blocks.py) -------------------------------------------------------
 *
 * b200_gemm_f16_resadd: out32[M,N] += drop(A[M,K] * B[N,K]^T + bias[n]).  `out` already holds the fp32 residual (what
 *   LayerNorm wrote), so `LN(dropout(dense(x)) + residual)` (bert_model.py:371-375, :449-453) needs no residual read in the
 *   epilogue: tiles leave through TMA reduce-add.
 * b200_gemm_f16_dgrad_delta: out16[M,N] = A[M,K] * B[K,N] (B row-major [K,N]: the output projection's dgrad), and
 *   delta[b,h,q] = sum_d out[b*Sq+q, 64h+d] * ctx[b*Sq+q, 64h+d] — the row statistic of attention backward, written to the
 *   head of the attention-backward workspace (b200_attn_bwd_delta_ptr) so that b200_attn_bwd_ext can skip its own pass.
 * b200_gemm_f16_dgelu_colsum: dz = (A . W) o dact (the FFN-down dgrad with the dGELU epilogue: A = d_dense [M,K] fp16,
 *   W [K,N] row-major = the dense layer's [out,in] weight read MN-major in place, dact = gelu'(z) saved by the forward) AND,
 *   in the same epilogue, colsum[n] += *col_alpha * sum_m dz[m,n] — the bias gradient of BertIntermediate.dense
 *   (bert_model.py:436-439), taken from the fp16-rounded tile while it sits in the store staging slab instead of re-reading
 *   dz from HBM (b200_colsum).
 * b200_attn_bwd_ext: b200_attn_bwd_drop + flags. */
#define B200_ATTN_BWD_DELTA_READY 1   /* workspace already holds delta (from b200_gemm_f16_dgrad_delta) */
#define B200_ATTN_BWD_DQ_HALF 2       /* dQ is accumulated as fp16 TMA reduce-adds straight into dq (zeroed by the call): no fp32
                                       * accumulator, memset or cast pass; at most Sk/128 roundings per element instead of one */
int b200_gemm_f16_resadd(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, float* out, int ld_out,
                         const uint32_t* seed, unsigned site, float p, void* stream);
int b200_gemm_f16_dgrad_delta(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const void* ctx, int ld_ctx, void* out,
                              int ld_out, float* delta, int heads, int Sq, void* stream);
int b200_gemm_f16_dgelu_colsum(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const void* dact, int ld_dact, void* out,
                               int ld_out, float* colsum, const float* col_alpha, void* stream);
float* b200_attn_bwd_delta_ptr(void* workspace);
int b200_attn_bwd_ext(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const void* dctx, int ld_dctx,
                      const void* ctx, int ld_ctx, const float* key_bias, const int32_t* kv_len, const float* lse2, void* workspace,
                      void* dq, int ld_dq, int dq_col0, void* dkv, int ld_dkv, int dk_col0, int dv_col0, int B, int heads, int Sq, int Sk,
                      const uint32_t* seed, unsigned site, float p, int flags, void* stream);

/* ---- loss heads of the topic-segmentation wrapper on the labelled [BOS] rows (csrc/heads.cuh) ---------
 * Replaces the per-example Python loops, boolean-mask gathers and host syncs of
 *   emnlp2023-topic_segmentation/src/models/modules/loss_calculator.py:38-71, utils.py:116-182, tssp.py:26-34, cssl.py:20-273.
 * All activations are fp32 (the dtype the drop-in encoder hands to autograd).  Index lists are int32 device arrays:
 *   b200_heads_compact    positions with key != ignore, row-major: idx (flat b*S+s), ex (example), rank (inside the example),
 *                         cnt/start [B], totals = {n, max cnt} (read back by the host once; depends on the labels only)
 *   b200_heads_topic_ids  topic id of every labelled row (cssl.py:252-263)
 *   gather/scatter_rows   R[i] = h[idx[i]] / dh[idx[i]] += scale * dR[i]
 *   segmax_fwd/bwd        cssl.py:236-247 segment amax + slot gather (gradient split evenly between ties)
 *   normalize(_bwd)       x / max(|x|, eps): the operands of F.cosine_similarity
 *   pair_cos_fwd/bwd      utils.py:116-138 cosine of each labelled row with the next one of its example (cyclic), / temp,
 *                         scattered into out[b, rank] (caller pre-fills -100)
 *   bce_fwd/bwd           the "cos" score predictor's BCE-with-logits terms (loss_calculator.py:45-49)
 *   cssl_matrix_fwd/bwd   cssl.py:20-72 ("eop_matrix"): out[0] += weight * loss; E [n,n], num, den, coef feed the backward
 *   cssl_list_fwd/bwd     cssl.py:86-166 ("eop_list") on host-drawn index lists pos [kp,n], neg [kn,n]
 *   rows_ce_fwd/bwd       Linear(H,C) + CE on compact rows (TSSP, tssp.py:26-34): stats[0] += sum nll
 *   cls_fwd / focal_stats / cls_bwd   the full-position Linear(H,C) head with CE or FocalLoss (utils.py:141-182) on fp32
 *                         activations: stats = {sum w nll, sum w, sum_i (1-p_i)^gamma}; dh/dW/db are ACCUMULATED.
 * Backward entry points take `scale` (host) and `gscale` (optional device scalar: the upstream gradient of the loss). */
int b200_heads_compact(const int64_t* key, long long ignore, int B, int S, int32_t* tmp, int32_t* cnt, int32_t* start, int32_t*
    totals, int32_t* idx, int32_t* ex, int32_t* rank, void* stream);
int b200_heads_gather_keys(const int64_t* key, const int32_t* idx, int n, int32_t* vals, void* stream);
int b200_heads_topic_ids(const int32_t* lab, const int32_t* ex, int n, int32_t* seg, void* stream);
int b200_heads_gather_rows(const float* h, const int32_t* idx, int n, int H, float* R, void* stream);
int b200_heads_scatter_rows(const float* dR, const int32_t* idx, int n, int H, float scale, float* dh, void* stream);
int b200_heads_segmax_fwd(const float* h, const int64_t* seg_ids, const int32_t* slot_ex, const int32_t* slot_id, int nf, int S,
    int H, float* F, void* stream);
int b200_heads_segmax_bwd(const float* h, const int64_t* seg_ids, const int32_t* slot_ex, const int32_t* slot_id, const float*
    F, const float* dF, int nf, int S, int H, float scale, float* dh, void* stream);
int b200_heads_normalize(const float* X, int n, int H, float eps, float* Xn, float* inv, void* stream);
int b200_heads_normalize_bwd(const float* Xn, const float* inv, const float* dXn, int n, int H, float scale, const float*
    gscale, float* dX, void* stream);
int b200_heads_pair_cos_fwd(const float* Rn, const int32_t* ex, const int32_t* rank, const int32_t* start, const int32_t* cnt,
    int n, int H, float temp, float* cos_rows, float* out, int ld, void* stream);
int b200_heads_pair_cos_bwd(const float* Rn, const int32_t* ex, const int32_t* rank, const int32_t* start, const int32_t* cnt,
    const float* g, int n, int H, float temp, float* dRn, void* stream);
int b200_heads_bce_fwd(const float* cos_rows, const int32_t* lab, int n, float* stats, void* stream);
int b200_heads_bce_bwd(const float* cos_rows, const int32_t* lab, int n, float scale, const float* gscale, float* g, void*
    stream);
int b200_heads_cssl_matrix_fwd(const float* Fn, const int32_t* seg, int n, int H, float temp, float weight, float* E, float*
    num, float* den, float* coef, float* out, void* stream);
int b200_heads_cssl_matrix_bwd(const float* Fn, const int32_t* seg, const float* E, const float* num, const float* den, const
    float* coef, int n, int H, float temp, float* dFn, void* stream);
int b200_heads_cssl_list_fwd(const float* Fn, const int32_t* pos, const int32_t* neg, int kp, int kn, int n, int H, float temp,
    float weight, float* out, float* gw, void* stream);
int b200_heads_cssl_list_bwd(const float* Fn, const int32_t* pos, const int32_t* neg, const float* gw, int kp, int kn, int n,
    int H, float* dFn, void* stream);
int b200_heads_rows_ce_fwd(const float* R, const float* W, const float* bias, const int32_t* tgt, int n, int H, int C, float*
    probs, float* stats, void* stream);
int b200_heads_rows_ce_bwd(const float* R, const float* W, const int32_t* tgt, const float* probs, int n, int H, int C, float
    scale, const float* gscale, float* dR, float* dW, float* db, void* stream);
int b200_heads_cls_fwd(const float* h, const float* W, const float* b, float* logits, int32_t* argmax, int rows, int H, int C,
    void* stream);
int b200_heads_focal_stats(const float* logits, const int64_t* labels, const float* class_weight, float gamma, int rows, int C,
    float* stats, void* stream);
int b200_heads_cls_bwd(const float* h, const float* logits, const int64_t* labels, const float* class_weight, const float*
    stats, const float* W, float gamma, float scale, const float* gscale, int rows, int H, int C, float* dh, float* dW, float*
    db, void* stream);

/* ---- packed (variable-length) rows: SURVEY.md §8f rank 2 --------------------------------------------------------------------
 * The reference right-pads every window to max_seq_length (ts_sentence_seq_labeling.py:862-873) and the encoder computes the
 * padding.  Packing keeps only the valid tokens: b200_heads_compact on the attention mask (ignore = 0) IS the packer — idx =
 * flat positions of the valid tokens in row-major order, start = cu_seqlens [B+1], ex / rank = sequence and position of every
 * packed row, totals = {rows, longest sequence}.  Every row-wise kernel and GEMM then runs on `rows` rows; attention walks
 * sequence b over rows [cu[b], cu[b+1]) with S_max shaping lse2 and the dropout indices (so a packed run draws the same
 * attention masks as the padded one).  b200_gather_i64 packs ids / token types / labels; b200_unpack_rows scatters fp32 rows back
 * into the padded layout. */
int b200_gather_i64(const int64_t* key, const int32_t* idx, int n, int64_t* out, void* stream);
int b200_unpack_rows(const float* src, const int32_t* idx, int n, int H, float* dst, void* stream);
int b200_attn_fwd_varlen(const void* qkv, int ld, int q_col0, int k_col0, int v_col0, const int32_t* cu_seqlens, long long rows, void* ctx,
                         int ld_out, float* lse2, int B, int heads, int S_max, const uint32_t* seed, unsigned site, float p, void* stream);
size_t b200_attn_bwd_workspace_varlen(int B, int heads, int S_max, long long rows);
int b200_attn_bwd_varlen(const void* qkv, int ld, int q_col0, int k_col0, int v_col0, const void* dctx, int ld_dctx, const void* ctx, int ld_ctx,
                         const int32_t* cu_seqlens, const int32_t* row_ex, const int32_t* row_rank, long long rows, const float* lse2,
                         void* workspace, void* dqkv, int ld_d, int dq_col0, int dk_col0, int dv_col0, int B, int heads, int S_max,
                         const uint32_t* seed, unsigned site, float p, void* stream);

/* ---- bf16 operands (BASELINE config 4: "mmvts ... fine-tune bf16") ----------------------------------------------------------
 * The same 2-CTA tcgen05 GEMM with bf16 A / B (instruction-descriptor formats 1/1, fp32 accumulation): forward Linear
 * (a_layout = b_layout = 0; EPI_STORE / EPI_BIAS; result bf16 or fp32), dgrad (b_layout = 1, EPI_STORE, bf16) and wgrad
 * (a_layout = b_layout = 1, EPI_ATOMIC split-K, fp32).  Same tensor-core rate as fp16, 8 mantissa bits: the hidden-state
 * parity bar of the BERT path (1e-3) is NOT met with bf16 operands (5.4e-3 at layer 12), which is why the encoder
 * runs fp16; the mmvts projector (the largest GEMM of that model, K up to 3328) offers it as an option and reports its error. */
int b200_gemm_bf16(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, int M, int N, int K, int epilogue, const float* bias,
                   void* out, int ld_out, int out_dtype, const float* alpha, int k_splits, void* stream);
int b200_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200ENC_H_ */
