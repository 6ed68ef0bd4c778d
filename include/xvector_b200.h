/*
 * xvector_b200 -- C ABI of the B200-native x-vector hot path (libxvector_b200.so).
 *
 * The reference (mycrazycracy/tf-kaldi-speaker) has no FFI: its operator layer is Python
 * (model/tdnn.py, model/pooling.py, model/loss.py, model/trainer.py) lowering to stock
 * TensorFlow kernels.  Each entry point below therefore cites the reference graph op(s) it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, a negative xv_status otherwise; xv_last_error()
 *     gives a thread-local message.  Nothing here allocates device memory or synchronises:
 *     the caller owns all buffers and passes raw device pointers + a cudaStream_t (as void*).
 *   - activations are channels-last, "flat-time": a batch [B, T, C] is a row-major matrix
 *     [B*T, ld] whose row m = b*T + t; a layer's output keeps the SAME row stride T and the
 *     rows with t >= T - shrink are don't-care ("invalid") rows.
 *   - bf16 = __nv_bfloat16 bits (uint16_t), f32 = float.
 *   - there is NO CPU fallback: on a machine without an sm_100 GPU every compute entry
 *     returns XV_ERR_CUDA.
 */
#ifndef XVECTOR_B200_H_
#define XVECTOR_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define XV_API __attribute__((visibility("default")))
#else
#define XV_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  XV_OK = 0,
  XV_ERR_INVALID = -1,     /* bad argument (shape, alignment, null pointer)            */
  XV_ERR_UNSUPPORTED = -2, /* valid request this build does not implement               */
  XV_ERR_CUDA = -3         /* CUDA runtime / driver error (message holds the CUDA text) */
} xv_status;

XV_API const char* xv_last_error(void);
XV_API int xv_version(void);
/* Device properties the host layer sizes grids with. out[0]=SM count, out[1]=cc major, out[2]=cc minor. */
XV_API int xv_device_info(int32_t out[3]);

/* ------------------------------------------------------------------------------------------
 * Generic implicit-GEMM on tcgen05/TMEM fed by TMA:  D[M,N] (+)= sum_k A(m,k) * B(n,k)
 * Replaces: tf.layers.conv2d (model/tdnn.py:39,57,75) and tf.layers.dense (tdnn.py:96,115,147,166)
 * forward, plus the Conv2DBackpropInput / Conv2DBackpropFilter / MatMul-grad kernels TF autodiff
 * inserts for them (model/trainer.py:403), and tf.matmul of the heads (model/loss.py:108,214,300).
 *
 * An operand is a row-major bf16 matrix [rows, cols] (cols contiguous, row stride ld elements).
 *   mn_major = 0 ("K-major"):  rows index M (or N), cols index K.
 *       tile coordinates for reduction index kk:  col = kk % div,  row = mn0 + (kk / div) * tap_rows
 *       (div = 0: col = kk, row = mn0).  This is how a width-k temporal convolution becomes a GEMM
 *       without im2col: tap j of the window of output row m is input row m + j.
 *   mn_major = 1 ("MN-major"): rows index K, cols index M (or N).
 *       col = (mn0 % div) + 64*c,  row = kk + (mn0 / div) * tap_rows   (div = 0: col = mn0 + 64*c, row = kk)
 * Out-of-range rows/cols (including negative rows) read as zero (TMA zero fill).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* ptr;
  int64_t rows, cols, ld;
  int32_t mn_major;
  int32_t div;
  int32_t tap_rows;
  int32_t _pad;
} xv_operand;

typedef enum {
  XV_EPI_BF16 = 0,      /* out bf16 [M, ldc] = acc (+ bias[n]); optional per-column sum / sum-of-squares of out */
  XV_EPI_F32 = 1,       /* out f32  [M, ldc] = acc (+ bias[n]); splits > 1 => atomic accumulate into out  */
  XV_EPI_HEAD_FWD = 2,  /* fused margin-softmax forward: online log-sum-exp partials, logits never stored  */
  XV_EPI_HEAD_BWD = 3   /* fused margin-softmax backward: recompute logits tile, emit dLoss/dlogit as bf16 */
} xv_epilogue;

/* Per-row margin description shared by the two head epilogues (model/loss.py:97-159,207-247,293-345). */
typedef enum { XV_HEAD_SOFTMAX = 0, XV_HEAD_ASOFTMAX = 1, XV_HEAD_AM = 2, XV_HEAD_AAM = 3 } xv_head_type;

typedef struct {
  int32_t type;          /* xv_head_type                                                            */
  int32_t asoftmax_m;    /* 1, 2 or 4 (XV_HEAD_ASOFTMAX)                                             */
  float margin;          /* m of AM / AAM                                                           */
  float cos_m, sin_m;    /* cos(m), sin(m) of AAM, evaluated by the host                            */
  float threshold;       /* cos(pi - m) of AAM (loss.py:321)                                        */
  const float* sched;    /* device [2] = {fa, fs} = {1/(1+lambda), 1-fa}; lambda schedule of loss.py:144-147 is
                            evaluated by the host each step and lives in device memory so a captured CUDA graph
                            of the step stays valid as global_step advances                                */
  const int32_t* labels; /* [M]                                                                     */
  const float* xnorm;    /* [M] max(||x_i||, 1e-12)                                                 */
  /* forward outputs */
  float* part_max;       /* [2*ceil(N/128), M]: online-LSE running max per 64-column half tile       */
  float* part_sum;       /* [2*ceil(N/128), M]: matching sum of exp(z - part_max)                   */
  float* target_logit;   /* [M] modified target logit z'_{i,y_i}                                    */
  float* logits_out;     /* optional [M, ldc] f32 pre-margin logits (endpoints["logits"]); may be 0 */
  /* backward inputs / outputs */
  const float* lse;      /* [M] log-sum-exp of the modified logits                                  */
  float inv_batch;       /* 1 / global batch                                                        */
  float* gnorm;          /* [M] dLoss/d||x_i|| through the margin term                              */
} xv_head_args;

/* Optional fusion for a dgrad GEMM (XV_EPI_BF16 producing dX = dLoss/d act(BN(y))): the epilogue also accumulates the
 * batch-norm backward reductions of the layer that owns y,
 *   col_sum[n]   += sum_m g,   col_sumsq[n] += sum_m g * (y - mean[n]) * rstd[n],   g = dX * act'(y*scale[n] + shift[n])
 * (= dbeta / dgamma of tf.layers.batch_normalization), saving the separate xv_bn_act_bwd_reduce pass over y and dX.
 * act' is 1 for z > 0 and neg_slope otherwise (relu 0, leaky_relu 0.2, identity 1).  y = NULL disables the fusion. */
typedef struct {
  const void* y;         /* bf16 [M, ldy]                                                            */
  int64_t ldy;
  const float* scale;    /* [N] gamma * rstd                                                         */
  const float* shift;    /* [N] beta - mean * scale                                                  */
  const float* mean;     /* [N] batch mean saved by the forward pass                                 */
  const float* rstd;     /* [N]                                                                      */
  float neg_slope;
  int32_t _pad;
} xv_bn_bwd_args;

typedef struct {
  xv_operand a, b;
  int32_t M, N, K;
  int32_t splits;        /* split-K factor (>=1); only XV_EPI_F32 accepts > 1                        */
  int32_t epilogue;      /* xv_epilogue                                                             */
  int32_t seg_len;       /* row validity for statistics: row m valid iff (m % seg_len) < seg_valid   */
  int32_t seg_valid;     /* (seg_len = 0: all rows < M valid)                                       */
  int32_t accumulate;    /* XV_EPI_BF16 only: out = bf16(acc + out) (gradient fan-in of a shared activation)  */
  void* out;
  int64_t ldc;
  const float* bias;     /* optional [N]                                                            */
  float* col_sum;        /* optional [N]: += sum over valid rows of the STORED bf16 output           */
  float* col_sumsq;      /* optional [N]: += sum over valid rows of its square (needs a TMA-storable out, even N) */
  xv_head_args head;
  xv_bn_bwd_args bn_bwd; /* XV_EPI_BF16 only; needs col_sum (-> dbeta) and col_sumsq (-> dgamma), N % 32 == 0 */
  /* XV_EPI_BF16, inference: out = act(acc * affine_scale[n] + affine_shift[n]) with act(z) = z > 0 ? z : neg_slope * z --
   * batch-norm with moving statistics (tf.layers.batch_normalization, training=False; model/trainer.py:210-225) folded
   * into scale / shift by xv_bn_finalize_infer, plus relu / leaky_relu / identity, in the GEMM epilogue; rows outside the
   * valid frames (seg_len / seg_valid) are written as zeros.  NULL disables.  Needs a TMA-storable output. */
  const float* affine_scale;
  const float* affine_shift;
  float affine_neg_slope;
  int32_t _pad2;
} xv_gemm_args;

XV_API int xv_gemm_bf16(const xv_gemm_args* args, void* stream);
/* Cap the persistent GEMM grid at max_ctas CTAs (0 = all SMs) for subsequently enqueued launches: the data-parallel
 * step reserves SMs for the NCCL kernels of an overlapped gradient all-reduce (process-wide setting, host side only). */
XV_API int xv_gemm_set_cta_limit(int max_ctas);

/* ------------------------------------------------------------------------------------------
 * Frame-level elementwise / reduction kernels (bf16 activations [rows, ld], channels-last).
 * Row validity everywhere: row m (b = m / seg_len, t = m % seg_len) is valid iff
 * t < (lengths ? lengths[b] : seg_valid); seg_len = 0 means every row is valid.
 * Activations: 0 none, 1 relu, 2 leaky_relu(0.2) (tdnn.py:29-30), 3 prelu (common.py:27-42), 4 tanh.
 * ------------------------------------------------------------------------------------------ */
typedef enum { XV_ACT_NONE = 0, XV_ACT_RELU = 1, XV_ACT_LRELU = 2, XV_ACT_PRELU = 3, XV_ACT_TANH = 4 } xv_activation;

/* features f32 [B,T,D] -> bf16 rows [B*T, ldo], out[m, j*dpad + c] = x[b, t+j, c] (zero padded): the K-major
 * A operand of tdnn1_conv (tf.expand_dims + conv2d input side, model/tdnn.py:35-44). */
XV_API int xv_pack_input(const float* x, void* out, int B, int T, int D, int k, int dpad, int64_t ldo, void* stream);

/* tf.layers.batch_normalization (model/tdnn.py:46,64,82,102,121), training mode: per-channel sums produced
 * by the GEMM epilogue -> scale/shift (+ saved mean/rstd for backward, + moving-stat update, eps 1e-3).
 * The sums are those of the STORED pre-BN tensor, which carries no layer bias: scale / shift / save_mean refer to that
 * tensor; `bias` (optional) only enters the moving mean (the statistic of the reference's biased tensor). */
XV_API int xv_bn_finalize_train(const float* col_sum, const float* col_sumsq, const float* bias, float count,
                                const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                                float momentum, float eps, int unbiased_moving_var, float* scale, float* shift,
                                float* save_mean, float* save_rstd, int C, void* stream);
/* xv_bn_finalize_train + xv_bn_act_apply in one launch: every block derives scale / shift from the batch statistics,
 * the first row chunk publishes scale / shift / save_mean / save_rstd and updates the moving statistics. */
XV_API int xv_bn_train_apply(const void* y, void* a, const float* col_sum, const float* col_sumsq, const float* bias,
                             float count, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                             float momentum, float eps, int unbiased_moving_var, float* scale, float* shift,
                             float* save_mean, float* save_rstd, const float* alpha, int act, int64_t rows, int C,
                             int64_t ld, int seg_len, int seg_valid, const int32_t* lengths, void* stream);
/* Per-channel sum and sum of squares of (y - bias) over the valid rows (+= into col_sum / col_sumsq): the batch
 * statistics of tf.layers.batch_normalization for layers whose GEMM is too short to hide a reduction epilogue. */
XV_API int xv_col_stats(const void* y, const float* bias, int64_t rows, int C, int64_t ld, int seg_len, int seg_valid,
                        const int32_t* lengths, float* col_sum, float* col_sumsq, void* stream);
/* ... inference mode (is_training=False, model/trainer.py:210-225): scale/shift from the moving statistics.
 * bias (optional): the layer bias; frame-level pre-BN tensors are stored bias-free (the bias cancels in training-mode
 * BN), so shift = beta - (moving_mean - bias) * scale. */
XV_API int xv_bn_finalize_infer(const float* gamma, const float* beta, const float* moving_mean,
                                const float* moving_var, const float* bias, float eps, float* scale, float* shift, int C,
                                void* stream);
/* a = act(y*scale + shift) on valid rows, 0 on invalid rows  (BN apply + relu, tdnn.py:46-52 etc.). */
XV_API int xv_bn_act_apply(const void* y, void* a, const float* scale, const float* shift, const float* alpha, int act,
                           int64_t rows, int C, int64_t ld, int seg_len, int seg_valid, const int32_t* lengths,
                           void* stream);
/* Backward of act(BN(y)) (what tf.gradients emits for FusedBatchNormGrad/ReluGrad, trainer.py:403):
 * reduce: dgamma += sum g*yhat, dbeta += sum g, dalpha += sum da*min(z,0);  apply: dy (bf16), 0 on invalid rows.
 * Fused tdnn5 path: with pooled/dpooled ([B, 2*pool_cpad] = [mean | std] and its gradient) the upstream gradient
 * da is not read but evaluated on the fly as the statistics-pooling backward of a = act(BN(y)) (da may be NULL). */
XV_API int xv_bn_act_bwd_reduce(const void* y, const void* da, const float* scale, const float* shift,
                                const float* save_mean, const float* save_rstd, const float* alpha, int act,
                                int64_t rows, int C, int64_t ld, int seg_len, int seg_valid, const int32_t* lengths,
                                float* dgamma, float* dbeta, float* dalpha, const float* pooled, const float* dpooled,
                                int pool_cpad, int pool_c_real, void* stream);
XV_API int xv_bn_act_bwd_apply(const void* y, const void* da, void* dy, const float* scale, const float* shift,
                               const float* save_mean, const float* save_rstd, const float* dgamma, const float* dbeta,
                               float count, const float* alpha, int act, int64_t rows, int C, int64_t ld, int seg_len,
                               int seg_valid, const int32_t* lengths, const float* pooled, const float* dpooled,
                               int pool_cpad, int pool_c_real, void* stream);

/* statistics_pooling (model/pooling.py:9-34) and its length-masked form statistics_pooling_v2
 * (model/multitask_v1/pooling.py:9-40): out f32 [B, 2*cpad] = [mean | std]; channels >= c_real read as 0.
 * out_split (optional) = bf16 [B, 6*cpad] = [hi | hi | lo] of out, the A operand of tdnn6_dense.
 * scale != NULL fuses the preceding BN + activation: x is then the pre-BN tensor and a = act(x*scale + shift) is
 * pooled without tdnn5_relu ever being written to HBM.
 * bwd_sums (optional, needs the fused BN and save_mean / save_rstd; not for prelu): f32 [B, 4, cpad] per-(segment,
 * channel) sums S1..S4 of act'(z), act'(z) a, act'(z) yhat, act'(z) a yhat from which xv_pool_bn_bwd_reduce forms the
 * BN dgamma / dbeta of that layer without another pass over the activation. */
XV_API int xv_stats_pool_fwd(const void* x, float* out, void* out_split, int B, int seg_len, int seg_valid,
                             const int32_t* lengths, int c_real, int cpad, int64_t ld, const float* scale,
                             const float* shift, const float* alpha, int act, const float* save_mean,
                             const float* save_rstd, float* bwd_sums, void* stream);
/* Ragged form for extraction (egs/voxceleb/v1/nnet/lib/extract.py:65-94 batched): the utterances of a batch are
 * CONCATENATED in one flat row space -- no padding -- and segment b occupies rows [starts[b], starts[b] + lengths[b]).
 * Frame layers in inference mode are row-local apart from the temporal taps, and a valid output row only ever reads
 * rows of its own utterance, so the rows straddling two utterances are simply never pooled. */
XV_API int xv_stats_pool_ragged(const void* x, float* out, void* out_split, int num_segments, const int32_t* starts,
                                const int32_t* lengths, int c_real, int cpad, int64_t ld, const float* scale,
                                const float* shift, const float* alpha, int act, void* stream);
/* dgamma[c] += sum_b ca S3 + cb S4, dbeta[c] += sum_b ca S1 + cb S2 with (ca, cb) the pooling-gradient coefficients
 * da_t = ca + cb a_t derived from pooled = [mean | std] and dpooled (model/pooling.py:22-32 backward). */
XV_API int xv_pool_bn_bwd_reduce(const float* pooled, const float* dpooled, const float* bwd_sums, int B, int seg_valid,
                                 const int32_t* lengths, int c_real, int cpad, float* dgamma, float* dbeta, void* stream);
XV_API int xv_stats_pool_bwd(const void* x, const float* pooled, const float* dpooled, void* dx, int B, int seg_len,
                             int seg_valid, const int32_t* lengths, int c_real, int cpad, int64_t ld, void* stream);

/* ------------------------------------------------------------------------------------------
 * self_attention pooling (model/pooling.py:37-192).  The key / value nets are frame layers (xv_gemm_bf16 +
 * xv_bn_act_*); these entries replace the einsum / softmax / weighted-moment / penalty ops of pooling.py:147-189
 * and the gradients TF autodiff derives for them.  key / value: bf16 flat-time [B*seg_len, ld]; scores, weights,
 * dweights: f32 [B, H, seg_len]; frames t >= lengths[b] (or seg_valid) get weight 0.
 * qpad f32 [H, ldk] is the query expanded to the key's padded width: qpad[h, d] = query[h, d] (att_split_key false)
 * or qpad[h, h*dq + d] = query[h, d] and 0 elsewhere (att_split_key true, pooling.py:150-153), so that
 * scores[b,h,t] = scale * <key[b,t,:], qpad[h,:]>  (scale = rsqrt(dq) when att_use_scale, pooling.py:155-156).
 * ------------------------------------------------------------------------------------------ */
XV_API int xv_att_expand_query(const float* query, float* qpad, int H, int dq, int ldk, int split_key, void* stream);
XV_API int xv_att_fold_query_grad(const float* dqpad, float* dquery, int H, int dq, int ldk, int split_key, void* stream);
XV_API int xv_att_scores_fwd(const void* key, const float* qpad, float* scores, int B, int seg_len, int seg_valid,
                             const int32_t* lengths, int H, int ldk, float scale, void* stream);
/* weights = softmax over the valid frames (pooling.py:159); may run in place. */
XV_API int xv_att_softmax_fwd(const float* scores, float* weights, int B, int H, int seg_len, int seg_valid,
                              const int32_t* lengths, void* stream);
/* out f32 [B, 2*cpad] = [weighted mean | sqrt(max(weighted var, 1e-12))] (pooling.py:162-170); head h owns channels
 * [h*c_real/H, (h+1)*c_real/H) (split_heads, model/common.py:239-249); out_split as in xv_stats_pool_fwd. */
XV_API int xv_att_pool_fwd(const void* value, const float* weights, float* out, void* out_split, int B, int H,
                           int seg_len, int seg_valid, const int32_t* lengths, int c_real, int cpad, int64_t ld,
                           void* stream);
/* dvalue bf16 [B*seg_len, ld] (= or += when accumulate) and dweights f32 [B, H, seg_len] from dpooled. */
XV_API int xv_att_pool_bwd(const void* value, const float* weights, const float* pooled, const float* dpooled,
                           void* dvalue, float* dweights, int B, int H, int seg_len, int seg_valid,
                           const int32_t* lengths, int c_real, int cpad, int64_t ld, int accumulate, void* stream);
/* penalty[0] += coef/B * sum_b ||W_b W_b^T - I||_F^2 (pooling.py:185-188); gram f32 [B, H, H] = W W^T - I. */
XV_API int xv_att_penalty_fwd(const float* weights, float* gram, float* penalty, int B, int H, int seg_len,
                              int seg_valid, const int32_t* lengths, float coef, void* stream);
/* dweights (in) -> dscores (out, in place): adds the penalty gradient 4 coef/B (gram W) when gram != NULL, applies the
 * softmax Jacobian and the score scale. */
XV_API int xv_att_softmax_bwd(const float* weights, float* dweights, const float* gram, int B, int H, int seg_len,
                              int seg_valid, const int32_t* lengths, float penalty_coef, float scale, void* stream);
/* dkey bf16 [B*seg_len, ldk] (= or +=) = sum_h dscores[b,h,t] qpad[h,:];  dqpad f32 [H, ldk] += sum_m dscores * key. */
XV_API int xv_att_scores_bwd(const void* key, const float* qpad, const float* dscores, void* dkey, float* dqpad, int B,
                             int seg_len, int seg_valid, const int32_t* lengths, int H, int ldk, int accumulate,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * NetVLAD / GhostVLAD pooling (replaces model/pooling.py:195-277 `ghost_vlad`; known answer model/test_utils.py:421-436).
 * K = vlad_num_centers real clusters, KG = K + vlad_num_ghosts <= 64.  logits bf16 [B*seg_len, ldl] is the output of the
 * `vlad_weight_affine` frame layer; post f32 [B, seg_len, KG] = softmax over the clusters (0 for frames >= length);
 * centers f32 [KG, ldc]; res / gres f32 [B, K, cpad]; mass / sumsq / gc f32 [B, K]; out f32 [B, K*cpad] =
 * l2_normalize(res, per cluster) (then over the whole row when final_norm), out_split (optional) bf16 [B, 3*K*cpad] =
 * [hi | hi | lo] of out.  xv_vlad_pool_bwd: dout f32 [B, K*cpad] -> dlogits bf16 [B*seg_len, ldl] (softmax Jacobian
 * applied, padded columns and frames zeroed), dvalue bf16 [B*seg_len, ld] (= or +=), dcenters f32 [KG, ldc] (+=).
 * ------------------------------------------------------------------------------------------ */
XV_API int xv_vlad_post_fwd(const void* logits, float* post, int B, int seg_len, int seg_valid, const int32_t* lengths, int KG,
                            int ldl, void* stream);
XV_API int xv_vlad_pool_fwd(const void* value, const float* post, const float* centers, float* res, float* mass, float* sumsq,
                            float* out, void* out_split, int B, int seg_len, int seg_valid, const int32_t* lengths, int K,
                            int KG, int c_real, int cpad, int64_t ld, int ldc, int final_norm, void* stream);
XV_API int xv_vlad_pool_bwd(const void* value, const float* post, const float* centers, const float* mass, const float* sumsq,
                            const float* out, const float* dout, float* gres, float* gc, void* dlogits, void* dvalue,
                            float* dcenters, int B, int seg_len, int seg_valid, const int32_t* lengths, int K, int KG,
                            int c_real, int cpad, int64_t ld, int ldl, int ldc, int final_norm, int accumulate_dvalue,
                            void* stream);

/* ------------------------------------------------------------------------------------------
 * Metric-learning losses on the [B, E] embeddings (replace model/loss.py:358-498 semihard_triplet_loss, :501-634
 * angular_triplet_loss, :637-705 e2e_valid_loss and model/common.py:61-110 pairwise_euc_distances / pairwise_cos_similarity;
 * known answers model/test_utils.py:21-154, 439-650).  B <= 2048 rows.  All of them are functions of the fp32 Gram matrix:
 *   xv_gram_f32          gram f32 [B, B] = x x^T
 *   xv_semihard_triplet  loss[0] += scale * mean over the positive pairs of max(margin + D_ap - D_an(semi-hard), 0);
 *   xv_angular_triplet   kind 0 = asoftmax (margin 1 / 2 / 4), 1 = additive margin, 2 = additive angular margin; hard = 0:
 *                        mean over the violating triplets, hard = 1: hardest positive / negative per anchor;
 *                        both also write coef f32 [B, B] (symmetric) and diag f32 [B] with
 *   xv_pairwise_bwd      dx f32 [B, ldx] = coef x + diag o x = dLoss/dx.
 *   work: f32 scratch of 2*B*B + 8 floats.  scale folds the replica weight of data-parallel runs.
 *   xv_e2e_valid_loss    forward only; rows are speaker-ordered (num_speakers x num_segments, >= 2 segments), logits =
 *                        20 * cosine to the speaker centres, the own speaker scored against the centre of its OTHER
 *                        segments; loss[0] += scale * mean cross entropy.  work: (B + num_speakers) * E + num_speakers floats.
 * ------------------------------------------------------------------------------------------ */
XV_API int xv_gram_f32(const float* x, float* gram, int B, int E, int64_t ldx, void* stream);
XV_API int xv_semihard_triplet(const float* gram, const int32_t* labels, int B, float margin, int squared, float scale,
                               float* loss, float* coef, float* diag, float* work, void* stream);
XV_API int xv_angular_triplet(const float* gram, const int32_t* labels, int B, int kind, float margin, int hard, float scale,
                              float* loss, float* coef, float* diag, float* work, void* stream);
XV_API int xv_pairwise_bwd(const float* coef, const float* diag, const float* x, float* dx, int B, int E, int64_t ldx,
                           void* stream);
/* Generalized angular triplet loss against class centres (replaces model/loss.py:708-901, loss_compute "raw"; known answer
 * model/test_utils.py:653-852).  cos f32 [B, ldc] = cosine of every sample to every centre (XV_EPI_F32 GEMM of the
 * l2-normalised features with the xv_head_prep_weights operand).  xv_center_triplet: loss[0] += scale * (w_triplet *
 * triplet + w_center * centre); d (optional) bf16 [B, ldc] = dLoss/dcos, the operand of the head's dW / dx GEMMs; topn 1 / k:
 * hardest k non-target centres, 0: all; counters: 4 floats of scratch.  xv_center_between: t f32 [E] = sum of the normalised
 * centres, loss[0] += coef * between-class term; xv_center_between_bwd: dwn += coef * d between / d(normalised centres)
 * (before xv_head_finish_dw).  xv_center_update (triplet_center "average", training): w[:, y_i] -= decay * (w_old[:, y_i] -
 * feats[i, :]) summed over the samples of a class (loss.py:766-783); delta: B*E floats of scratch. */
XV_API int xv_center_triplet(const float* cosm, const int32_t* labels, int B, int C, int64_t ldc, float margin,
                             float target_margin, int topn, float w_triplet, float w_center, float scale, float* loss, void* d,
                             float* counters, void* stream);
XV_API int xv_center_between(const float* w, const float* inv_norm, int E, int C, int64_t ldw, float coef, float* t, float* loss,
                             void* stream);
XV_API int xv_center_between_bwd(float* dwn, const float* w, const float* inv_norm, const float* t, int E, int C, int64_t ldw,
                                 float coef, void* stream);
XV_API int xv_center_update(float* w, const float* feats, const int32_t* labels, float* delta, int B, int E, int64_t ldw,
                            float decay, void* stream);
XV_API int xv_e2e_valid_loss(const float* x, int num_speakers, int num_segments, int E, int64_t ldx, float scale, float* loss,
                             float* work, void* stream);

/* ------------------------------------------------------------------------------------------
 * Utterance-level layers (tdnn6/tdnn7 BN + activation on f32 [B, C], model/tdnn.py:147-189).
 * mode: 0 = no BN (last_layer_no_bn), 1 = training (batch statistics), 2 = inference (moving statistics).
 * a_split: optional bf16 copy of the activation, split_terms = 1 (plain) or 3 ([hi | hi | lo]).
 * ------------------------------------------------------------------------------------------ */
XV_API int xv_bn_rows_fwd(const float* y, int B, int C, int mode, const float* gamma, const float* beta,
                          float* moving_mean, float* moving_var, float momentum, float eps, const float* alpha, int act,
                          float* bn_out, float* a, void* a_split, int split_terms, float* save_mean, float* save_rstd,
                          void* stream);
XV_API int xv_bn_rows_bwd(const float* y, const float* da, int B, int C, int mode, const float* gamma, const float* beta,
                          const float* save_mean, const float* save_rstd, const float* alpha, int act, float* dy,
                          void* dy_bf16, float* dgamma, float* dbeta, float* dalpha, float* dbias, void* stream);
XV_API int xv_cast_split(const float* x, void* out, int64_t rows, int cols, int terms, void* stream);

/* ------------------------------------------------------------------------------------------
 * Margin-softmax head helpers around the XV_EPI_HEAD_* GEMM epilogues (model/loss.py, model/common.py:45-58).
 * ------------------------------------------------------------------------------------------ */
/* w f32 [E,C] -> bf16 [3E, ldw] = [hi; lo; hi] of l2_normalize(w, dim=0) (loss.py:104,213,299); inv_norm[C]. */
XV_API int xv_head_prep_weights(const float* w, void* wn3, float* inv_norm, int E, int C, int64_t ldw, int normalize,
                                void* stream);
/* u f32 [B,E] -> x = l2_scaling(u, scaling) if scaling > 0 (trainer.py:183-186) else u; x3 bf16 [B,3E] = [hi|hi|lo];
 * xnorm[B] = max(||x||, 1e-12) (loss.py:121,220,306); u_rinv[B] = rsqrt(max(||u||^2, 1e-12)). */
XV_API int xv_head_prep_features(const float* u, float scaling, float* x, void* x3, float* xnorm, float* u_rinv, int B,
                                 int E, void* stream);
/* tf.losses.sparse_softmax_cross_entropy (loss.py:35,112,156,244,342): lse from tile partials; loss += mean CE. */
XV_API int xv_head_combine(const float* part_max, const float* part_sum, const float* target_logit, int nblk, int B,
                           float inv_batch, float* lse, float* loss_rows, float* loss, void* stream);
XV_API int xv_head_finish_dx(const float* dx_gemm, const float* gnorm, const float* x, const float* xnorm, const float* u,
                             const float* u_rinv, float scaling, float* du, int B, int E, void* stream);
XV_API int xv_head_finish_dw(float* dw, const float* w, const float* inv_norm, int E, int C, void* stream);
/* Auxiliary losses of the heads (model/loss.py:985-1037, aux_loss_func).
 * Ring loss  lambda * mean_i (||x_i|| - r)^2 on the head's input features (xnorm from xv_head_prep_features, r = the
 * trainable scalar softmax_ringloss/r); scale = lambda / global batch.  loss (optional) += scale * sum_i (n_i - r)^2;
 * gnorm (optional) [B] += 2 scale (n_i - r) -- the dLoss/d||x_i|| input of xv_head_finish_dx -- and dr += -2 scale sum_i (n_i - r). */
XV_API int xv_ring_loss(const float* xnorm, const float* r, int B, float scale, float* loss, float* gnorm, float* dr,
                        void* stream);
/* MHE  lambda / (mean_{i,j} (2 - 2 <wn_{y_i}, wn_j>) + 1e-6), wn = column-normalised speaker matrix (w f32 [E, ldw],
 * inv_norm [C] from xv_head_prep_weights).  forward: t[E] = sum_j wn_j, S[E] = sum_i wn_{y_i}, hist[C] (zero on entry) =
 * label histogram, loss += scale * L, kappa[0] = scale * dL/d<S,t>-coefficient;  backward: dWn[e,j] += kappa (hist_j t_e + S_e)
 * on the un-projected gradient buffer, before xv_head_finish_dw. */
XV_API int xv_mhe_forward(const float* w, const float* inv_norm, const int32_t* labels, int B, int E, int C, int64_t ldw,
                          float lambda, float scale, float* t, float* S, float* hist, float* kappa, float* loss, void* stream);
XV_API int xv_mhe_backward(float* dwn, const float* t, const float* S, const float* hist, const float* kappa, int E, int C,
                           int64_t ldw, void* stream);
/* Class-sharded head (north_star "Data parallelism": the speaker matrix split by columns over the ranks; the reference
 * has no multi-device path, README.md:1,82,115).  Each rank runs the XV_EPI_HEAD_* epilogues on its own columns for
 * the all-gathered rows:
 *   xv_head_local_labels    labels[R] (global class ids) -> out[R] = label - lo inside [0, n_local), else -1 (no target);
 *   xv_head_shard_partials  per-tile partials -> out f32 [3, R] = this shard's per-row (max, sum exp(z'-max), target|0);
 *   xv_head_combine_shards  gathered parts f32 [S, 3, R] -> lse[R] (the all-reduce(max) / all-reduce(sum) pair of a
 *                           sharded softmax evaluated on the gathered pairs) and loss += sum_i (lse_i - target_i)*inv_batch. */
XV_API int xv_head_local_labels(const int32_t* labels, int lo, int n_local, int32_t* out, int R, void* stream);
XV_API int xv_head_shard_partials(const float* part_max, const float* part_sum, const float* target_logit, int nblk, int R,
                                  float* out, void* stream);
XV_API int xv_head_combine_shards(const float* parts, int S, int R, float inv_batch, float* lse, float* loss_rows,
                                  float* loss, void* stream);

/* ------------------------------------------------------------------------------------------
 * Host feeder: dequantise + transpose Kaldi compressed-matrix ('CM ', format 1) segment crops on the device.
 * Replaces the NumPy maps of dataset/kaldi_io.py:780-797 (uint16 percentile -> float, three-piece uint8 -> float) and
 * the column-major -> row-major transpose of :811/:868 that the reference's loader processes run per segment
 * (dataset/data_loader.py:229-307).  Bit-exact with that reader.
 *   data u8 [B, D, ld_t] (column d of segment b: T frames), headers u16 [B, D, 4], glob f32 [B, 2] = (min, range),
 *   out f32 [B, T, ldo].
 * ------------------------------------------------------------------------------------------ */
XV_API int xv_cm_decode(const void* data, const void* headers, const float* glob, float* out, int B, int T, int D,
                        int64_t ld_t, int64_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer over one flat f32 parameter buffer whose tensors start at multiples of 1024 elements
 * (tf.train.GradientDescent/Momentum/AdamOptimizer + l2_regularizer + clip_by_global_norm,
 * model/trainer.py:328-347, 357-358, 403-436).  opt: 0 sgd, 1 momentum, 2 nesterov, 3 adam.
 * hyper (device f32[8]): lr, momentum, beta1, beta2, adam_eps, adam_t, clip_norm (<=0 off), unused.
 * blk_* arrays have one entry per 1024-element block: L2 coefficient, offset of the block's bf16 shadow
 * (-1: none) and, for the 3-term split shadows, the element stride between the hi / lo / hi copies (0: plain).
 * l2_loss_out (optional): += sum l2/2 * w^2 of the PRE-update parameters, i.e. the regularisation loss of this step
 * (tf.losses.get_regularization_loss, trainer.py:357) without a separate pass over the parameters.
 * ------------------------------------------------------------------------------------------ */
XV_API int xv_grad_sumsq(const float* params, const float* grads, const float* blk_l2, int64_t n, float* out, void* stream);
XV_API int xv_opt_step(float* params, const float* grads, float* state1, float* state2, const float* blk_l2,
                       const int64_t* blk_shadow, const int64_t* blk_split_stride, void* shadow, int64_t n, int opt,
                       const float* hyper, const float* gsumsq, float* l2_loss_out, void* stream);
XV_API int xv_shadow_refresh(const float* params, const int64_t* blk_shadow, const int64_t* blk_split_stride,
                             void* shadow, int64_t n, void* stream);
/* Data-parallel gradient all-reduce as one kernel over NVSwitch multicast memory (no reference counterpart: the
 * reference is single-device).  multicast_ptr: multicast (NVLS) mapping of the flat f32 gradient buffers of all ranks
 * (CUDA symmetric memory, same offset on every rank); rank r reduces slice r with multimem.ld_reduce.add.f32 and
 * broadcasts it with multimem.st.  flag_ptrs_dev: device array of `world` pointers to each rank's u32 flag buffer
 * (>= grid * world entries, zero-initialised once); block_epoch: local u32[grid] launch counters (zero-initialised).
 * Reduces floats [offset, offset + n) of the buffers (both multiples of 4).  Every rank must call it the same number of
 * times with the same grid (<= SM count); in place, sum. */
XV_API int xv_dp_allreduce_multimem(void* multicast_ptr, void* const* flag_ptrs_dev, void* block_epoch, int rank, int world,
                                    int64_t offset, int64_t n, int grid, void* stream);
/* Same exchange without multicast: buf_ptrs_dev = device array of `world` pointers to the ranks' symmetric gradient
 * buffers; rank r sums slice r with direct peer loads and stores it to every peer.  Preferred at world = 2. */
XV_API int xv_dp_allreduce_p2p(void* const* buf_ptrs_dev, void* const* flag_ptrs_dev, void* block_epoch, int rank, int world,
                               int64_t offset, int64_t n, int grid, void* stream);
/* Optional bf16 gradient exchange of the data-parallel step: round the flat f32 gradient buffer to bf16 before the
 * all-reduce (half the NVLink bytes) and widen the reduced values again for xv_opt_step.  n % 8 == 0. */
XV_API int xv_grad_pack_bf16(const float* grads, void* out_bf16, int64_t n, void* stream);
XV_API int xv_grad_unpack_bf16(const void* in_bf16, float* grads, int64_t n, void* stream);
XV_API int xv_l2_loss(const float* params, const float* blk_l2, int64_t n, float* out, void* stream);
/* dst[0..n) = host_vals[0..n) (n <= 16), passed as kernel arguments: the learning_rate / global_step placeholders of
 * model/trainer.py:229-231,326 fed per step without a host-buffer race under CUDA-graph replay. */
XV_API int xv_set_scalars(float* dst, const float* host_vals, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XVECTOR_B200_H_ */
