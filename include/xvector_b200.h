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
  XV_EPI_BF16 = 0,      /* out bf16 [M, ldc] = acc (+ bias[n]); optional per-column sum / sum-of-squares   */
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
  float fa, fs;          /* 1/(1+lambda), 1-fa (lambda schedule evaluated by the host, loss.py:144) */
  const int32_t* labels; /* [M]                                                                     */
  const float* xnorm;    /* [M] max(||x_i||, 1e-12)                                                 */
  /* forward outputs */
  float* part_max;       /* [num_n_blocks, M]                                                       */
  float* part_sum;       /* [num_n_blocks, M]                                                       */
  float* target_logit;   /* [M] modified target logit z'_{i,y_i}                                    */
  float* logits_out;     /* optional [M, ldc] f32 pre-margin logits (endpoints["logits"]); may be 0 */
  /* backward inputs / outputs */
  const float* lse;      /* [M] log-sum-exp of the modified logits                                  */
  float inv_batch;       /* 1 / global batch                                                        */
  float* gnorm;          /* [M] dLoss/d||x_i|| through the margin term                              */
} xv_head_args;

typedef struct {
  xv_operand a, b;
  int32_t M, N, K;
  int32_t splits;        /* split-K factor (>=1); only XV_EPI_F32 accepts > 1                        */
  int32_t epilogue;      /* xv_epilogue                                                             */
  int32_t seg_len;       /* row validity for statistics: row m valid iff (m % seg_len) < seg_valid   */
  int32_t seg_valid;     /* (seg_len = 0: all rows < M valid)                                       */
  int32_t _pad;
  void* out;
  int64_t ldc;
  const float* bias;     /* optional [N]                                                            */
  float* col_sum;        /* optional [N]: += sum over valid rows of acc (bias excluded)             */
  float* col_sumsq;      /* optional [N]: += sum over valid rows of acc^2                           */
  xv_head_args head;
} xv_gemm_args;

XV_API int xv_gemm_bf16(const xv_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XVECTOR_B200_H_ */
