// Utterance-level (post-pooling) batch-norm + activation on small fp32 [B, C] matrices (tdnn6 / tdnn7,
// model/tdnn.py:147-189) and the attention key nets.  These tensors are a few hundred KB: the kernels are
// latency-bound, one thread per channel with loads coalesced across channels.
#include <cuda_bf16.h>

#include "xv_internal.h"

namespace xv {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_PRELU = 3, ACT_TANH = 4 };

__device__ __forceinline__ float uact_fwd(int act, float z, float alpha) {
  switch (act) {
    case ACT_RELU: return fmaxf(z, 0.f);
    case ACT_LRELU: return z > 0.f ? z : 0.2f * z;
    case ACT_PRELU: return z > 0.f ? z : alpha * z;
    case ACT_TANH: return tanhf(z);
    default: return z;
  }
}
__device__ __forceinline__ float uact_grad(int act, float z, float alpha) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_LRELU: return z > 0.f ? 1.f : 0.2f;
    case ACT_PRELU: return z > 0.f ? 1.f : alpha;
    case ACT_TANH: { const float t = tanhf(z); return 1.f - t * t; }
    default: return 1.f;
  }
}

// Block = 32 channels x UROWG row groups (1024 threads); sums over the B rows are combined through shared memory.
// (Latency-bound: with 8 row groups a thread walked 16 dependent rows per pass for B = 128.)
constexpr int UROWG = 32;
__device__ __forceinline__ float rows_block_sum(float v, float (*red)[32], int tx, int ty) {
  __syncthreads();
  red[ty][tx] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < UROWG; ++k) t += red[k][tx];
  return t;
}

// mode: 0 = no BN (identity), 1 = training (batch statistics over the B rows), 2 = inference (moving stats)
__global__ void __launch_bounds__(32 * UROWG) bn_rows_fwd_kernel(const float* __restrict__ y, int B, int C, int mode,
                                   const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* moving_mean, float* moving_var, float momentum,
                                   float eps, const float* __restrict__ alpha, int act, float* __restrict__ bn_out,
                                   float* __restrict__ a, __nv_bfloat16* __restrict__ a_split, int split_terms,
                                   float* save_mean, float* save_rstd) {
  pdl_entry();
  __shared__ float red[UROWG][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const bool ok = c < C;
  float mean = 0.f, rstd = 1.f, g = 1.f, bt = 0.f;
  if (mode == 1) {
    float s = 0.f;
    if (ok) for (int i = ty; i < B; i += UROWG) s += y[static_cast<long long>(i) * C + c];
    mean = rows_block_sum(s, red, tx, ty) / B;
    float q = 0.f;
    if (ok) for (int i = ty; i < B; i += UROWG) { const float d = y[static_cast<long long>(i) * C + c] - mean; q += d * d; }
    const float var = rows_block_sum(q, red, tx, ty) / B;
    rstd = rsqrtf(var + eps);
    if (ok && ty == 0 && moving_mean) {   // rank-2 tensors take TF's unfused path: biased variance in the moving average
      moving_mean[c] = moving_mean[c] * momentum + mean * (1.f - momentum);
      moving_var[c] = moving_var[c] * momentum + var * (1.f - momentum);
    }
  } else if (mode == 2 && ok) {
    mean = moving_mean[c];
    rstd = rsqrtf(moving_var[c] + eps);
  }
  if (!ok) return;
  if (mode != 0) { g = gamma[c]; bt = beta[c]; }
  if (save_mean && ty == 0) { save_mean[c] = mean; save_rstd[c] = rstd; }
  const float al = (act == ACT_PRELU) ? alpha[c] : 0.f;
  for (int i = ty; i < B; i += UROWG) {
    const long long idx = static_cast<long long>(i) * C + c;
    const float z = (mode == 0) ? y[idx] : ((y[idx] - mean) * rstd * g + bt);
    if (bn_out) bn_out[idx] = z;
    const float v = uact_fwd(act, z, al);
    if (a) a[idx] = v;
    if (a_split) {
      __nv_bfloat16* row = a_split + static_cast<long long>(i) * split_terms * C;
      const __nv_bfloat16 h = __float2bfloat16(v);
      row[c] = h;
      if (split_terms == 3) {   // [hi | hi | lo]
        row[C + c] = h;
        row[2 * C + c] = __float2bfloat16(v - __bfloat162float(h));
      }
    }
  }
}

// Backward of act(BN(y)) over B rows.  dy in fp32 and bf16 (the latter feeds the dgrad/wgrad GEMMs).
__global__ void __launch_bounds__(32 * UROWG) bn_rows_bwd_kernel(const float* __restrict__ y, const float* __restrict__ da, int B,
                                   int C, int mode,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                   const float* __restrict__ alpha, int act, float* __restrict__ dy,
                                   __nv_bfloat16* __restrict__ dy_bf16, float* dgamma, float* dbeta, float* dalpha,
                                   float* dbias) {
  pdl_entry();
  __shared__ float red[UROWG][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const bool ok = c < C;
  const float mean = (mode != 0 && ok) ? save_mean[c] : 0.f, rstd = (mode != 0 && ok) ? save_rstd[c] : 1.f;
  const float g = (mode != 0 && ok) ? gamma[c] : 1.f, bt = (mode != 0 && ok) ? beta[c] : 0.f;
  const float al = (act == ACT_PRELU && ok) ? alpha[c] : 0.f;
  float sg = 0.f, sgy = 0.f, sal = 0.f;
  if (ok) {
    for (int i = ty; i < B; i += UROWG) {
      const long long idx = static_cast<long long>(i) * C + c;
      const float yh = (y[idx] - mean) * rstd;
      const float z = (mode == 0) ? y[idx] : (yh * g + bt);
      const float gr = da[idx] * uact_grad(act, z, al);
      sg += gr;
      sgy += gr * yh;
      if (act == ACT_PRELU) sal += da[idx] * fminf(z, 0.f);
    }
  }
  sg = rows_block_sum(sg, red, tx, ty);
  sgy = rows_block_sum(sgy, red, tx, ty);
  if (act == ACT_PRELU) sal = rows_block_sum(sal, red, tx, ty);
  if (ok && ty == 0) {
    if (mode != 0) {
      if (dgamma) atomicAdd(dgamma + c, sgy);
      if (dbeta) atomicAdd(dbeta + c, sg);
    }
    if (act == ACT_PRELU && dalpha) atomicAdd(dalpha + c, sal);
  }
  float bias_grad = 0.f;
  if (ok) {
    for (int i = ty; i < B; i += UROWG) {
      const long long idx = static_cast<long long>(i) * C + c;
      const float yh = (y[idx] - mean) * rstd;
      const float z = (mode == 0) ? y[idx] : (yh * g + bt);
      const float gr = da[idx] * uact_grad(act, z, al);
      float d;
      if (mode == 1) d = g * rstd * (gr - sg / B - yh * sgy / B);
      else if (mode == 2) d = g * rstd * gr;
      else d = gr;
      bias_grad += d;
      if (dy) dy[idx] = d;
      if (dy_bf16) dy_bf16[idx] = __float2bfloat16(d);
    }
  }
  bias_grad = rows_block_sum(bias_grad, red, tx, ty);
  if (ok && ty == 0 && dbias) atomicAdd(dbias + c, bias_grad);   // bias of the affine layer feeding this BN
}

// f32 [rows, cols] -> bf16, optionally as the [hi | hi | lo] split (terms = 3).
__global__ void cast_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int cols,
                                  int terms) {
  pdl_entry();
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i % cols);
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16(v);
    __nv_bfloat16* row = out + r * terms * cols;
    row[c] = h;
    if (terms == 3) {
      row[cols + c] = h;
      row[2 * cols + c] = __float2bfloat16(v - __bfloat162float(h));
    }
  }
}

}  // namespace xv

using namespace xv;

extern "C" int xv_bn_rows_fwd(const float* y, int B, int C, int mode, const float* gamma, const float* beta,
                              float* moving_mean, float* moving_var, float momentum, float eps, const float* alpha,
                              int act, float* bn_out, float* a, void* a_split, int split_terms, float* save_mean,
                              float* save_rstd, void* stream) {
  if (!y || B <= 0 || C <= 0 || mode < 0 || mode > 2) return set_error(XV_ERR_INVALID, "xv_bn_rows_fwd: bad arguments");
  if (mode != 0 && (!gamma || !beta)) return set_error(XV_ERR_INVALID, "xv_bn_rows_fwd: BN needs gamma/beta");
  if (mode == 2 && (!moving_mean || !moving_var)) return set_error(XV_ERR_INVALID, "xv_bn_rows_fwd: inference needs moving stats");
  if (act == ACT_PRELU && !alpha) return set_error(XV_ERR_INVALID, "xv_bn_rows_fwd: prelu needs alpha");
  if (a_split && split_terms != 1 && split_terms != 3) return set_error(XV_ERR_INVALID, "xv_bn_rows_fwd: split_terms must be 1 or 3");
  ::xv::launch_pdl((bn_rows_fwd_kernel), ceil_div(C, 32), 32 * UROWG, 0, static_cast<cudaStream_t>(stream), 
      y, B, C, mode, gamma, beta, moving_mean, moving_var, momentum, eps, alpha, act, bn_out, a,
      static_cast<__nv_bfloat16*>(a_split), split_terms, save_mean, save_rstd);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_rows_bwd(const float* y, const float* da, int B, int C, int mode, const float* gamma,
                              const float* beta, const float* save_mean, const float* save_rstd, const float* alpha,
                              int act, float* dy, void* dy_bf16, float* dgamma, float* dbeta, float* dalpha,
                              float* dbias, void* stream) {
  if (!y || !da || B <= 0 || C <= 0 || mode < 0 || mode > 2) return set_error(XV_ERR_INVALID, "xv_bn_rows_bwd: bad arguments");
  if (mode != 0 && (!gamma || !beta || !save_mean || !save_rstd)) return set_error(XV_ERR_INVALID, "xv_bn_rows_bwd: BN needs saved statistics");
  ::xv::launch_pdl((bn_rows_bwd_kernel), ceil_div(C, 32), 32 * UROWG, 0, static_cast<cudaStream_t>(stream), 
      y, da, B, C, mode, gamma, beta, save_mean, save_rstd, alpha, act, dy, static_cast<__nv_bfloat16*>(dy_bf16),
      dgamma, dbeta, dalpha, dbias);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_cast_split(const float* x, void* out, int64_t rows, int cols, int terms, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || (terms != 1 && terms != 3)) return set_error(XV_ERR_INVALID, "xv_cast_split: bad arguments");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  long long g = (rows * cols + 255) / 256;
  if (g > sms * 8LL) g = sms * 8LL;
  ::xv::launch_pdl((cast_split_kernel), static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream), 
      x, static_cast<__nv_bfloat16*>(out), rows, cols, terms);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
