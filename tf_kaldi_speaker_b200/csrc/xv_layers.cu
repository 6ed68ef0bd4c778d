// HBM-bound kernels of the frame-level path: input packing, batch-norm (+activation) forward/backward,
// statistics pooling forward/backward.  Activations are bf16, channels-last, "flat-time" [B*T, ld];
// every access is a 16-byte vector (8 channels) and warps run along the channel axis, so loads/stores
// are fully coalesced.  Reductions over rows/time keep 8 fp32 partials per thread, combine the warps of a
// block through shared memory and issue one atomic per (block, channel).
#include <cuda_bf16.h>

#include "xv_internal.h"

namespace xv {

struct alignas(16) Bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const Bf16x8 r = *reinterpret_cast<const Bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(r.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  Bf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<Bf16x8*>(p) = r;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_PRELU = 3, ACT_TANH = 4 };

__device__ __forceinline__ float act_fwd(int act, float z, float alpha) {
  switch (act) {
    case ACT_RELU: return fmaxf(z, 0.f);
    case ACT_LRELU: return z > 0.f ? z : 0.2f * z;
    case ACT_PRELU: return z > 0.f ? z : alpha * z;   // relu(z) + alpha*(z-|z|)/2  (model/common.py:40-42)
    case ACT_TANH: return tanhf(z);
    default: return z;
  }
}
// d act / d z
__device__ __forceinline__ float act_grad(int act, float z, float alpha) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_LRELU: return z > 0.f ? 1.f : 0.2f;
    case ACT_PRELU: return z > 0.f ? 1.f : alpha;
    case ACT_TANH: { const float t = tanhf(z); return 1.f - t * t; }
    default: return 1.f;
  }
}

__device__ __forceinline__ bool row_is_valid(long long m, int seg_len, int seg_valid, const int* lengths) {
  if (seg_len <= 0) return true;
  const int b = static_cast<int>(m / seg_len), t = static_cast<int>(m % seg_len);
  return t < (lengths ? lengths[b] : seg_valid);
}

// ------------------------------------------------------------------------------------------------
// Input packing: features f32 [B, T, D] -> bf16 im2col rows [B*T, ldo], out[m, j*dpad + c] = x[b, t+j, c].
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int T, int D,
                                  int k, int dpad, long long ldo) {
  const long long total = static_cast<long long>(B) * T * ldo;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / ldo;
    const int col = static_cast<int>(i % ldo);
    const int j = col / dpad, c = col % dpad;
    const int b = static_cast<int>(m / T), t = static_cast<int>(m % T);
    float v = 0.f;
    if (j < k && c < D && t + j < T) v = x[(static_cast<long long>(b) * T + t + j) * D + c];
    out[i] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------
// BN finalize (C threads total).
__global__ void bn_finalize_train_kernel(const float* __restrict__ col_sum, const float* __restrict__ col_sumsq,
                                         const float* __restrict__ bias, float count, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* moving_mean, float* moving_var,
                                         float momentum, float eps, int unbiased, float* scale, float* shift,
                                         float* save_mean, float* save_rstd, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean_nb = col_sum[c] / count;                       // mean of the bias-free accumulator
  const float var = fmaxf(col_sumsq[c] / count - mean_nb * mean_nb, 0.f);
  const float mean = mean_nb + (bias ? bias[c] : 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  save_mean[c] = mean;
  save_rstd[c] = rstd;
  if (moving_mean) {
    const float mv = unbiased ? var * (count / fmaxf(count - 1.f, 1.f)) : var;
    moving_mean[c] = moving_mean[c] * momentum + mean * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + mv * (1.f - momentum);
  }
}
__global__ void bn_finalize_infer_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                         const float* __restrict__ mm, const float* __restrict__ mv, float eps,
                                         float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(mv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - mm[c] * sc;
}

// a = act(y*scale + shift) on valid rows, 0 on invalid rows.
__global__ void bn_act_apply_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ a,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ alpha, int act, long long rows, int C, long long ld,
                                    int seg_len, int seg_valid, const int* __restrict__ lengths) {
  const int cv = C / 8;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / cv;
    const int c0 = static_cast<int>(i % cv) * 8;
    float o[8];
    if (row_is_valid(m, seg_len, seg_valid, lengths)) {
      float v[8], sc[8], sh[8];
      load8(y + m * ld + c0, v);
      load8f(scale + c0, sc);
      load8f(shift + c0, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = act_fwd(act, fmaf(v[j], sc[j], sh[j]), act == ACT_PRELU ? alpha[c0 + j] : 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    store8(a + m * ld + c0, o);
  }
}

// Column reductions for the BN backward: dbeta += sum g, dgamma += sum g*yhat, dalpha += sum da*min(z,0).
// grid = (C/256, row chunks); block = 256 threads = 8 warps; warp w takes rows w, w+8, ... of the chunk.
constexpr int RED_ROWS_PER_BLOCK = 128;
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ da, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
    const float* __restrict__ alpha, int act, long long rows, int C, long long ld, int seg_len, int seg_valid,
    const int* __restrict__ lengths, float* dgamma, float* dbeta, float* dalpha) {
  __shared__ float red[8][3][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const bool c_ok = c0 < C;
  float sg[8], sgy[8], sal[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sg[j] = sgy[j] = sal[j] = 0.f;
  if (c_ok) {
    float sc[8], sh[8], mu[8], rs[8], al[8];
    load8f(scale + c0, sc); load8f(shift + c0, sh); load8f(save_mean + c0, mu); load8f(save_rstd + c0, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) al[j] = (act == ACT_PRELU) ? alpha[c0 + j] : 0.f;
    const long long r0 = static_cast<long long>(blockIdx.y) * RED_ROWS_PER_BLOCK;
    const long long r1 = min(r0 + RED_ROWS_PER_BLOCK, rows);
    for (long long m = r0 + w; m < r1; m += 8) {
      if (!row_is_valid(m, seg_len, seg_valid, lengths)) continue;
      float v[8], d[8];
      load8(y + m * ld + c0, v);
      load8(da + m * ld + c0, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(v[j], sc[j], sh[j]);
        const float g = d[j] * act_grad(act, z, al[j]);
        sg[j] += g;
        sgy[j] += g * (v[j] - mu[j]) * rs[j];
        if (act == ACT_PRELU) sal[j] += d[j] * fminf(z, 0.f);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[w][0][lane * 8 + j] = sg[j];
    red[w][1][lane * 8 + j] = sgy[j];
    red[w][2][lane * 8 + j] = sal[j];
  }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < C) {
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[k][0][threadIdx.x]; b += red[k][1][threadIdx.x]; d += red[k][2][threadIdx.x]; }
    atomicAdd(dbeta + c, a);
    atomicAdd(dgamma + c, b);
    if (act == ACT_PRELU && dalpha) atomicAdd(dalpha + c, d);
  }
}

// dy = scale * (g - dbeta/n - yhat*dgamma/n) on valid rows, 0 elsewhere (scale = gamma*rstd).
__global__ void bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ da,
                                        __nv_bfloat16* __restrict__ dy, const float* __restrict__ scale,
                                        const float* __restrict__ shift, const float* __restrict__ save_mean,
                                        const float* __restrict__ save_rstd, const float* __restrict__ dgamma,
                                        const float* __restrict__ dbeta, float inv_count,
                                        const float* __restrict__ alpha, int act, long long rows, int C, long long ld,
                                        int seg_len, int seg_valid, const int* __restrict__ lengths) {
  const int cv = C / 8;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / cv;
    const int c0 = static_cast<int>(i % cv) * 8;
    float o[8];
    if (row_is_valid(m, seg_len, seg_valid, lengths)) {
      float v[8], d[8], sc[8], sh[8], mu[8], rs[8], dg[8], db[8];
      load8(y + m * ld + c0, v);
      load8(da + m * ld + c0, d);
      load8f(scale + c0, sc); load8f(shift + c0, sh); load8f(save_mean + c0, mu); load8f(save_rstd + c0, rs);
      load8f(dgamma + c0, dg); load8f(dbeta + c0, db);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(v[j], sc[j], sh[j]);
        const float g = d[j] * act_grad(act, z, act == ACT_PRELU ? alpha[c0 + j] : 0.f);
        const float yh = (v[j] - mu[j]) * rs[j];
        o[j] = sc[j] * (g - db[j] * inv_count - yh * dg[j] * inv_count);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    store8(dy + m * ld + c0, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Statistics pooling (model/pooling.py:22-32; masked form multitask_v1/pooling.py:22-38).
// grid = (Cpad/256, B); block = 256 = 8 warps striding over time; shifted one-pass moments
// (shift = first frame) so that var = E[(x-x0)^2] - E[x-x0]^2 does not cancel catastrophically.
__global__ void __launch_bounds__(256) stats_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                             __nv_bfloat16* __restrict__ out3, int seg_len,
                                                             int seg_valid, const int* __restrict__ lengths,
                                                             int c_real, int cpad, long long ld) {
  __shared__ float red[8][2][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int L = lengths ? lengths[b] : seg_valid;
  const __nv_bfloat16* xb = x + static_cast<long long>(b) * seg_len * ld;
  float s1[8], s2[8], x0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = x0[j] = 0.f;
  if (c0 < cpad && L > 0) {
    load8(xb + c0, x0);
    for (int t = w; t < L; t += 8) {
      float v[8];
      load8(xb + static_cast<long long>(t) * ld + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[j] - x0[j];
        s1[j] += d;
        s2[j] += d * d;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[w][0][lane * 8 + j] = s1[j]; red[w][1][lane * 8 + j] = s2[j]; }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cpad) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[k][0][threadIdx.x]; q += red[k][1][threadIdx.x]; }
    float mean = 0.f, sd = 0.f;
    if (c < c_real && L > 0) {
      const float first = __bfloat162float(xb[c]);
      const float invl = 1.0f / (static_cast<float>(L) + 1e-16f);
      const float md = a * invl;
      mean = first + md;
      float var = q * invl - md * md;
      var = (var <= 1e-12f) ? 1e-12f : var;       // VAR2STD_EPSILON floor (mask blend, pooling.py:28-29)
      sd = sqrtf(var);
    }
    float* ob = out + static_cast<long long>(b) * 2 * cpad;
    ob[c] = mean;
    ob[cpad + c] = sd;
    if (out3) {   // [hi | hi | lo] split copy: operand of the tdnn6 GEMM (K = 3 * 2*cpad)
      __nv_bfloat16* o3 = out3 + static_cast<long long>(b) * 6 * cpad;
      const __nv_bfloat16 mh = __float2bfloat16(mean), sh = __float2bfloat16(sd);
      o3[c] = mh; o3[cpad + c] = sh;
      o3[2 * cpad + c] = mh; o3[3 * cpad + c] = sh;
      o3[4 * cpad + c] = __float2bfloat16(mean - __bfloat162float(mh));
      o3[5 * cpad + c] = __float2bfloat16(sd - __bfloat162float(sh));
    }
  }
}

// dx_t = gmean/L + 1[var>floor] * gstd * (x_t - mean) / (L * std) on valid frames, 0 elsewhere.
__global__ void stats_pool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pooled,
                                      const float* __restrict__ dpooled, __nv_bfloat16* __restrict__ dx, int B,
                                      int seg_len, int seg_valid, const int* __restrict__ lengths, int c_real, int cpad,
                                      long long ld) {
  const int cv = cpad / 8;
  const long long total = static_cast<long long>(B) * seg_len * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / cv;
    const int c0 = static_cast<int>(i % cv) * 8;
    const int b = static_cast<int>(m / seg_len), t = static_cast<int>(m % seg_len);
    const int L = lengths ? lengths[b] : seg_valid;
    float o[8];
    if (t < L) {
      float v[8], mu[8], sd[8], gm[8], gs[8];
      load8(x + m * ld + c0, v);
      const float* pb = pooled + static_cast<long long>(b) * 2 * cpad;
      const float* gb = dpooled + static_cast<long long>(b) * 2 * cpad;
      load8f(pb + c0, mu); load8f(pb + cpad + c0, sd); load8f(gb + c0, gm); load8f(gb + cpad + c0, gs);
      const float invl = 1.0f / (static_cast<float>(L) + 1e-16f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float g = 0.f;
        if (c0 + j < c_real) {
          g = gm[j] * invl;
          if (sd[j] * sd[j] > 1.0000001e-12f) g += gs[j] * (v[j] - mu[j]) * invl / sd[j];
        }
        o[j] = g;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    store8(dx + m * ld + c0, o);
  }
}

static inline int grid_for(long long work_items, int block, int sms) {
  long long g = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(sms) * 16;   // grid-stride beyond 16 resident-ish blocks per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace xv

using namespace xv;

extern "C" int xv_pack_input(const float* x, void* out, int B, int T, int D, int k, int dpad, int64_t ldo, void* stream) {
  if (!x || !out || B <= 0 || T <= 0 || D <= 0 || D > dpad || k * dpad > ldo) return set_error(XV_ERR_INVALID, "xv_pack_input: bad arguments");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  const long long total = static_cast<long long>(B) * T * ldo;
  pack_input_kernel<<<grid_for(total, 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(out), B, T, D, k, dpad, ldo);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_finalize_train(const float* col_sum, const float* col_sumsq, const float* bias, float count,
                                    const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                                    float momentum, float eps, int unbiased, float* scale, float* shift,
                                    float* save_mean, float* save_rstd, int C, void* stream) {
  if (!col_sum || !col_sumsq || !gamma || !beta || !scale || !shift || !save_mean || !save_rstd || C <= 0 || count <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_finalize_train: bad arguments");
  bn_finalize_train_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      col_sum, col_sumsq, bias, count, gamma, beta, moving_mean, moving_var, momentum, eps, unbiased, scale, shift,
      save_mean, save_rstd, C);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_finalize_infer(const float* gamma, const float* beta, const float* moving_mean,
                                    const float* moving_var, float eps, float* scale, float* shift, int C, void* stream) {
  if (!gamma || !beta || !moving_mean || !moving_var || !scale || !shift || C <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_finalize_infer: bad arguments");
  bn_finalize_infer_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(gamma, beta, moving_mean,
                                                                                            moving_var, eps, scale, shift, C);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

static int check_act_layout(const char* who, int C, int64_t ld, int act, const float* alpha) {
  if (C <= 0 || C % 8 || ld % 8 || ld < C) return set_error(XV_ERR_INVALID, "%s: C and ld must be multiples of 8, ld >= C", who);
  if (act < ACT_NONE || act > ACT_TANH) return set_error(XV_ERR_INVALID, "%s: unknown activation %d", who, act);
  if (act == ACT_PRELU && !alpha) return set_error(XV_ERR_INVALID, "%s: prelu needs alpha", who);
  return XV_OK;
}

extern "C" int xv_bn_act_apply(const void* y, void* a, const float* scale, const float* shift, const float* alpha,
                               int act, int64_t rows, int C, int64_t ld, int seg_len, int seg_valid,
                               const int32_t* lengths, void* stream) {
  if (!y || !a || !scale || !shift || rows <= 0) return set_error(XV_ERR_INVALID, "xv_bn_act_apply: bad arguments");
  int rc = check_act_layout("xv_bn_act_apply", C, ld, act, alpha); if (rc) return rc;
  int sms; rc = device_sm_count(&sms); if (rc) return rc;
  bn_act_apply_kernel<<<grid_for(rows * (C / 8), 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(a), scale, shift, alpha, act, rows, C, ld,
      seg_len, seg_valid, lengths);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_act_bwd_reduce(const void* y, const void* da, const float* scale, const float* shift,
                                    const float* save_mean, const float* save_rstd, const float* alpha, int act,
                                    int64_t rows, int C, int64_t ld, int seg_len, int seg_valid, const int32_t* lengths,
                                    float* dgamma, float* dbeta, float* dalpha, void* stream) {
  if (!y || !da || !scale || !shift || !save_mean || !save_rstd || !dgamma || !dbeta || rows <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_act_bwd_reduce: bad arguments");
  int rc = check_act_layout("xv_bn_act_bwd_reduce", C, ld, act, alpha); if (rc) return rc;
  dim3 grid(ceil_div(C, 256), ceil_div(rows, RED_ROWS_PER_BLOCK));
  bn_act_bwd_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(da), scale, shift, save_mean, save_rstd,
      alpha, act, rows, C, ld, seg_len, seg_valid, lengths, dgamma, dbeta, dalpha);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_act_bwd_apply(const void* y, const void* da, void* dy, const float* scale, const float* shift,
                                   const float* save_mean, const float* save_rstd, const float* dgamma,
                                   const float* dbeta, float count, const float* alpha, int act, int64_t rows, int C,
                                   int64_t ld, int seg_len, int seg_valid, const int32_t* lengths, void* stream) {
  if (!y || !da || !dy || !scale || !shift || !save_mean || !save_rstd || !dgamma || !dbeta || rows <= 0 || count <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_act_bwd_apply: bad arguments");
  int rc = check_act_layout("xv_bn_act_bwd_apply", C, ld, act, alpha); if (rc) return rc;
  int sms; rc = device_sm_count(&sms); if (rc) return rc;
  bn_act_bwd_apply_kernel<<<grid_for(rows * (C / 8), 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(da), static_cast<__nv_bfloat16*>(dy),
      scale, shift, save_mean, save_rstd, dgamma, dbeta, 1.0f / count, alpha, act, rows, C, ld, seg_len, seg_valid, lengths);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_stats_pool_fwd(const void* x, float* out, void* out_split, int B, int seg_len, int seg_valid,
                                 const int32_t* lengths, int c_real, int cpad, int64_t ld, void* stream) {
  if (!x || !out || B <= 0 || seg_len <= 0 || cpad % 8 || c_real > cpad || ld % 8 || ld < cpad)
    return set_error(XV_ERR_INVALID, "xv_stats_pool_fwd: bad arguments");
  dim3 grid(ceil_div(cpad, 256), B);
  stats_pool_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), out, static_cast<__nv_bfloat16*>(out_split), seg_len, seg_valid, lengths,
      c_real, cpad, ld);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_stats_pool_bwd(const void* x, const float* pooled, const float* dpooled, void* dx, int B, int seg_len,
                                 int seg_valid, const int32_t* lengths, int c_real, int cpad, int64_t ld, void* stream) {
  if (!x || !pooled || !dpooled || !dx || B <= 0 || seg_len <= 0 || cpad % 8 || c_real > cpad || ld % 8 || ld < cpad)
    return set_error(XV_ERR_INVALID, "xv_stats_pool_bwd: bad arguments");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  const long long total = static_cast<long long>(B) * seg_len * (cpad / 8);
  stats_pool_bwd_kernel<<<grid_for(total, 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), pooled, dpooled, static_cast<__nv_bfloat16*>(dx), B, seg_len, seg_valid,
      lengths, c_real, cpad, ld);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
