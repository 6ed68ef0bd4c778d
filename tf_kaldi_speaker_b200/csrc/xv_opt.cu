// Multi-tensor optimizer step over ONE flat fp32 parameter buffer (model/trainer.py:328-347, 403-436):
// L2 regulariser gradient (kernels only), clip_by_global_norm, SGD / Momentum / Nesterov / Adam, and the bf16
// "shadow" copies the tensor-core GEMMs read (plain, or the [hi; lo; hi] 3-term split of the utterance layers).
// Layout contract: every tensor starts at a multiple of XV_OPT_BLOCK (1024) elements in the flat buffer, so a
// 1024-element block never straddles two tensors and the per-tensor attributes are one table lookup per block.
#include <cuda_bf16.h>

#include "xv_internal.h"

namespace xv {

constexpr int OPT_BLOCK = 1024;
enum { OPT_SGD = 0, OPT_MOMENTUM = 1, OPT_NESTEROV = 2, OPT_ADAM = 3 };

// sum over all elements of (g + l2*w)^2  -> out[0] (atomic); blk_l2[b] is the L2 coefficient of block b.
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ params, const float* __restrict__ grads,
                                                         const float* __restrict__ blk_l2, long long n, float* out) {
  pdl_entry();
  __shared__ float sh[8];
  float acc = 0.f;
  for (long long blk = blockIdx.x; blk * OPT_BLOCK < n; blk += gridDim.x) {
    const float l2 = blk_l2[blk];
    const long long base = blk * OPT_BLOCK + threadIdx.x * 4;
    if (base + 3 < n) {
      const float4 g = *reinterpret_cast<const float4*>(grads + base);
      const float4 w = *reinterpret_cast<const float4*>(params + base);
      const float a = g.x + l2 * w.x, b = g.y + l2 * w.y, c = g.z + l2 * w.z, d = g.w + l2 * w.w;
      acc += a * a + b * b + c * c + d * d;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

// hyper (device): [0] lr, [1] momentum, [2] beta1, [3] beta2, [4] adam eps, [5] adam step t (>=1), [6] clip norm (<=0: off)
// Each CTA walks OPT_UNROLL 1024-element blocks per iteration with all their loads issued up front (one block per
// iteration left ~2 x 16 B in flight per thread: 40 % of HBM).  l2_out (optional) += sum l2/2 * w_old^2, the
// regularisation loss of the step (tf.losses.get_regularization_loss, trainer.py:357), so no separate pass reads w.
constexpr int OPT_UNROLL = 4;
__global__ void __launch_bounds__(256) opt_step_kernel(float* __restrict__ params, const float* __restrict__ grads,
                                                       float* __restrict__ s1, float* __restrict__ s2,
                                                       const float* __restrict__ blk_l2,
                                                       const long long* __restrict__ blk_shadow,
                                                       const long long* __restrict__ blk_split_stride,
                                                       __nv_bfloat16* __restrict__ shadow, long long n, int opt,
                                                       const float* __restrict__ hyper, const float* __restrict__ gsumsq,
                                                       float* l2_out) {
  pdl_entry();
  __shared__ float sh[8];
  const float lr = hyper[0];
  float gscale = 1.f;
  if (hyper[6] > 0.f && gsumsq) {
    const float gn = sqrtf(*gsumsq);
    gscale = hyper[6] / fmaxf(gn, hyper[6]);      // tf.clip_by_global_norm
  }
  float lr_t = lr;
  if (opt == OPT_ADAM) lr_t = lr * sqrtf(1.f - powf(hyper[3], hyper[5])) / (1.f - powf(hyper[2], hyper[5]));
  const long long nblk = (n + OPT_BLOCK - 1) / OPT_BLOCK;
  float l2acc = 0.f;
  for (long long blk0 = static_cast<long long>(blockIdx.x) * OPT_UNROLL; blk0 < nblk;
       blk0 += static_cast<long long>(gridDim.x) * OPT_UNROLL) {
    float4 g4[OPT_UNROLL], w4[OPT_UNROLL], a4[OPT_UNROLL], b4[OPT_UNROLL];
    float l2[OPT_UNROLL];
    bool ok[OPT_UNROLL];
#pragma unroll
    for (int u = 0; u < OPT_UNROLL; ++u) {
      const long long blk = blk0 + u;
      const long long base = blk * OPT_BLOCK + threadIdx.x * 4;
      ok[u] = blk < nblk && base + 3 < n;
      a4[u] = b4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[u]) {
        l2[u] = blk_l2[blk];
        g4[u] = *reinterpret_cast<const float4*>(grads + base);
        w4[u] = *reinterpret_cast<const float4*>(params + base);
        if (opt != OPT_SGD) a4[u] = *reinterpret_cast<const float4*>(s1 + base);
        if (opt == OPT_ADAM) b4[u] = *reinterpret_cast<const float4*>(s2 + base);
      }
    }
#pragma unroll
    for (int u = 0; u < OPT_UNROLL; ++u) {
      if (!ok[u]) continue;
      const long long blk = blk0 + u;
      const long long base = blk * OPT_BLOCK + threadIdx.x * 4;
      float g[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
      float w[4] = {w4[u].x, w4[u].y, w4[u].z, w4[u].w};
      float a[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
      float b[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w};
      l2acc += 0.5f * l2[u] * (w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + w[3] * w[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gr = (g[j] + l2[u] * w[j]) * gscale;
        if (opt == OPT_SGD) {
          w[j] -= lr * gr;
        } else if (opt == OPT_MOMENTUM || opt == OPT_NESTEROV) {
          a[j] = a[j] * hyper[1] + gr;                                 // accum = momentum*accum + grad
          w[j] -= lr * ((opt == OPT_NESTEROV) ? (gr + hyper[1] * a[j]) : a[j]);
        } else {
          a[j] = a[j] * hyper[2] + (1.f - hyper[2]) * gr;
          b[j] = b[j] * hyper[3] + (1.f - hyper[3]) * gr * gr;
          w[j] -= lr_t * a[j] / (sqrtf(b[j]) + hyper[4]);
        }
      }
      *reinterpret_cast<float4*>(params + base) = make_float4(w[0], w[1], w[2], w[3]);
      if (opt != OPT_SGD) *reinterpret_cast<float4*>(s1 + base) = make_float4(a[0], a[1], a[2], a[3]);
      if (opt == OPT_ADAM) *reinterpret_cast<float4*>(s2 + base) = make_float4(b[0], b[1], b[2], b[3]);
      const long long so = blk_shadow[blk];
      if (so >= 0) {
        const long long stride = blk_split_stride[blk];
        __nv_bfloat16* d = shadow + so + threadIdx.x * 4;
        __nv_bfloat16 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __float2bfloat16(w[j]);
        *reinterpret_cast<uint2*>(d) = *reinterpret_cast<uint2*>(h);
        if (stride > 0) {
          __nv_bfloat16 l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) l[j] = __float2bfloat16(w[j] - __bfloat162float(h[j]));
          *reinterpret_cast<uint2*>(d + stride) = *reinterpret_cast<uint2*>(l);
          *reinterpret_cast<uint2*>(d + 2 * stride) = *reinterpret_cast<uint2*>(h);
        }
      }
    }
  }
  if (l2_out) {
    for (int o = 16; o > 0; o >>= 1) l2acc += __shfl_xor_sync(0xffffffffu, l2acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = l2acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += sh[i];
      atomicAdd(l2_out, t);
    }
  }
}

// Refresh the bf16 shadows from the fp32 masters without touching them (after load / init).
__global__ void __launch_bounds__(256) shadow_refresh_kernel(const float* __restrict__ params,
                                                             const long long* __restrict__ blk_shadow,
                                                             const long long* __restrict__ blk_split_stride,
                                                             __nv_bfloat16* __restrict__ shadow, long long n) {
  pdl_entry();
  for (long long blk = blockIdx.x; blk * OPT_BLOCK < n; blk += gridDim.x) {
    const long long so = blk_shadow[blk];
    const long long base = blk * OPT_BLOCK + threadIdx.x * 4;
    if (so < 0 || base + 3 >= n) continue;
    const float4 w4 = *reinterpret_cast<const float4*>(params + base);
    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
    const long long stride = blk_split_stride[blk];
    __nv_bfloat16* d = shadow + so + threadIdx.x * 4;
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { h[j] = __float2bfloat16(w[j]); l[j] = __float2bfloat16(w[j] - __bfloat162float(h[j])); }
    *reinterpret_cast<uint2*>(d) = *reinterpret_cast<uint2*>(h);
    if (stride > 0) {
      *reinterpret_cast<uint2*>(d + stride) = *reinterpret_cast<uint2*>(l);
      *reinterpret_cast<uint2*>(d + 2 * stride) = *reinterpret_cast<uint2*>(h);
    }
  }
}

// loss += sum over blocks of l2/2 * w^2   (tf.losses.get_regularization_loss, trainer.py:357)
__global__ void __launch_bounds__(256) l2_loss_kernel(const float* __restrict__ params, const float* __restrict__ blk_l2,
                                                      long long n, float* out) {
  pdl_entry();
  __shared__ float sh[8];
  float acc = 0.f;
  for (long long blk = blockIdx.x; blk * OPT_BLOCK < n; blk += gridDim.x) {
    const float l2 = blk_l2[blk];
    if (l2 == 0.f) continue;
    const long long base = blk * OPT_BLOCK + threadIdx.x * 4;
    if (base + 3 < n) {
      const float4 w = *reinterpret_cast<const float4*>(params + base);
      acc += 0.5f * l2 * (w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

// Host scalars -> device memory through kernel ARGUMENTS (captured by value at launch time): lets the per-step
// learning rate / margin schedule change between replays of a captured CUDA graph without any host-buffer race.
struct Scalars16 { float v[16]; };
__global__ void set_scalars_kernel(float* dst, Scalars16 s, int n) {
  pdl_entry();
  if (threadIdx.x < n) dst[threadIdx.x] = s.v[threadIdx.x];
}

static int opt_grid(long long n, int sms) {
  long long blocks = (n + OPT_BLOCK - 1) / OPT_BLOCK;
  const long long cap = static_cast<long long>(sms) * 8;
  return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}


// Gradient exchange in bf16 (optional, parallel.DataParallel grad_dtype="bf16"): the flat fp32 gradient buffer is rounded
// to bf16 for the all-reduce (half the NVLink bytes) and widened again for the optimizer.  8 elements per thread.
__global__ void __launch_bounds__(256) grad_pack_bf16_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_entry();
  for (long long i = (blockIdx.x * 256LL + threadIdx.x) * 8; i + 7 < n; i += gridDim.x * 2048LL) {
    const float4 a = *reinterpret_cast<const float4*>(g + i), b = *reinterpret_cast<const float4*>(g + i + 4);
    __nv_bfloat162 h[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                           __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<uint4*>(h);
  }
}
__global__ void __launch_bounds__(256) grad_unpack_bf16_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ g, long long n) {
  pdl_entry();
  for (long long i = (blockIdx.x * 256LL + threadIdx.x) * 8; i + 7 < n; i += gridDim.x * 2048LL) {
    const uint4 r = *reinterpret_cast<const uint4*>(in + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
    *reinterpret_cast<float4*>(g + i) = make_float4(f0.x, f0.y, f1.x, f1.y);
    *reinterpret_cast<float4*>(g + i + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
  }
}

}  // namespace xv

using namespace xv;

extern "C" int xv_grad_sumsq(const float* params, const float* grads, const float* blk_l2, int64_t n, float* out, void* stream) {
  if (!params || !grads || !blk_l2 || !out || n <= 0 || n % OPT_BLOCK) return set_error(XV_ERR_INVALID, "xv_grad_sumsq: n must be a positive multiple of 1024");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  ::xv::launch_pdl((grad_sumsq_kernel), opt_grid(n, sms), 256, 0, static_cast<cudaStream_t>(stream), params, grads, blk_l2, n, out);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_opt_step(float* params, const float* grads, float* state1, float* state2, const float* blk_l2,
                           const int64_t* blk_shadow, const int64_t* blk_split_stride, void* shadow, int64_t n, int opt,
                           const float* hyper, const float* gsumsq, float* l2_loss_out, void* stream) {
  if (!params || !grads || !blk_l2 || !blk_shadow || !blk_split_stride || !hyper || n <= 0 || n % OPT_BLOCK)
    return set_error(XV_ERR_INVALID, "xv_opt_step: bad arguments (n must be a positive multiple of 1024)");
  if (opt < OPT_SGD || opt > OPT_ADAM) { set_error(XV_ERR_INVALID, "Optimizer %d is not supported.", opt); return XV_ERR_INVALID; }
  if (opt != OPT_SGD && !state1) return set_error(XV_ERR_INVALID, "xv_opt_step: momentum/adam need state1");
  if (opt == OPT_ADAM && !state2) return set_error(XV_ERR_INVALID, "xv_opt_step: adam needs state2");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  ::xv::launch_pdl((opt_step_kernel), opt_grid((n + OPT_UNROLL - 1) / OPT_UNROLL, sms), 256, 0, static_cast<cudaStream_t>(stream), 
      params, grads, state1, state2, blk_l2, reinterpret_cast<const long long*>(blk_shadow),
      reinterpret_cast<const long long*>(blk_split_stride), static_cast<__nv_bfloat16*>(shadow), n, opt, hyper, gsumsq,
      l2_loss_out);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_shadow_refresh(const float* params, const int64_t* blk_shadow, const int64_t* blk_split_stride,
                                 void* shadow, int64_t n, void* stream) {
  if (!params || !blk_shadow || !blk_split_stride || !shadow || n <= 0 || n % OPT_BLOCK)
    return set_error(XV_ERR_INVALID, "xv_shadow_refresh: bad arguments");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  ::xv::launch_pdl((shadow_refresh_kernel), opt_grid(n, sms), 256, 0, static_cast<cudaStream_t>(stream), 
      params, reinterpret_cast<const long long*>(blk_shadow), reinterpret_cast<const long long*>(blk_split_stride),
      static_cast<__nv_bfloat16*>(shadow), n);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_l2_loss(const float* params, const float* blk_l2, int64_t n, float* out, void* stream) {
  if (!params || !blk_l2 || !out || n <= 0 || n % OPT_BLOCK) return set_error(XV_ERR_INVALID, "xv_l2_loss: bad arguments");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  ::xv::launch_pdl((l2_loss_kernel), opt_grid(n, sms), 256, 0, static_cast<cudaStream_t>(stream), params, blk_l2, n, out);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_set_scalars(float* dst, const float* host_vals, int n, void* stream) {
  if (!dst || !host_vals || n <= 0 || n > 16) return set_error(XV_ERR_INVALID, "xv_set_scalars: n must be in [1, 16]");
  Scalars16 s;
  for (int i = 0; i < 16; ++i) s.v[i] = i < n ? host_vals[i] : 0.f;
  ::xv::launch_pdl((set_scalars_kernel), 1, 32, 0, static_cast<cudaStream_t>(stream), dst, s, n);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_grad_pack_bf16(const float* grads, void* out, int64_t n, void* stream) {
  if (!grads || !out || n <= 0 || n % 8) return set_error(XV_ERR_INVALID, "xv_grad_pack_bf16: n must be a positive multiple of 8");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  long long g = (n / 8 + 255) / 256; if (g > sms * 8LL) g = sms * 8LL;
  ::xv::launch_pdl((grad_pack_bf16_kernel), static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream), grads,
                   static_cast<__nv_bfloat16*>(out), static_cast<long long>(n));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_grad_unpack_bf16(const void* in, float* grads, int64_t n, void* stream) {
  if (!grads || !in || n <= 0 || n % 8) return set_error(XV_ERR_INVALID, "xv_grad_unpack_bf16: n must be a positive multiple of 8");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  long long g = (n / 8 + 255) / 256; if (g > sms * 8LL) g = sms * 8LL;
  ::xv::launch_pdl((grad_unpack_bf16_kernel), static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream),
                   static_cast<const __nv_bfloat16*>(in), grads, static_cast<long long>(n));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
