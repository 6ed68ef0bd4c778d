// NetVLAD / GhostVLAD pooling (model/pooling.py:195-277) -- forward and backward.
//
//   post[b,t,:]  = softmax_j(logits[b,t,j]),  j over the K real + G ghost clusters        (pooling.py:249-250)
//   res[b,k,:]   = sum_t post[b,t,k] * (value[b,t,:] - centers[k,:])   for the K real clusters (pooling.py:259-269)
//                = sum_t post[b,t,k] value[b,t,:]  -  mass[b,k] * centers[k,:],   mass[b,k] = sum_t post[b,t,k]
//   out[b,k,:]   = res[b,k,:] * rsqrt(max(|res[b,k,:]|^2, 1e-12))                        (pooling.py:271)
//   out[b,:]    *= rsqrt(max(|out[b,:]|^2, 1e-12))     when vlad_final_l2_norm           (pooling.py:273-274)
//
// The key / value networks and the `vlad_weight_affine` layer in front of this are ordinary frame layers on the tcgen05
// GEMM; what lives here is HBM / L2-bound CUDA-core work on a [K, C] accumulator per utterance.  Layouts: value bf16
// flat-time [B*T, ld]; logits bf16 [B*T, ldl]; post f32 [B, T, KG]; res / gres f32 [B, K, cpad]; out f32 [B, K*cpad] with
// the [hi | hi | lo] bf16 split copy that feeds the tdnn6 GEMM.  Frames t >= length of a segment carry post = 0.
#include <cuda_bf16.h>
#include <math.h>

#include "xv_internal.h"

namespace xv {

constexpr int VLAD_MAX_KG = 64;      // real + ghost clusters (two logits per lane of the per-frame warp)
constexpr int VLAD_KC = 8;           // clusters per accumulator pass of the pooling kernels
constexpr float VLAD_L2_EPS = 1e-12f;      // tf.nn.l2_normalize epsilon

struct alignas(16) VBf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void v_load8(const __nv_bfloat16* p, float (&f)[8]) {
  const VBf16x8 r = *reinterpret_cast<const VBf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(r.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void v_store8(__nv_bfloat16* p, const float (&f)[8]) {
  VBf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<VBf16x8*>(p) = r;
}
__device__ __forceinline__ void v_load8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float v_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float v_warp_max(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// Cluster posteriors: one warp per frame, lane j holds logits j and j + 32.
__global__ void __launch_bounds__(256) vlad_post_fwd_kernel(const __nv_bfloat16* __restrict__ logits, float* __restrict__ post,
                                                            int rows, int seg_len, int seg_valid,
                                                            const int* __restrict__ lengths, int KG, int ldl) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= rows) return;
  const int b = m / seg_len, t = m - b * seg_len;
  const int L = lengths ? lengths[b] : seg_valid;
  float* o = post + static_cast<long long>(m) * KG;
  if (t >= L) {
    if (lane < KG) o[lane] = 0.f;
    if (lane + 32 < KG) o[lane + 32] = 0.f;
    return;
  }
  const __nv_bfloat16* lr = logits + static_cast<long long>(m) * ldl;
  const float e0 = (lane < KG) ? __bfloat162float(lr[lane]) : -INFINITY;
  const float e1 = (lane + 32 < KG) ? __bfloat162float(lr[lane + 32]) : -INFINITY;
  const float mx = v_warp_max(fmaxf(e0, e1));
  const float p0 = (lane < KG) ? expf(e0 - mx) : 0.f;
  const float p1 = (lane + 32 < KG) ? expf(e1 - mx) : 0.f;
  const float inv = 1.0f / v_warp_sum(p0 + p1);
  if (lane < KG) o[lane] = p0 * inv;
  if (lane + 32 < KG) o[lane + 32] = p1 * inv;
}

// ------------------------------------------------------------------------------------------------
// Residual aggregation.  grid = (cpad/256, B, ceil(K/8)); 8 warps stride over the frames, a lane owns 8 channels and
// VLAD_KC clusters: acc[k][c] += post[t,k] * value[t,c].  Cross-warp reduction through shared memory, one cluster at a time.
__global__ void __launch_bounds__(256) vlad_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ post,
                                                            const float* __restrict__ centers, float* __restrict__ res,
                                                            float* __restrict__ mass, float* __restrict__ sumsq, int seg_len,
                                                            int seg_valid, const int* __restrict__ lengths, int K, int KG,
                                                            int c_real, int cpad, long long ld, int ldc) {
  pdl_entry();
  __shared__ float red[8][256];
  __shared__ float red_m[8][VLAD_KC];
  __shared__ float red_s[8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int b = blockIdx.y, k0 = blockIdx.z * VLAD_KC;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int L = lengths ? lengths[b] : seg_valid;
  const __nv_bfloat16* xb = x + static_cast<long long>(b) * seg_len * ld;
  const float* pb = post + static_cast<long long>(b) * seg_len * KG;
  float acc[VLAD_KC][8], am[VLAD_KC];
#pragma unroll
  for (int k = 0; k < VLAD_KC; ++k) {
    am[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  }
  const bool live = c0 < cpad;
  for (int t = wp; t < L; t += 8) {
    float v[8];
    if (live) v_load8(xb + static_cast<long long>(t) * ld + c0, v);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < VLAD_KC; ++k) {
      const float a = (k0 + k < K) ? pb[static_cast<long long>(t) * KG + k0 + k] : 0.f;
      am[k] += a;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(a, v[j], acc[k][j]);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < VLAD_KC; ++k) red_m[wp][k] = am[k];
  }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  for (int k = 0; k < VLAD_KC; ++k) {
    if (k0 + k >= K) break;                 // uniform over the block
#pragma unroll
    for (int j = 0; j < 8; ++j) red[wp][lane * 8 + j] = acc[k][j];
    __syncthreads();
    float m = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) m += red_m[w][k];
    float r = 0.f;
    if (c < cpad) {
#pragma unroll
      for (int w = 0; w < 8; ++w) r += red[w][threadIdx.x];
      r = (c < c_real) ? r - m * centers[static_cast<long long>(k0 + k) * ldc + c] : 0.f;
      res[(static_cast<long long>(b) * K + k0 + k) * cpad + c] = r;
    }
    const float sq = v_warp_sum(r * r);
    if (lane == 0) red_s[wp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red_s[w];
      atomicAdd(sumsq + static_cast<long long>(b) * K + k0 + k, s);
      if (blockIdx.x == 0) mass[static_cast<long long>(b) * K + k0 + k] = m;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Intra-cluster (+ optional final) L2 normalisation and the bf16 split copy.  grid = B.
__device__ __forceinline__ float vlad_final_sq(const float* __restrict__ sq, int K) {
  float tot = 0.f;
  for (int k = 0; k < K; ++k) tot += sq[k] / fmaxf(sq[k], VLAD_L2_EPS);
  return tot;
}

__global__ void __launch_bounds__(256) vlad_norm_fwd_kernel(const float* __restrict__ res, const float* __restrict__ sumsq,
                                                            float* __restrict__ out, __nv_bfloat16* __restrict__ out3, int K,
                                                            int cpad, int final_norm) {
  pdl_entry();
  const int b = blockIdx.x;
  const float* sq = sumsq + static_cast<long long>(b) * K;
  const float inv_f = final_norm ? rsqrtf(fmaxf(vlad_final_sq(sq, K), VLAD_L2_EPS)) : 1.0f;
  const long long W = static_cast<long long>(K) * cpad;
  const float* rb = res + static_cast<long long>(b) * W;
  float* ob = out + static_cast<long long>(b) * W;
  for (long long i = threadIdx.x; i < W; i += 256) {
    const int k = static_cast<int>(i / cpad);
    const float v = rb[i] * rsqrtf(fmaxf(sq[k], VLAD_L2_EPS)) * inv_f;
    ob[i] = v;
    if (out3) {
      __nv_bfloat16* o3 = out3 + static_cast<long long>(b) * 3 * W;
      const __nv_bfloat16 h = __float2bfloat16(v);
      o3[i] = h;
      o3[W + i] = h;
      o3[2 * W + i] = __float2bfloat16(v - __bfloat162float(h));
    }
  }
}

// Backward of the two normalisations: gres[b,k,:] = dL/dres[b,k,:],  gc[b,k] = <gres[b,k,:], centers[k,:]>.  grid = B.
//   final:   u = out * nf (cluster-normalised), g1 = (g - out <g, out>) / nf
//   cluster: gres = (g1 - u_k <g1_k, u_k>) / n_k
__global__ void __launch_bounds__(256) vlad_norm_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                            const float* __restrict__ sumsq, const float* __restrict__ centers,
                                                            float* __restrict__ gres, float* __restrict__ gc, int K, int cpad,
                                                            int ldc, int final_norm) {
  pdl_entry();
  __shared__ float red[8];
  __shared__ float bc;
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float* sq = sumsq + static_cast<long long>(b) * K;
  const long long W = static_cast<long long>(K) * cpad;
  const float* gb = dout + static_cast<long long>(b) * W;
  const float* ob = out + static_cast<long long>(b) * W;
  float nf = 1.f, p = 0.f;
  bool f_clamped = true;
  if (final_norm) {
    const float tot = vlad_final_sq(sq, K);
    f_clamped = tot <= VLAD_L2_EPS;
    nf = sqrtf(fmaxf(tot, VLAD_L2_EPS));
    float a = 0.f;
    for (long long i = threadIdx.x; i < W; i += 256) a = fmaf(gb[i], ob[i], a);
    a = v_warp_sum(a);
    if (lane == 0) red[wp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w];
      bc = s;
    }
    __syncthreads();
    p = f_clamped ? 0.f : bc;
  }
  const float inv_f = 1.0f / nf;
  for (int k = wp; k < K; k += 8) {
    const float* gk = gb + static_cast<long long>(k) * cpad;
    const float* ok = ob + static_cast<long long>(k) * cpad;
    const float* ck = centers + static_cast<long long>(k) * ldc;
    float* rk = gres + (static_cast<long long>(b) * K + k) * cpad;
    const bool clamped = sq[k] <= VLAD_L2_EPS;
    const float inv_n = rsqrtf(fmaxf(sq[k], VLAD_L2_EPS));
    float q = 0.f;
    for (int c = lane; c < cpad; c += 32) {
      const float o = ok[c], u = o * nf;
      const float g1 = (gk[c] - o * p) * inv_f;
      q = fmaf(g1, u, q);
    }
    q = clamped ? 0.f : v_warp_sum(q);
    float d = 0.f;
    for (int c = lane; c < cpad; c += 32) {
      const float o = ok[c], u = o * nf;
      const float g1 = (gk[c] - o * p) * inv_f;
      const float r = (g1 - u * q) * inv_n;
      rk[c] = r;
      d = fmaf(r, ck[c], d);
    }
    d = v_warp_sum(d);
    if (lane == 0) gc[static_cast<long long>(b) * K + k] = d;
  }
}

// ------------------------------------------------------------------------------------------------
// dL/dlogits: one warp per frame.  dpost[t,k] = <gres[b,k,:], value[t,:]> - gc[b,k] for the real clusters (0 for the
// ghosts), then the softmax backward  dlogit_j = post_j (dpost_j - sum_i post_i dpost_i).
__global__ void __launch_bounds__(256) vlad_dlogits_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ post,
                                                           const float* __restrict__ gres, const float* __restrict__ gc,
                                                           __nv_bfloat16* __restrict__ dlogits, int rows, int seg_len,
                                                           int seg_valid, const int* __restrict__ lengths, int K, int KG,
                                                           int cpad, long long ld, int ldl) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= rows) return;
  const int b = m / seg_len, t = m - b * seg_len;
  const int L = lengths ? lengths[b] : seg_valid;
  __nv_bfloat16* dl = dlogits + static_cast<long long>(m) * ldl;
  if (t >= L) {
    for (int j = lane; j < ldl; j += 32) dl[j] = __float2bfloat16(0.f);
    return;
  }
  const __nv_bfloat16* xr = x + static_cast<long long>(m) * ld;
  const float* gb = gres + static_cast<long long>(b) * K * cpad;
  float dp0 = 0.f, dp1 = 0.f;           // dpost of clusters lane and lane + 32
  for (int k0 = 0; k0 < K; k0 += VLAD_KC) {
    float acc[VLAD_KC];
#pragma unroll
    for (int k = 0; k < VLAD_KC; ++k) acc[k] = 0.f;
    for (int c0 = lane * 8; c0 < cpad; c0 += 256) {
      float v[8];
      v_load8(xr + c0, v);
#pragma unroll
      for (int k = 0; k < VLAD_KC; ++k) {
        if (k0 + k < K) {
          float g[8];
          v_load8f(gb + static_cast<long long>(k0 + k) * cpad + c0, g);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(g[j], v[j], acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VLAD_KC; ++k) {
      if (k0 + k < K) {
        const float s = v_warp_sum(acc[k]) - gc[static_cast<long long>(b) * K + k0 + k];
        if (((k0 + k) & 31) == lane) {
          if (k0 + k < 32) dp0 = s; else dp1 = s;
        }
      }
    }
  }
  const float* pr = post + static_cast<long long>(m) * KG;
  const float a0 = (lane < KG) ? pr[lane] : 0.f;
  const float a1 = (lane + 32 < KG) ? pr[lane + 32] : 0.f;
  const float s = v_warp_sum(a0 * dp0 + a1 * dp1);
  for (int j = lane; j < ldl; j += 32) {
    float d = 0.f;
    if (j == lane && j < KG) d = a0 * (dp0 - s);
    else if (j == lane + 32 && j < KG) d = a1 * (dp1 - s);
    dl[j] = __float2bfloat16(d);
  }
}

// dL/dvalue[t,c] = sum_{k<K} post[t,k] gres[b,k,c]   (+= when accumulate).  grid = (cpad/256, B); the [K, 256] slice of
// gres sits in shared memory, 8 warps stride over the frames, 8 channels per lane.
__global__ void __launch_bounds__(256) vlad_dvalue_kernel(const float* __restrict__ post, const float* __restrict__ gres,
                                                          __nv_bfloat16* __restrict__ dx, int seg_len, int seg_valid,
                                                          const int* __restrict__ lengths, int K, int KG, int cpad,
                                                          long long ld, int accumulate) {
  pdl_entry();
  extern __shared__ float sg[];          // [K][256]
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int cb = blockIdx.x * 256;
  const int c0 = cb + lane * 8;
  const int L = lengths ? lengths[b] : seg_valid;
  for (int i = threadIdx.x; i < K * 256; i += 256) {
    const int k = i >> 8, c = cb + (i & 255);
    sg[i] = (c < cpad) ? gres[(static_cast<long long>(b) * K + k) * cpad + c] : 0.f;
  }
  __syncthreads();
  if (c0 >= cpad) return;
  const float* pb = post + static_cast<long long>(b) * seg_len * KG;
  __nv_bfloat16* db = dx + static_cast<long long>(b) * seg_len * ld;
  for (int t = wp; t < seg_len; t += 8) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    __nv_bfloat16* dst = db + static_cast<long long>(t) * ld + c0;
    if (t < L) {
      for (int k = 0; k < K; ++k) {
        const float a = pb[static_cast<long long>(t) * KG + k];
        float g[8];
        v_load8f(sg + k * 256 + lane * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(a, g[j], o[j]);
      }
      if (accumulate) {
        float prev[8];
        v_load8(dst, prev);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += prev[j];
      }
    } else if (accumulate) {
      continue;
    }
    v_store8(dst, o);
  }
}

// dL/dcenters[k,c] += -sum_b mass[b,k] gres[b,k,c]   (ghost centres receive no data gradient).  grid = (cpad/256, K).
__global__ void __launch_bounds__(256) vlad_dcenters_kernel(const float* __restrict__ mass, const float* __restrict__ gres,
                                                            float* __restrict__ dcenters, int B, int K, int c_real, int cpad,
                                                            int ldc) {
  pdl_entry();
  const int k = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
  if (c >= c_real) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s = fmaf(mass[static_cast<long long>(b) * K + k], gres[(static_cast<long long>(b) * K + k) * cpad + c], s);
  dcenters[static_cast<long long>(k) * ldc + c] -= s;
}

}  // namespace xv

using namespace xv;

static int vlad_check(const char* who, int B, int seg_len, int K, int KG, int c_real, int cpad) {
  if (B <= 0 || seg_len <= 0 || K < 1 || KG < K || KG > VLAD_MAX_KG)
    return set_error(XV_ERR_UNSUPPORTED, "%s: need 1 <= vlad_num_centers <= centers + ghosts <= %d", who, VLAD_MAX_KG);
  if (c_real <= 0 || c_real > cpad || cpad % 8) return set_error(XV_ERR_INVALID, "%s: bad channel counts", who);
  return XV_OK;
}

extern "C" int xv_vlad_post_fwd(const void* logits, float* post, int B, int seg_len, int seg_valid, const int32_t* lengths, int KG,
                                int ldl, void* stream) {
  if (!logits || !post || B <= 0 || seg_len <= 0 || KG < 1 || KG > VLAD_MAX_KG || ldl < KG)
    return set_error(XV_ERR_INVALID, "xv_vlad_post_fwd: bad arguments (1 <= clusters <= %d <= ldl)", VLAD_MAX_KG);
  const long long rows = static_cast<long long>(B) * seg_len;
  if (rows > 0x7fffffffLL) return set_error(XV_ERR_INVALID, "xv_vlad_post_fwd: rows must fit in int32");
  ::xv::launch_pdl((vlad_post_fwd_kernel), ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream),
                   static_cast<const __nv_bfloat16*>(logits), post, static_cast<int>(rows), seg_len, seg_valid, lengths, KG, ldl);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_vlad_pool_fwd(const void* value, const float* post, const float* centers, float* res, float* mass, float* sumsq,
                                float* out, void* out_split, int B, int seg_len, int seg_valid, const int32_t* lengths, int K,
                                int KG, int c_real, int cpad, int64_t ld, int ldc, int final_norm, void* stream) {
  if (!value || !post || !centers || !res || !mass || !sumsq || !out || ld % 8 || ld < cpad || ldc < c_real)
    return set_error(XV_ERR_INVALID, "xv_vlad_pool_fwd: bad arguments");
  int rc = vlad_check("xv_vlad_pool_fwd", B, seg_len, K, KG, c_real, cpad); if (rc) return rc;
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  XV_CUDA_CHECK(cudaMemsetAsync(sumsq, 0, sizeof(float) * static_cast<size_t>(B) * K, s_));
  dim3 grid(ceil_div(cpad, 256), B, ceil_div(K, VLAD_KC));
  ::xv::launch_pdl((vlad_pool_fwd_kernel), grid, 256, 0, s_, static_cast<const __nv_bfloat16*>(value), post, centers, res, mass,
                   sumsq, seg_len, seg_valid, lengths, K, KG, c_real, cpad, static_cast<long long>(ld), ldc);
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((vlad_norm_fwd_kernel), B, 256, 0, s_, static_cast<const float*>(res), static_cast<const float*>(sumsq), out,
                   static_cast<__nv_bfloat16*>(out_split), K, cpad, final_norm);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_vlad_pool_bwd(const void* value, const float* post, const float* centers, const float* mass, const float* sumsq,
                                const float* out, const float* dout, float* gres, float* gc, void* dlogits, void* dvalue,
                                float* dcenters, int B, int seg_len, int seg_valid, const int32_t* lengths, int K, int KG,
                                int c_real, int cpad, int64_t ld, int ldl, int ldc, int final_norm, int accumulate_dvalue,
                                void* stream) {
  if (!value || !post || !centers || !mass || !sumsq || !out || !dout || !gres || !gc || !dlogits || !dvalue || !dcenters ||
      ld % 8 || ld < cpad || ldl < KG || ldc < c_real)
    return set_error(XV_ERR_INVALID, "xv_vlad_pool_bwd: bad arguments");
  int rc = vlad_check("xv_vlad_pool_bwd", B, seg_len, K, KG, c_real, cpad); if (rc) return rc;
  const long long rows = static_cast<long long>(B) * seg_len;
  if (rows > 0x7fffffffLL) return set_error(XV_ERR_INVALID, "xv_vlad_pool_bwd: rows must fit in int32");
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  ::xv::launch_pdl((vlad_norm_bwd_kernel), B, 256, 0, s_, dout, out, sumsq, centers, gres, gc, K, cpad, ldc, final_norm);
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((vlad_dlogits_kernel), ceil_div(rows, 8), 256, 0, s_, static_cast<const __nv_bfloat16*>(value), post,
                   static_cast<const float*>(gres), static_cast<const float*>(gc), static_cast<__nv_bfloat16*>(dlogits),
                   static_cast<int>(rows), seg_len, seg_valid, lengths, K, KG, cpad, static_cast<long long>(ld), ldl);
  XV_CUDA_CHECK(cudaGetLastError());
  const size_t smem = static_cast<size_t>(K) * 256 * sizeof(float);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      XV_CUDA_CHECK(cudaFuncSetAttribute(vlad_dvalue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VLAD_MAX_KG * 1024));
      configured = true;
    }
  }
  dim3 gv(ceil_div(cpad, 256), B);
  ::xv::launch_pdl((vlad_dvalue_kernel), gv, 256, smem, s_, post, static_cast<const float*>(gres),
                   static_cast<__nv_bfloat16*>(dvalue), seg_len, seg_valid, lengths, K, KG, cpad, static_cast<long long>(ld),
                   accumulate_dvalue);
  XV_CUDA_CHECK(cudaGetLastError());
  dim3 gcn(ceil_div(cpad, 256), K);
  ::xv::launch_pdl((vlad_dcenters_kernel), gcn, 256, 0, s_, mass, static_cast<const float*>(gres), dcenters, B, K, c_real, cpad, ldc);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
