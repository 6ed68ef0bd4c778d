// Multi-head attentive statistics pooling (model/pooling.py:37-192) -- the HBM-bound part.  The key / value
// networks are ordinary frame layers (tcgen05 GEMM + BN/activation kernels); what lives here is
//   scores  e[b,h,t] = scale * <key[b,t,:], q_h>            (one pass over the key tensor, warp per frame)
//   weights w[b,h,:] = softmax_t(e[b,h,:])                   (masked to the valid frames)
//   pooling mean[b,c] = sum_t w[b,h(c),t] v[b,t,c],  std[b,c] = sqrt(max(sum_t w (v-mean)^2, 1e-12))
//   penalty coef/B * sum_b ||W_b W_b^T - I||_F^2            (pooling.py:185-189)
// and their backward passes.  Layouts: key / value bf16 flat-time [B*T, ld]; scores / weights f32 [B, H, T];
// the query is expanded to "qpad" f32 [H, ldk] (zero outside the head's key slice when att_split_key) so that every
// kernel runs the same dense inner product.  Heads own contiguous channel ranges of width dv/H (split_heads,
// model/common.py:239-249), which need not be a multiple of 8: an 8-channel vector straddles at most two heads.
#include <cuda_bf16.h>
#include <math.h>

#include "xv_internal.h"

namespace xv {

constexpr int ATT_MAX_HEADS = 16;

struct alignas(16) ABf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void a_load8(const __nv_bfloat16* p, float (&f)[8]) {
  const ABf16x8 r = *reinterpret_cast<const ABf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(r.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void a_unpack8(const ABf16x8& r, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(r.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void a_store8(__nv_bfloat16* p, const float (&f)[8]) {
  ABf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<ABf16x8*>(p) = r;
}
__device__ __forceinline__ void a_load8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// query [H, dq] -> qpad [H, ldk]:  non-split: qpad[h, d] = q[h, d] (d < dq);  split: qpad[h, h*dq + d] = q[h, d].
__global__ void att_expand_query_kernel(const float* __restrict__ q, float* __restrict__ qpad, int H, int dq, int ldk,
                                        int split) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * ldk) return;
  const int h = i / ldk, d = i % ldk;
  const int dd = split ? d - h * dq : d;
  qpad[i] = (dd >= 0 && dd < dq) ? q[h * dq + dd] : 0.f;
}
// dqpad [H, ldk] -> dq [H, dq] (+=)
__global__ void att_fold_query_grad_kernel(const float* __restrict__ dqpad, float* __restrict__ dq, int H, int dqn,
                                           int ldk, int split) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * dqn) return;
  const int h = i / dqn, d = i % dqn;
  dq[i] += dqpad[h * ldk + (split ? h * dqn + d : d)];
}

// ------------------------------------------------------------------------------------------------
// scores: one warp per frame; the key row is read once (16-byte vectors), qpad comes from L1.
__global__ void __launch_bounds__(256) att_scores_fwd_kernel(const __nv_bfloat16* __restrict__ key,
                                                             const float* __restrict__ qpad, float* __restrict__ scores,
                                                             int rows, int seg_len, int seg_valid,
                                                             const int* __restrict__ lengths, int H, int ldk, float scale) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= rows) return;
  const int b = m / seg_len, t = m - b * seg_len;
  const int L = lengths ? lengths[b] : seg_valid;
  if (t >= L) return;
  float acc[ATT_MAX_HEADS];
#pragma unroll
  for (int h = 0; h < ATT_MAX_HEADS; ++h) acc[h] = 0.f;
  const __nv_bfloat16* kr = key + static_cast<long long>(m) * ldk;
  for (int c0 = lane * 8; c0 < ldk; c0 += 256) {
    float kv[8];
    a_load8(kr + c0, kv);
#pragma unroll
    for (int h = 0; h < ATT_MAX_HEADS; ++h) {
      if (h < H) {
        float qv[8];
        a_load8f(qpad + static_cast<long long>(h) * ldk + c0, qv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[h] = fmaf(kv[j], qv[j], acc[h]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < ATT_MAX_HEADS; ++h) {
    if (h < H) {
      const float s = warp_sum(acc[h]);
      if (lane == 0) scores[(static_cast<long long>(b) * H + h) * seg_len + t] = s * scale;
    }
  }
}

// Same scores, ldk <= 2048 and H * ldk floats of shared memory: the query sits in shared memory, a warp takes four frames of a
// 32-frame block and issues all of a frame's 16-byte loads (kept packed) before the first use.  The kernel above keeps one
// load in flight per lane and leaves after one frame (46 us for 78 MB at config 4).
template <int HT>
__global__ void __launch_bounds__(256) att_scores_fwd_fast_kernel(const __nv_bfloat16* __restrict__ key,
                                                                  const float* __restrict__ qpad, float* __restrict__ scores,
                                                                  int rows, int seg_len, int seg_valid,
                                                                  const int* __restrict__ lengths, int H, int ldk, float scale) {
  pdl_entry();
  extern __shared__ float sq[];       // [H][ldk]
  for (int i = threadIdx.x; i < H * ldk; i += 256) sq[i] = qpad[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int f = 0; f < 4; ++f) {
    const int m = blockIdx.x * 32 + f * 8 + wp;
    if (m >= rows) break;
    const int b = m / seg_len, t = m - b * seg_len;
    const int L = lengths ? lengths[b] : seg_valid;
    if (t >= L) continue;
    const __nv_bfloat16* kr = key + static_cast<long long>(m) * ldk;
    ABf16x8 raw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c0 = lane * 8 + 256 * i;
      if (c0 < ldk) raw[i] = *reinterpret_cast<const ABf16x8*>(kr + c0);
    }
    float acc[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) acc[h] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c0 = lane * 8 + 256 * i;
      if (c0 < ldk) {
        float kv[8];
        a_unpack8(raw[i], kv);
#pragma unroll
        for (int h = 0; h < HT; ++h) {
          if (h < H) {
            float qv[8];
            a_load8f(sq + h * ldk + c0, qv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[h] = fmaf(kv[j], qv[j], acc[h]);
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      if (h < H) {
        const float s = warp_sum(acc[h]);
        if (lane == 0) scores[(static_cast<long long>(b) * H + h) * seg_len + t] = s * scale;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// softmax over the valid frames of one (segment, head); frames >= L get weight 0.  In place allowed.
__global__ void __launch_bounds__(256) att_softmax_fwd_kernel(const float* __restrict__ scores, float* __restrict__ w,
                                                              int H, int seg_len, int seg_valid,
                                                              const int* __restrict__ lengths) {
  pdl_entry();
  __shared__ float red[8];
  __shared__ float bc;
  const int bh = blockIdx.x, b = bh / H;
  const int L = lengths ? lengths[b] : seg_valid;
  const float* e = scores + static_cast<long long>(bh) * seg_len;
  float* o = w + static_cast<long long>(bh) * seg_len;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < L; t += 256) mx = fmaxf(mx, e[t]);
  mx = warp_max(mx);
  if (lane == 0) red[wid] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, red[i]);
    bc = v;
  }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (int t = threadIdx.x; t < L; t += 256) s += expf(e[t] - mx);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[i];
    bc = v;
  }
  __syncthreads();
  const float inv = 1.0f / bc;
  for (int t = threadIdx.x; t < seg_len; t += 256) o[t] = (t < L) ? expf(e[t] - mx) * inv : 0.f;
}

// ------------------------------------------------------------------------------------------------
// weighted statistics pooling.  grid = (cpad/256, B), 8 warps striding over time, 8 channels per thread;
// shifted one-pass moments (shift = first frame): mean = x0 + S1, var = S2 - S1^2 with S1 = sum w (x-x0),
// S2 = sum w (x-x0)^2 (the weights of a head sum to 1).
struct HeadPair { int lo, hi, split; };   // channels [c0, c0+split) belong to head lo, the rest to head hi
__device__ __forceinline__ HeadPair head_pair(int c0, int dvh, int H) {
  HeadPair hp;
  hp.lo = min(c0 / dvh, H - 1);
  hp.hi = min((c0 + 7) / dvh, H - 1);
  hp.split = (hp.hi == hp.lo) ? 8 : (hp.hi * dvh - c0);
  return hp;
}

__global__ void __launch_bounds__(256) att_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                           float* __restrict__ out, __nv_bfloat16* __restrict__ out3,
                                                           int H, int seg_len, int seg_valid,
                                                           const int* __restrict__ lengths, int c_real, int cpad,
                                                           long long ld) {
  pdl_entry();
  __shared__ float red[8][2][256];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int L = lengths ? lengths[b] : seg_valid;
  const int dvh = c_real / H;
  const __nv_bfloat16* xb = x + static_cast<long long>(b) * seg_len * ld;
  float s1[8], s2[8], x0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = x0[j] = 0.f;
  if (c0 < cpad && L > 0) {
    const HeadPair hp = head_pair(c0, dvh, H);
    const float* wl = w + (static_cast<long long>(b) * H + hp.lo) * seg_len;
    const float* wh = w + (static_cast<long long>(b) * H + hp.hi) * seg_len;
    a_load8(xb + c0, x0);
    for (int t = wp; t < L; t += 16) {
      float v[2][8];
      float wa[2], wb[2];
      const bool ok1 = (t + 8) < L;
      a_load8(xb + static_cast<long long>(t) * ld + c0, v[0]);
      wa[0] = wl[t]; wb[0] = wh[t];
      if (ok1) {
        a_load8(xb + static_cast<long long>(t + 8) * ld + c0, v[1]);
        wa[1] = wl[t + 8]; wb[1] = wh[t + 8];
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !ok1) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float wt = (j < hp.split) ? wa[u] : wb[u];
          const float d = v[u][j] - x0[j];
          s1[j] = fmaf(wt, d, s1[j]);
          s2[j] = fmaf(wt * d, d, s2[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[wp][0][lane * 8 + j] = s1[j]; red[wp][1][lane * 8 + j] = s2[j]; }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cpad) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[k][0][threadIdx.x]; q += red[k][1][threadIdx.x]; }
    float mean = 0.f, sd = 0.f;
    if (c < c_real && L > 0) {
      const float first = __bfloat162float(xb[c]);
      mean = first + a;
      float var = q - a * a;
      var = (var <= 1e-12f) ? 1e-12f : var;       // VAR2STD_EPSILON floor (pooling.py:166-168)
      sd = sqrtf(var);
    }
    float* ob = out + static_cast<long long>(b) * 2 * cpad;
    ob[c] = mean;
    ob[cpad + c] = sd;
    if (out3) {   // [hi | hi | lo] split copy feeding the tdnn6 GEMM
      __nv_bfloat16* o3 = out3 + static_cast<long long>(b) * 6 * cpad;
      const __nv_bfloat16 mh = __float2bfloat16(mean), sh = __float2bfloat16(sd);
      o3[c] = mh; o3[cpad + c] = sh;
      o3[2 * cpad + c] = mh; o3[3 * cpad + c] = sh;
      o3[4 * cpad + c] = __float2bfloat16(mean - __bfloat162float(mh));
      o3[5 * cpad + c] = __float2bfloat16(sd - __bfloat162float(sh));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pooling backward.  grid = (ceil(T/64), B); per block the per-channel coefficients (mean, gm, 2*gv) are staged in
// shared memory once, then warp wp walks frames t0+wp, t0+wp+8, ...:
//   dv[t,c]  = w[h(c),t] * (gm_c + 2 gv_c (v - mean_c))          (+= when accumulate)
//   dw[h,t]  = sum_{c in h} gm_c v + gv_c (v - mean_c)^2
// with gm = dL/dmean, gv = dL/dstd / (2 std) where the variance is above the floor, else 0.
template <bool ONE_HEAD, bool ACC>
__global__ void __launch_bounds__(256, ONE_HEAD ? 4 : 1) att_pool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ pooled,
                                                           const float* __restrict__ dpooled,
                                                           __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, int H,
                                                           int seg_len, int seg_valid, const int* __restrict__ lengths,
                                                           int c_real, int cpad, long long ld, int accumulate) {
  pdl_entry();
  extern __shared__ float sm[];       // [3][cpad]: mean | gm | gv
  float* s_mu = sm;
  float* s_gm = sm + cpad;
  float* s_gv = sm + 2 * cpad;
  const int b = blockIdx.y;
  const int L = lengths ? lengths[b] : seg_valid;
  const float* pb = pooled + static_cast<long long>(b) * 2 * cpad;
  const float* gb = dpooled + static_cast<long long>(b) * 2 * cpad;
  float csum = 0.f;
  for (int c = threadIdx.x; c < cpad; c += 256) {
    float mu = 0.f, gm = 0.f, gv = 0.f;
    if (c < c_real) {
      mu = pb[c];
      gm = gb[c];
      const float sd = pb[cpad + c];
      if (sd * sd > 1.0000001e-12f) gv = gb[cpad + c] / (2.0f * sd);
    }
    if (ONE_HEAD) {
      // dv = w (A v + Bc), dw = sum_c v (A/2 v + Bc) + sum_c Cc   with A = 2 gv, Bc = gm - 2 gv mu, Cc = gv mu^2:
      // two per-channel vectors instead of three, 5 flops per element instead of 8
      s_mu[c] = 2.0f * gv;
      s_gm[c] = gm - 2.0f * gv * mu;
      csum += gv * mu * mu;
    } else {
      s_mu[c] = mu; s_gm[c] = gm; s_gv[c] = gv;
    }
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  __shared__ float s_red[8];
  if (ONE_HEAD) {
    csum = warp_sum(csum);
    if (lane == 0) s_red[wp] = csum;
  }
  __syncthreads();
  if (ONE_HEAD) {
    csum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) csum += s_red[k];
  }
  const int dvh = c_real / H;
  const int t0 = blockIdx.x * 64;
  const int t1 = min(t0 + 64, seg_len);
  if (ONE_HEAD) {
    // one head (the shipped configuration), cpad <= 2048: a single weight per frame, no head bookkeeping, and ALL of a frame's
    // 16-byte loads issued before the first use (up to 8 per lane) -- the generic loop below keeps one load in flight per
    // lane and is latency-bound (95 us for 157 MB at config 4)
    for (int t = t0 + wp; t < t1; t += 8) {
      const long long m = static_cast<long long>(b) * seg_len + t;
      const bool valid = t < L;
      __nv_bfloat16* drow = dx + m * ld;
      if (!valid) {
        if (!ACC) {
          const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int c0 = lane * 8; c0 < cpad; c0 += 256) a_store8(drow + c0, z);
        }
        if (lane == 0) dw[static_cast<long long>(b) * seg_len + t] = 0.f;
        continue;
      }
      const float wa = w[static_cast<long long>(b) * seg_len + t];
      const __nv_bfloat16* xrow = x + m * ld;
      ABf16x8 raw[8], praw[8];          // rows stay packed (4 registers per 8 channels) until they are used
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c0 = lane * 8 + 256 * i;
        if (c0 < cpad) {
          raw[i] = *reinterpret_cast<const ABf16x8*>(xrow + c0);
          if (ACC) praw[i] = *reinterpret_cast<const ABf16x8*>(drow + c0);
        }
      }
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c0 = lane * 8 + 256 * i;
        if (c0 < cpad) {
          float v[8], prev[8], ca[8], cb[8], o[8];
          a_unpack8(raw[i], v);
          if (ACC) a_unpack8(praw[i], prev);
          a_load8f(s_mu + c0, ca); a_load8f(s_gm + c0, cb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a = fmaf(v[j], fmaf(0.5f * ca[j], v[j], cb[j]), a);
            o[j] = wa * fmaf(ca[j], v[j], cb[j]) + (ACC ? prev[j] : 0.f);
          }
          a_store8(drow + c0, o);
        }
      }
      a = warp_sum(a);
      if (lane == 0) dw[static_cast<long long>(b) * seg_len + t] = a + csum;
    }
    return;
  }
  for (int t = t0 + wp; t < t1; t += 8) {
    const long long m = static_cast<long long>(b) * seg_len + t;
    const bool valid = t < L;
    float acc[ATT_MAX_HEADS];
#pragma unroll
    for (int h = 0; h < ATT_MAX_HEADS; ++h) acc[h] = 0.f;
    for (int c0 = lane * 8; c0 < cpad; c0 += 256) {
      float o[8];
      if (valid) {
        const HeadPair hp = head_pair(c0, dvh, H);
        const float wa = w[(static_cast<long long>(b) * H + hp.lo) * seg_len + t];
        const float wb = (hp.hi != hp.lo) ? w[(static_cast<long long>(b) * H + hp.hi) * seg_len + t] : wa;
        float v[8], mu[8], gm[8], gv[8], prev[8];
        a_load8(x + m * ld + c0, v);
        a_load8f(s_mu + c0, mu); a_load8f(s_gm + c0, gm); a_load8f(s_gv + c0, gv);
        if (accumulate) a_load8(dx + m * ld + c0, prev);
        float plo = 0.f, phi = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[j] - mu[j];
          const float contrib = fmaf(gm[j], v[j], gv[j] * d * d);
          const bool lo = j < hp.split;
          plo += lo ? contrib : 0.f;
          phi += lo ? 0.f : contrib;
          o[j] = (lo ? wa : wb) * fmaf(2.0f * gv[j], d, gm[j]) + (accumulate ? prev[j] : 0.f);
        }
#pragma unroll
        for (int h = 0; h < ATT_MAX_HEADS; ++h)
          if (h < H) acc[h] += ((h == hp.lo) ? plo : 0.f) + ((h == hp.hi && hp.hi != hp.lo) ? phi : 0.f);
        a_store8(dx + m * ld + c0, o);
      } else if (!accumulate) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        a_store8(dx + m * ld + c0, o);
      }
    }
#pragma unroll
    for (int h = 0; h < ATT_MAX_HEADS; ++h) {
      if (h < H) {
        const float s = warp_sum(acc[h]);
        if (lane == 0) dw[(static_cast<long long>(b) * H + h) * seg_len + t] = valid ? s : 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// penalty: one block per segment; thread (i, j) of the H x H Gram matrix.  gram_out[b,i,j] = (W W^T - I)[i,j].
__global__ void __launch_bounds__(256) att_penalty_fwd_kernel(const float* __restrict__ w, float* __restrict__ gram_out,
                                                              float* __restrict__ penalty, int H, int seg_len,
                                                              int seg_valid, const int* __restrict__ lengths, float coef_over_b) {
  pdl_entry();
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int L = lengths ? lengths[b] : seg_valid;
  const int i = threadIdx.x / H, j = threadIdx.x % H;
  float g2 = 0.f;
  if (threadIdx.x < H * H) {
    const float* wi = w + (static_cast<long long>(b) * H + i) * seg_len;
    const float* wj = w + (static_cast<long long>(b) * H + j) * seg_len;
    float g = 0.f;
    for (int t = 0; t < L; ++t) g = fmaf(wi[t], wj[t], g);
    g -= (i == j) ? 1.f : 0.f;
    gram_out[(static_cast<long long>(b) * H + i) * H + j] = g;
    g2 = g * g;
  }
  g2 = warp_sum(g2);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = g2;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    atomicAdd(penalty, s * coef_over_b);
  }
}

// ------------------------------------------------------------------------------------------------
// softmax backward (+ penalty gradient): de[t] = scale * w[t] * (g[t] - sum_s w[s] g[s]),
// g[t] = dw[t] + 4 coef/B sum_j gram[h, j] w[j, t].   One block per (segment, head).  Writes de over dw.
__global__ void __launch_bounds__(256) att_softmax_bwd_kernel(const float* __restrict__ w, float* __restrict__ dw,
                                                              const float* __restrict__ gram, int H, int seg_len,
                                                              int seg_valid, const int* __restrict__ lengths,
                                                              float pen4, float scale) {
  pdl_entry();
  __shared__ float red[8];
  __shared__ float bc;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int L = lengths ? lengths[b] : seg_valid;
  const float* wr = w + static_cast<long long>(bh) * seg_len;
  float* g = dw + static_cast<long long>(bh) * seg_len;
  const float* wb = w + static_cast<long long>(b) * H * seg_len;
  float dot = 0.f;
  for (int t = threadIdx.x; t < L; t += 256) {
    float gt = g[t];
    if (gram) {
      float p = 0.f;
      for (int j = 0; j < H; ++j) p = fmaf(gram[(static_cast<long long>(b) * H + h) * H + j], wb[static_cast<long long>(j) * seg_len + t], p);
      gt = fmaf(pen4, p, gt);
      g[t] = gt;
    }
    dot = fmaf(wr[t], gt, dot);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    bc = s;
  }
  __syncthreads();
  dot = bc;
  for (int t = threadIdx.x; t < seg_len; t += 256) g[t] = (t < L) ? scale * wr[t] * (g[t] - dot) : 0.f;
}

// ------------------------------------------------------------------------------------------------
// scores backward: dkey[m, d] = sum_h de[b,h,t] qpad[h,d] (bf16, zero on invalid frames) and
// dqpad[h, d] += sum_m de[b,h,t] key[m,d].  Grid (ldk/256, rows/256): a lane owns 8 channels (16-byte accesses), a warp
// walks 32 rows two at a time, the block's [H, 256] slice of qpad sits in shared memory, the per-head partial sums of dqpad
// stay in registers (HT = H rounded up to 1 / 2 / 4 / 8 / 16) and leave through one cross-warp reduction and H * 256 global
// atomics per 256 rows.  (Round 1 used 64-row blocks of 128 channels with 8-byte accesses: 4800 blocks, 614 k atomics on
// 1536 addresses -- 264 us at config 4, the largest single kernel of the attention step.)
constexpr int ASB_ROWS = 256;

template <int HT>
__global__ void __launch_bounds__(256) att_scores_bwd_kernel(const __nv_bfloat16* __restrict__ key,
                                                             const float* __restrict__ qpad, const float* __restrict__ de,
                                                             __nv_bfloat16* __restrict__ dkey, float* __restrict__ dqpad,
                                                             int rows, int seg_len, int seg_valid,
                                                             const int* __restrict__ lengths, int H, int ldk, int accumulate) {
  pdl_entry();
  extern __shared__ float sred[];     // [HT][256] qpad slice, then reused as [8][HT][256] reduction scratch
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int cb = blockIdx.x * 256;
  const int c0 = cb + lane * 8;
  const bool c_ok = c0 < ldk;
  for (int i = threadIdx.x; i < HT * 256; i += 256) {
    const int h = i >> 8, c = cb + (i & 255);
    sred[i] = (h < H && c < ldk) ? qpad[static_cast<long long>(h) * ldk + c] : 0.f;
  }
  __syncthreads();
  float qv[HT][8], acc[HT][8];
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    a_load8f(sred + h * 256 + lane * 8, qv[h]);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[h][j] = 0.f;
  }
  __syncthreads();                     // the slice is in registers: the scratch may be overwritten below
  if (c_ok) {
    const int r0 = blockIdx.y * ASB_ROWS;
    const int r1 = min(r0 + ASB_ROWS, rows);
    for (int mb = r0 + wp; mb < r1; mb += 16) {
      float kv[2][8], dh[2][HT];
      bool live[2], inb[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {           // two independent rows in flight per warp
        const int m = mb + 8 * u;
        inb[u] = m < r1;
        live[u] = false;
        if (inb[u]) {
          const int b = m / seg_len, t = m - b * seg_len;
          const int L = lengths ? lengths[b] : seg_valid;
          live[u] = t < L;
          if (live[u]) {
            a_load8(key + static_cast<long long>(m) * ldk + c0, kv[u]);
#pragma unroll
            for (int h = 0; h < HT; ++h) dh[u][h] = (h < H) ? de[(static_cast<long long>(b) * H + h) * seg_len + t] : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!inb[u]) continue;
        const int m = mb + 8 * u;
        __nv_bfloat16* dst = dkey + static_cast<long long>(m) * ldk + c0;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        if (live[u]) {
#pragma unroll
          for (int h = 0; h < HT; ++h) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o[j] = fmaf(dh[u][h], qv[h][j], o[j]);
              acc[h][j] = fmaf(dh[u][h], kv[u][j], acc[h][j]);
            }
          }
          if (accumulate) {
            float prev[8];
            a_load8(dst, prev);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += prev[j];
          }
        } else if (accumulate) {
          continue;
        }
        a_store8(dst, o);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < HT; ++h) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sred[(wp * HT + h) * 256 + lane * 8 + j] = acc[h][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * 256; i += 256) {
    const int h = i >> 8, cc = i & 255;
    const int c = cb + cc;
    if (c < ldk) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += sred[(k * HT + h) * 256 + cc];
      atomicAdd(dqpad + static_cast<long long>(h) * ldk + c, s);
    }
  }
}

}  // namespace xv

using namespace xv;

static int att_check_heads(const char* who, int H) {
  if (H < 1 || H > ATT_MAX_HEADS) return set_error(XV_ERR_UNSUPPORTED, "%s: att_num_heads must be in [1, %d]", who, ATT_MAX_HEADS);
  return XV_OK;
}

extern "C" int xv_att_expand_query(const float* query, float* qpad, int H, int dq, int ldk, int split_key, void* stream) {
  if (!query || !qpad || dq <= 0 || ldk % 8 || (split_key ? H * dq : dq) > ldk)
    return set_error(XV_ERR_INVALID, "xv_att_expand_query: bad arguments");
  int rc = att_check_heads("xv_att_expand_query", H); if (rc) return rc;
  ::xv::launch_pdl((att_expand_query_kernel), ceil_div(static_cast<long long>(H) * ldk, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      query, qpad, H, dq, ldk, split_key);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_fold_query_grad(const float* dqpad, float* dquery, int H, int dq, int ldk, int split_key, void* stream) {
  if (!dqpad || !dquery || dq <= 0 || (split_key ? H * dq : dq) > ldk)
    return set_error(XV_ERR_INVALID, "xv_att_fold_query_grad: bad arguments");
  int rc = att_check_heads("xv_att_fold_query_grad", H); if (rc) return rc;
  ::xv::launch_pdl((att_fold_query_grad_kernel), ceil_div(static_cast<long long>(H) * dq, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      dqpad, dquery, H, dq, ldk, split_key);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_scores_fwd(const void* key, const float* qpad, float* scores, int B, int seg_len, int seg_valid,
                                 const int32_t* lengths, int H, int ldk, float scale, void* stream) {
  if (!key || !qpad || !scores || B <= 0 || seg_len <= 0 || ldk <= 0 || ldk % 8)
    return set_error(XV_ERR_INVALID, "xv_att_scores_fwd: bad arguments");
  int rc = att_check_heads("xv_att_scores_fwd", H); if (rc) return rc;
  const long long rows = static_cast<long long>(B) * seg_len;
  if (rows > 0x7fffffffLL) return set_error(XV_ERR_INVALID, "xv_att_scores_fwd: rows must fit in int32");
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(H) * ldk * sizeof(float);
  if (ldk <= 2048 && smem <= 96 * 1024) {
    const int ht = H <= 1 ? 1 : (H <= 2 ? 2 : (H <= 4 ? 4 : (H <= 8 ? 8 : 16)));
#define XV_ASF_LAUNCH(HT_)                                                                                                    \
  do {                                                                                                                        \
    if (smem > 48 * 1024) {                                                                                                   \
      static bool configured = false;                                                                                        \
      if (!configured) {                                                                                                      \
        XV_CUDA_CHECK(cudaFuncSetAttribute(att_scores_fwd_fast_kernel<HT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                           96 * 1024));                                                                       \
        configured = true;                                                                                                    \
      }                                                                                                                       \
    }                                                                                                                         \
    ::xv::launch_pdl((att_scores_fwd_fast_kernel<HT_>), ceil_div(rows, 32), 256, smem, s_,                                     \
                     static_cast<const __nv_bfloat16*>(key), qpad, scores, static_cast<int>(rows), seg_len, seg_valid, lengths, \
                     H, ldk, scale);                                                                                          \
  } while (0)
    switch (ht) {
      case 1: XV_ASF_LAUNCH(1); break;
      case 2: XV_ASF_LAUNCH(2); break;
      case 4: XV_ASF_LAUNCH(4); break;
      case 8: XV_ASF_LAUNCH(8); break;
      default: XV_ASF_LAUNCH(16); break;
    }
#undef XV_ASF_LAUNCH
  } else {
    ::xv::launch_pdl((att_scores_fwd_kernel), ceil_div(rows, 8), 256, 0, s_, static_cast<const __nv_bfloat16*>(key), qpad, scores,
                     static_cast<int>(rows), seg_len, seg_valid, lengths, H, ldk, scale);
  }
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_softmax_fwd(const float* scores, float* weights, int B, int H, int seg_len, int seg_valid,
                                  const int32_t* lengths, void* stream) {
  if (!scores || !weights || B <= 0 || seg_len <= 0) return set_error(XV_ERR_INVALID, "xv_att_softmax_fwd: bad arguments");
  int rc = att_check_heads("xv_att_softmax_fwd", H); if (rc) return rc;
  ::xv::launch_pdl((att_softmax_fwd_kernel), B * H, 256, 0, static_cast<cudaStream_t>(stream), scores, weights, H, seg_len, seg_valid, lengths);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_pool_fwd(const void* value, const float* weights, float* out, void* out_split, int B, int H,
                               int seg_len, int seg_valid, const int32_t* lengths, int c_real, int cpad, int64_t ld,
                               void* stream) {
  if (!value || !weights || !out || B <= 0 || seg_len <= 0 || cpad % 8 || c_real > cpad || c_real <= 0 || ld % 8 || ld < cpad)
    return set_error(XV_ERR_INVALID, "xv_att_pool_fwd: bad arguments");
  int rc = att_check_heads("xv_att_pool_fwd", H); if (rc) return rc;
  if (c_real % H || c_real / H < 8) return set_error(XV_ERR_INVALID, "xv_att_pool_fwd: value dim must be divisible by the heads, >= 8 channels per head");
  dim3 grid(ceil_div(cpad, 256), B);
  ::xv::launch_pdl((att_pool_fwd_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(value), weights, out, static_cast<__nv_bfloat16*>(out_split), H, seg_len,
      seg_valid, lengths, c_real, cpad, ld);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_pool_bwd(const void* value, const float* weights, const float* pooled, const float* dpooled,
                               void* dvalue, float* dweights, int B, int H, int seg_len, int seg_valid,
                               const int32_t* lengths, int c_real, int cpad, int64_t ld, int accumulate, void* stream) {
  if (!value || !weights || !pooled || !dpooled || !dvalue || !dweights || B <= 0 || seg_len <= 0 || cpad % 8 ||
      c_real > cpad || c_real <= 0 || ld % 8 || ld < cpad)
    return set_error(XV_ERR_INVALID, "xv_att_pool_bwd: bad arguments");
  int rc = att_check_heads("xv_att_pool_bwd", H); if (rc) return rc;
  if (c_real % H || c_real / H < 8) return set_error(XV_ERR_INVALID, "xv_att_pool_bwd: value dim must be divisible by the heads, >= 8 channels per head");
  const size_t smem = static_cast<size_t>(3) * cpad * sizeof(float);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      XV_CUDA_CHECK(cudaFuncSetAttribute(att_pool_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      XV_CUDA_CHECK(cudaFuncSetAttribute(att_pool_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      XV_CUDA_CHECK(cudaFuncSetAttribute(att_pool_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
    if (smem > 200 * 1024) return set_error(XV_ERR_UNSUPPORTED, "xv_att_pool_bwd: value dim too large");
  }
  dim3 grid(ceil_div(seg_len, 64), B);
#define XV_APB_ARGS static_cast<const __nv_bfloat16*>(value), weights, pooled, dpooled, static_cast<__nv_bfloat16*>(dvalue), \
                    dweights, H, seg_len, seg_valid, lengths, c_real, cpad, ld, accumulate
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  if (H == 1 && cpad <= 2048) {
    if (accumulate) ::xv::launch_pdl((att_pool_bwd_kernel<true, true>), grid, 256, smem, s_, XV_APB_ARGS);
    else ::xv::launch_pdl((att_pool_bwd_kernel<true, false>), grid, 256, smem, s_, XV_APB_ARGS);
  } else {
    ::xv::launch_pdl((att_pool_bwd_kernel<false, false>), grid, 256, smem, s_, XV_APB_ARGS);
  }
#undef XV_APB_ARGS
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_penalty_fwd(const float* weights, float* gram, float* penalty, int B, int H, int seg_len,
                                  int seg_valid, const int32_t* lengths, float coef, void* stream) {
  if (!weights || !gram || !penalty || B <= 0 || seg_len <= 0) return set_error(XV_ERR_INVALID, "xv_att_penalty_fwd: bad arguments");
  int rc = att_check_heads("xv_att_penalty_fwd", H); if (rc) return rc;
  ::xv::launch_pdl((att_penalty_fwd_kernel), B, 256, 0, static_cast<cudaStream_t>(stream), weights, gram, penalty, H, seg_len, seg_valid,
                                                                            lengths, coef / static_cast<float>(B));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_softmax_bwd(const float* weights, float* dweights, const float* gram, int B, int H, int seg_len,
                                  int seg_valid, const int32_t* lengths, float penalty_coef, float scale, void* stream) {
  if (!weights || !dweights || B <= 0 || seg_len <= 0) return set_error(XV_ERR_INVALID, "xv_att_softmax_bwd: bad arguments");
  int rc = att_check_heads("xv_att_softmax_bwd", H); if (rc) return rc;
  ::xv::launch_pdl((att_softmax_bwd_kernel), B * H, 256, 0, static_cast<cudaStream_t>(stream), 
      weights, dweights, gram, H, seg_len, seg_valid, lengths, 4.0f * penalty_coef / static_cast<float>(B), scale);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_att_scores_bwd(const void* key, const float* qpad, const float* dscores, void* dkey, float* dqpad, int B,
                                 int seg_len, int seg_valid, const int32_t* lengths, int H, int ldk, int accumulate,
                                 void* stream) {
  if (!key || !qpad || !dscores || !dkey || !dqpad || B <= 0 || seg_len <= 0 || ldk <= 0 || ldk % 8)
    return set_error(XV_ERR_INVALID, "xv_att_scores_bwd: bad arguments");
  int rc = att_check_heads("xv_att_scores_bwd", H); if (rc) return rc;
  const long long rows = static_cast<long long>(B) * seg_len;
  if (rows > 0x7fffffffLL) return set_error(XV_ERR_INVALID, "xv_att_scores_bwd: rows must fit in int32");
  const int ht = H <= 1 ? 1 : (H <= 2 ? 2 : (H <= 4 ? 4 : (H <= 8 ? 8 : 16)));
  const size_t smem = static_cast<size_t>(8) * ht * 256 * sizeof(float);
  dim3 grid(ceil_div(ldk, 256), ceil_div(rows, ASB_ROWS));
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
#define XV_ASB_LAUNCH(HT_)                                                                                                   \
  do {                                                                                                                       \
    if (smem > 48 * 1024) {                                                                                                  \
      static bool configured = false;                                                                                       \
      if (!configured) {                                                                                                     \
        XV_CUDA_CHECK(cudaFuncSetAttribute(att_scores_bwd_kernel<HT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                           static_cast<int>(smem)));                                                         \
        configured = true;                                                                                                   \
      }                                                                                                                      \
    }                                                                                                                        \
    ::xv::launch_pdl((att_scores_bwd_kernel<HT_>), grid, 256, smem, s_, static_cast<const __nv_bfloat16*>(key), qpad, dscores,  \
                     static_cast<__nv_bfloat16*>(dkey), dqpad, static_cast<int>(rows), seg_len, seg_valid, lengths, H, ldk,  \
                     accumulate);                                                                                            \
  } while (0)
  switch (ht) {
    case 1: XV_ASB_LAUNCH(1); break;
    case 2: XV_ASB_LAUNCH(2); break;
    case 4: XV_ASB_LAUNCH(4); break;
    case 8: XV_ASB_LAUNCH(8); break;
    default: XV_ASB_LAUNCH(16); break;
  }
#undef XV_ASB_LAUNCH
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
