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

__device__ __forceinline__ bool row_is_valid(int m, int seg_len, int seg_valid, const int* lengths) {
  if (seg_len <= 0) return true;
  const int b = m / seg_len, t = m - b * seg_len;
  return t < (lengths ? lengths[b] : seg_valid);
}

// Gradient of the statistics pooling evaluated on the fly (fused layer-5 backward): the upstream gradient of
// act(BN(y)) at frame (b, t) is  gmean/L + 1[var > floor] * gstd * (a - mean) / (L * std)  with a recomputed from y,
// so neither tdnn5_relu nor its gradient ever round-trips HBM.
struct PoolGradSrc {
  const float* pooled;    // [B, 2*cpad] = [mean | std]   (nullptr: read the upstream gradient from memory)
  const float* dpooled;   // [B, 2*cpad]
  int cpad, c_real;
};
// ------------------------------------------------------------------------------------------------
// Input packing: features f32 [B, T, D] -> bf16 im2col rows [B*T, ldo], out[m, j*dpad + c] = x[b, t+j, c].
// One thread per (row m, 8-channel group): eight feature loads (L1/L2 hits: the 3 MB input is read k times) and one
// 16-byte store, so a warp writes 512 contiguous bytes; slots j >= k and channels c >= D are zero padding.
// (v1: four integer divisions and a 2-byte store per ELEMENT, 23 us for 13 MB; v2: one thread per 64-byte tap slot with
// 30 strided scalar loads each, 13.5 us.)
__global__ void __launch_bounds__(256) pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B,
                                                         int T, int D, int k, int dpad, long long ldo) {
  pdl_entry();
  const int groups = static_cast<int>(ldo >> 3);
  const long long total = static_cast<long long>(B) * T * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / groups;
    const int col = static_cast<int>(i - m * groups) << 3;
    const int j = col / dpad, c0 = col - j * dpad;
    const int t = static_cast<int>(m % T);
    const bool live = j < k && t + j < T;
    const float* src = x + (m + j) * D + c0;     // frame (b, t + j): rows of one segment are contiguous
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = (live && c0 + q < D) ? __ldg(src + q) : 0.f;
    store8(out + m * ldo + col, f);
  }
}

// ------------------------------------------------------------------------------------------------
// BN finalize (C threads total).
__global__ void bn_finalize_train_kernel(const float* __restrict__ col_sum, const float* __restrict__ col_sumsq,
                                         const float* __restrict__ bias, float count, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* moving_mean, float* moving_var,
                                         float momentum, float eps, int unbiased, float* scale, float* shift,
                                         float* save_mean, float* save_rstd, int C) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // y is stored WITHOUT the layer bias (a bias in front of a batch-norm cancels): scale / shift / saved mean refer to the
  // stored tensor, only the moving mean -- the statistic of the reference's biased tensor -- adds the bias back
  const float mean_nb = col_sum[c] / count;
  const float var = fmaxf(col_sumsq[c] / count - mean_nb * mean_nb, 0.f);
  const float mean = mean_nb + (bias ? bias[c] : 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean_nb * sc;
  save_mean[c] = mean_nb;
  save_rstd[c] = rstd;
  if (moving_mean) {
    const float mv = unbiased ? var * (count / fmaxf(count - 1.f, 1.f)) : var;
    moving_mean[c] = moving_mean[c] * momentum + mean * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + mv * (1.f - momentum);
  }
}
__global__ void bn_finalize_infer_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                         const float* __restrict__ mm, const float* __restrict__ mv,
                                         const float* __restrict__ bias, float eps, float* scale, float* shift, int C) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(mv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - (mm[c] - (bias ? bias[c] : 0.f)) * sc;     // the stored tensor is bias-free: fold the bias in here
}

// Row-streaming layout shared by the BN kernels below: grid = (C/128, row chunks of STREAM_ROWS); a block is 8 warps,
// a warp covers 128 channels (4 per lane, one 8-byte vector) and walks rows w, w+8, ... of its chunk two at a time.
// Every per-channel constant (scale, shift, mean, rstd, dgamma, dbeta, alpha) is loaded ONCE per thread, and four
// channels per thread keep the kernels under 64-80 registers so that 3-4 blocks (24-32 warps) stay resident per SM.
#ifndef XV_STREAM_ROWS
#define XV_STREAM_ROWS 64
#endif
#ifndef XV_RU
#define XV_RU 2
#endif
#ifndef XV_FU
#define XV_FU 4
#endif
constexpr int STREAM_ROWS = XV_STREAM_ROWS;
constexpr int SV = 4;                 // channels per thread
constexpr int SCH = 32 * SV;          // channels per warp / block column
constexpr int RU = XV_RU;               // rows a warp keeps in flight per tensor (8 in flight measured SLOWER in the step:
                                      // 125 registers -> 2 blocks/SM and no load/compute overlap inside a block)

__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&f)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ void load4f(const float* p, float (&f)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}

// A block of the row-streaming kernels owns rows [t0, t1) of ONE segment b, so row validity is a compare (t < L) and the
// per-segment pooling coefficients are loaded once per block; the only integer division is this one, per block.
// (A division per row per thread made these kernels issue-bound: ~36 warp instructions per element.)
struct RowChunk {
  int b, t0, t1, L;
  long long m_base;     // first row of the segment
};
struct RowGrid {
  int seg_len, chunks, rows_per_chunk;   // seg_len = rows when the tensor has no segment structure
  int col_groups;                        // channel groups of SCH channels; blockIdx.x = row_chunk * col_groups + group
};
// Kernels that evaluate the pooling gradient on the fly load 4 x 8 per-(segment, channel) coefficients per thread and
// block: longer row chunks amortise them (64-row chunks made the coefficient traffic equal to the data traffic).
constexpr int FUSED_STREAM_ROWS = 128;
static inline RowGrid make_row_grid(long long rows, int seg_len, int C, int group_channels = SCH,
                                    int stream_rows = STREAM_ROWS) {
  RowGrid g;
  g.col_groups = (C + group_channels - 1) / group_channels;
  g.seg_len = seg_len > 0 ? seg_len : static_cast<int>(rows);
  g.chunks = (g.seg_len + stream_rows - 1) / stream_rows;
  g.rows_per_chunk = (g.seg_len + g.chunks - 1) / g.chunks;
  return g;
}
__device__ __forceinline__ RowChunk row_chunk(int block_y, const RowGrid& g, int seg_len, int seg_valid, const int* lengths) {
  RowChunk c;
  c.b = block_y / g.chunks;
  const int ch = block_y - c.b * g.chunks;
  c.t0 = ch * g.rows_per_chunk;
  c.t1 = min(c.t0 + g.rows_per_chunk, g.seg_len);
  c.L = seg_len > 0 ? (lengths ? lengths[c.b] : seg_valid) : g.seg_len;
  c.m_base = static_cast<long long>(c.b) * g.seg_len;
  return c;
}

template <int ACT>
__device__ __forceinline__ float actf(float z, float alpha) { return act_fwd(ACT, z, alpha); }
template <int ACT>
__device__ __forceinline__ float actg(float z, float alpha) { return act_grad(ACT, z, alpha); }

// Per-(segment, channel) coefficients of the on-the-fly pooling gradient: da = ca + cb * a.
struct PoolCoef {
  float ca[SV], cb[SV];
  int b;
};
__device__ __forceinline__ void pool_coef_load(PoolCoef& pc, const PoolGradSrc& ps, int b, int c0, int seg_valid,
                                               const int* lengths) {
  if (pc.b == b) return;
  pc.b = b;
  const float invl = 1.0f / (static_cast<float>(lengths ? lengths[b] : seg_valid) + 1e-16f);
  float mu[SV], sd[SV], gm[SV], gs[SV];
  const float* pb = ps.pooled + static_cast<long long>(b) * 2 * ps.cpad;
  const float* gb = ps.dpooled + static_cast<long long>(b) * 2 * ps.cpad;
  load4f(pb + c0, mu); load4f(pb + ps.cpad + c0, sd); load4f(gb + c0, gm); load4f(gb + ps.cpad + c0, gs);
#pragma unroll
  for (int j = 0; j < SV; ++j) {
    float ca = 0.f, cb = 0.f;
    if (c0 + j < ps.c_real) {
      ca = gm[j] * invl;
      if (sd[j] * sd[j] > 1.0000001e-12f) {       // variance above the 1e-12 floor: gradient flows through std
        cb = gs[j] * invl / sd[j];
        ca -= cb * mu[j];
      }
    }
    pc.ca[j] = ca;
    pc.cb[j] = cb;
  }
}

// ---- wide row-streaming kernels (BN apply / backward): 8 channels per thread (16-byte accesses), a block column of
// WCH = 256 channels, row pointers advanced by adds.  ncu on the first 4-channel versions: ~29 issued instructions per
// element (64-bit address multiplies and row bookkeeping amortised over 4 elements), issue slots 45 % busy at 35 %
// occupancy -- the kernels were instruction-bound at ~2.5 TB/s, not memory-bound.
constexpr int WV = 8;
constexpr int WCH = 32 * WV;
// Rows in flight per warp, from the sweep in tools/layers_bench.py on B200 (C = 512 / 1536, HBM-streaming):
//   plain kernels: 1 row (17 / 25 / 39 / 58 us) beats 2 (26 / 30 / 64 / 76) and 4 -- fewer registers, more resident warps;
//   fused-pooling kernels: 4 rows (49 us) beat 2 (56) and 1 (68) -- they amortise the per-block coefficient loads.
#ifndef XV_WROWS
#define XV_WROWS 1
#endif
#ifndef XV_WROWS_FUSED
#define XV_WROWS_FUSED 4
#endif

struct WideBlock {
  int c0, w, lane, cgroup;
  RowChunk rc;
};
__device__ __forceinline__ WideBlock wide_block(const RowGrid& rg, int seg_len, int seg_valid, const int* lengths) {
  WideBlock b;
  b.lane = threadIdx.x & 31;
  b.w = threadIdx.x >> 5;
  b.cgroup = blockIdx.x % rg.col_groups;
  b.c0 = b.cgroup * WCH + b.lane * WV;
  b.rc = row_chunk(blockIdx.x / rg.col_groups, rg, seg_len, seg_valid, lengths);
  return b;
}

// Training-mode BN finalisation folded into the apply kernel (col_sum != nullptr): every block derives scale / shift of
// its 8 channels per thread from the batch statistics the GEMM epilogue accumulated; the blocks of the first row
// chunk also publish scale / shift / saved mean / rstd for the backward pass and update the moving statistics
// (what xv_bn_finalize_train does as a separate 3 us launch per layer).
struct BnTrainSrc {
  const float* col_sum;
  const float* col_sumsq;
  const float* bias;
  const float* gamma;
  const float* beta;
  float* moving_mean;
  float* moving_var;
  float* scale_out;
  float* shift_out;
  float* save_mean;
  float* save_rstd;
  float count, momentum, eps;
  int unbiased;
};

// a = act(y*scale + shift) on valid rows, 0 on invalid rows.
template <int ACT>
__global__ void __launch_bounds__(256) bn_act_apply_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ a,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ alpha, RowGrid rg, int C, long long ld,
                                    int seg_len, int seg_valid, const int* __restrict__ lengths, BnTrainSrc bt) {
  pdl_entry();
  __shared__ float s_sc[WCH], s_sh[WCH];
  const WideBlock wb = wide_block(rg, seg_len, seg_valid, lengths);
  const int c0 = wb.c0;
  if (bt.col_sum != nullptr) {      // one channel per thread, shared through smem (not 8 per thread: 2x the kernel time)
    const int c = wb.cgroup * WCH + threadIdx.x;
    if (c < C) {
      const float mean_nb = bt.col_sum[c] / bt.count;                       // mean of the stored (bias-free) tensor
      const float var = fmaxf(bt.col_sumsq[c] / bt.count - mean_nb * mean_nb, 0.f);
      const float mean = mean_nb + (bt.bias ? bt.bias[c] : 0.f);            // moving mean only (see bn_finalize_train_kernel)
      const float rstd = rsqrtf(var + bt.eps);
      const float scv = bt.gamma[c] * rstd;
      const float shv = bt.beta[c] - mean_nb * scv;
      s_sc[threadIdx.x] = scv;
      s_sh[threadIdx.x] = shv;
      if ((blockIdx.x / rg.col_groups) == 0) {     // first row chunk: publish for the backward pass, update moving stats
        bt.scale_out[c] = scv;
        bt.shift_out[c] = shv;
        bt.save_mean[c] = mean_nb;
        bt.save_rstd[c] = rstd;
        if (bt.moving_mean) {
          const float mv = bt.unbiased ? var * (bt.count / fmaxf(bt.count - 1.f, 1.f)) : var;
          bt.moving_mean[c] = bt.moving_mean[c] * bt.momentum + mean * (1.f - bt.momentum);
          bt.moving_var[c] = bt.moving_var[c] * bt.momentum + mv * (1.f - bt.momentum);
        }
      }
    }
    __syncthreads();
  }
  if (c0 >= C) return;
  const RowChunk rc = wb.rc;
  float sc[WV], sh[WV], al[WV];
  if (bt.col_sum != nullptr) {
#pragma unroll
    for (int j = 0; j < WV; ++j) { sc[j] = s_sc[wb.lane * WV + j]; sh[j] = s_sh[wb.lane * WV + j]; }
  } else {
    load8f(scale + c0, sc);
    load8f(shift + c0, sh);
  }
#pragma unroll
  for (int j = 0; j < WV; ++j) al[j] = (ACT == ACT_PRELU) ? alpha[c0 + j] : 0.f;
  const long long step = 8 * ld;
  const __nv_bfloat16* yp = y + (rc.m_base + rc.t0 + wb.w) * ld + c0;
  __nv_bfloat16* ap = a + (rc.m_base + rc.t0 + wb.w) * ld + c0;
  constexpr int WROWS = XV_WROWS;
  for (int t = rc.t0 + wb.w; t < rc.t1; t += 8 * WROWS, yp += WROWS * step, ap += WROWS * step) {
    float v[WROWS][WV];
    bool in[WROWS], valid[WROWS];
#pragma unroll
    for (int u = 0; u < WROWS; ++u) {
      const int tt = t + 8 * u;
      in[u] = tt < rc.t1;
      valid[u] = tt < rc.t1 && tt < rc.L;
      if (valid[u]) load8(yp + u * step, v[u]);
    }
#pragma unroll
    for (int u = 0; u < WROWS; ++u) {
      if (!in[u]) continue;
      float o[WV];
#pragma unroll
      for (int j = 0; j < WV; ++j) o[j] = valid[u] ? actf<ACT>(fmaf(v[u][j], sc[j], sh[j]), al[j]) : 0.f;
      store8(ap + u * step, o);
    }
  }
}

// Per-channel sum / sum of squares of (y - bias) over the valid rows: the BN batch statistics of layers whose GEMM is
// too short (K <= 512) to hide a reduction epilogue.
__global__ void __launch_bounds__(256) col_stats_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ bias,
                                                        RowGrid rg, int C, long long ld, int seg_len, int seg_valid,
                                                        const int* __restrict__ lengths, float* col_sum, float* col_sumsq) {
  pdl_entry();
  __shared__ float red[8][2][SCH];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cg_ = blockIdx.x % rg.col_groups, by_ = blockIdx.x / rg.col_groups;
  const int c0 = cg_ * SCH + lane * SV;
  float s1[SV], s2[SV];
#pragma unroll
  for (int j = 0; j < SV; ++j) s1[j] = s2[j] = 0.f;
  if (c0 < C) {
    float bs[SV];
#pragma unroll
    for (int j = 0; j < SV; ++j) bs[j] = bias ? bias[c0 + j] : 0.f;
    const RowChunk rc = row_chunk(by_, rg, seg_len, seg_valid, lengths);
    const int t1 = min(rc.t1, rc.L);
    for (int t = rc.t0 + w; t < t1; t += 8 * RU) {
      float v[RU][SV];
      bool ok[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const int tt = t + 8 * u;
        ok[u] = tt < t1;
        if (ok[u]) load4(y + (rc.m_base + tt) * ld + c0, v[u]);
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        if (!ok[u]) continue;
#pragma unroll
        for (int j = 0; j < SV; ++j) {
          const float d = v[u][j] - bs[j];
          s1[j] += d;
          s2[j] = fmaf(d, d, s2[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < SV; ++j) { red[w][0][lane * SV + j] = s1[j]; red[w][1][lane * SV + j] = s2[j]; }
  __syncthreads();
  const int c = cg_ * SCH + threadIdx.x;
  if (threadIdx.x < SCH && c < C) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[k][0][threadIdx.x]; q += red[k][1][threadIdx.x]; }
    atomicAdd(col_sum + c, a);
    atomicAdd(col_sumsq + c, q);
  }
}

// Per-(segment, channel) coefficients of the on-the-fly pooling gradient, 8 channels: da = ca + cb * a.
struct PoolCoef8 { float ca[WV], cb[WV]; };
__device__ __forceinline__ void pool_coef_load8(PoolCoef8& pc, const PoolGradSrc& ps, int b, int c0, int L) {
  const float invl = 1.0f / (static_cast<float>(L) + 1e-16f);
  float mu[WV], sd[WV], gm[WV], gs[WV];
  const float* pb = ps.pooled + static_cast<long long>(b) * 2 * ps.cpad;
  const float* gb = ps.dpooled + static_cast<long long>(b) * 2 * ps.cpad;
  load8f(pb + c0, mu); load8f(pb + ps.cpad + c0, sd); load8f(gb + c0, gm); load8f(gb + ps.cpad + c0, gs);
#pragma unroll
  for (int j = 0; j < WV; ++j) {
    float ca = 0.f, cb = 0.f;
    if (c0 + j < ps.c_real) {
      ca = gm[j] * invl;
      if (sd[j] * sd[j] > 1.0000001e-12f) {       // variance above the 1e-12 floor: gradient flows through std
        cb = gs[j] * invl / sd[j];
        ca -= cb * mu[j];
      }
    }
    pc.ca[j] = ca;
    pc.cb[j] = cb;
  }
}

// Column reductions for the BN backward: dbeta += sum g, dgamma += sum g*yhat, dalpha += sum da*min(z,0).
template <bool FUSED_POOL, int ACT>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ da, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
    const float* __restrict__ alpha, RowGrid rg, int C, long long ld, int seg_len, int seg_valid,
    const int* __restrict__ lengths, float* dgamma, float* dbeta, float* dalpha, PoolGradSrc ps) {
  pdl_entry();
  constexpr int WROWS = FUSED_POOL ? XV_WROWS_FUSED : XV_WROWS;
  __shared__ float red[8][WCH];
  const WideBlock wb = wide_block(rg, seg_len, seg_valid, lengths);
  const int c0 = wb.c0, lane = wb.lane, w = wb.w;
  const bool c_ok = c0 < C;
  float sg[WV], sgy[WV], sal[WV];
#pragma unroll
  for (int j = 0; j < WV; ++j) sg[j] = sgy[j] = sal[j] = 0.f;
  if (c_ok) {
    const RowChunk rc = wb.rc;
    float sc[WV], sh[WV], mu[WV], al[WV];
    load8f(scale + c0, sc); load8f(shift + c0, sh); load8f(save_mean + c0, mu);
#pragma unroll
    for (int j = 0; j < WV; ++j) al[j] = (ACT == ACT_PRELU) ? alpha[c0 + j] : 0.f;
    PoolCoef8 pc;
    if (FUSED_POOL) pool_coef_load8(pc, ps, rc.b, c0, rc.L);
    const int t1 = min(rc.t1, rc.L);
    const long long step = 8 * ld;
    const long long off = (rc.m_base + rc.t0 + w) * ld + c0;
    const __nv_bfloat16* yp = y + off;
    const __nv_bfloat16* dp = FUSED_POOL ? nullptr : da + off;
    for (int t = rc.t0 + w; t < t1; t += 8 * WROWS, yp += WROWS * step) {
      float v[WROWS][WV], d[WROWS][WV];
      bool ok[WROWS];
#pragma unroll
      for (int u = 0; u < WROWS; ++u) {
        ok[u] = t + 8 * u < t1;
        if (ok[u]) {
          load8(yp + u * step, v[u]);
          if (!FUSED_POOL) load8(dp + u * step, d[u]);
        }
      }
      if (!FUSED_POOL) dp += WROWS * step;
#pragma unroll
      for (int u = 0; u < WROWS; ++u) {
        if (!ok[u]) continue;
#pragma unroll
        for (int j = 0; j < WV; ++j) {
          const float z = fmaf(v[u][j], sc[j], sh[j]);
          const float dd = FUSED_POOL ? fmaf(pc.cb[j], actf<ACT>(z, al[j]), pc.ca[j]) : d[u][j];
          const float g = dd * actg<ACT>(z, al[j]);
          sg[j] += g;
          sgy[j] = fmaf(g, v[u][j] - mu[j], sgy[j]);      // * rstd once, below
          if (ACT == ACT_PRELU) sal[j] = fmaf(dd, fminf(z, 0.f), sal[j]);
        }
      }
    }
    float rs[WV];
    load8f(save_rstd + c0, rs);
#pragma unroll
    for (int j = 0; j < WV; ++j) sgy[j] *= rs[j];
  }
  // three passes through one [8][WCH] staging tile: dbeta, dgamma, (dalpha)
#pragma unroll
  for (int pass = 0; pass < (ACT == ACT_PRELU ? 3 : 2); ++pass) {
    if (pass) __syncthreads();
#pragma unroll
    for (int j = 0; j < WV; ++j) red[w][lane * WV + j] = pass == 0 ? sg[j] : (pass == 1 ? sgy[j] : sal[j]);
    __syncthreads();
    const int c = wb.cgroup * WCH + threadIdx.x;
    if (c < C) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[k][threadIdx.x];
      if (pass == 0) atomicAdd(dbeta + c, a);
      else if (pass == 1) atomicAdd(dgamma + c, a);
      else if (dalpha) atomicAdd(dalpha + c, a);
    }
  }
}

// dy = scale * (g - dbeta/n - yhat*dgamma/n) on valid rows, 0 elsewhere (scale = gamma*rstd), evaluated as
// dy = scale*g + A*y + Bc with A = -scale*rstd*dgamma/n and Bc = -scale*dbeta/n - A*mean (per-channel constants).
template <bool FUSED_POOL, int ACT>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ dy,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ save_mean,
    const float* __restrict__ save_rstd, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
    float inv_count, const float* __restrict__ alpha, RowGrid rg, int C, long long ld, int seg_len,
    int seg_valid, const int* __restrict__ lengths, PoolGradSrc ps) {
  pdl_entry();
  constexpr int WROWS = FUSED_POOL ? XV_WROWS_FUSED : XV_WROWS;
  const WideBlock wb = wide_block(rg, seg_len, seg_valid, lengths);
  const int c0 = wb.c0;
  if (c0 >= C) return;
  const RowChunk rc = wb.rc;
  float sc[WV], sh[WV], ka[WV], kb[WV], al[WV];
  load8f(scale + c0, sc); load8f(shift + c0, sh);
  {
    float mu[WV], rs[WV], dg[WV], db[WV];
    load8f(save_mean + c0, mu); load8f(save_rstd + c0, rs); load8f(dgamma + c0, dg); load8f(dbeta + c0, db);
#pragma unroll
    for (int j = 0; j < WV; ++j) {
      ka[j] = -sc[j] * rs[j] * dg[j] * inv_count;
      kb[j] = -sc[j] * db[j] * inv_count - ka[j] * mu[j];
      al[j] = (ACT == ACT_PRELU) ? alpha[c0 + j] : 0.f;
    }
  }
  PoolCoef8 pc;
  if (FUSED_POOL) pool_coef_load8(pc, ps, rc.b, c0, rc.L);
  const long long step = 8 * ld;
  const long long off = (rc.m_base + rc.t0 + wb.w) * ld + c0;
  const __nv_bfloat16* yp = y + off;
  const __nv_bfloat16* dp = FUSED_POOL ? nullptr : da + off;
  __nv_bfloat16* op = dy + off;
  for (int t = rc.t0 + wb.w; t < rc.t1; t += 8 * WROWS, yp += WROWS * step, op += WROWS * step) {
    float v[WROWS][WV], d[WROWS][WV];
    bool in[WROWS], valid[WROWS];
#pragma unroll
    for (int u = 0; u < WROWS; ++u) {
      const int tt = t + 8 * u;
      in[u] = tt < rc.t1;
      valid[u] = tt < rc.t1 && tt < rc.L;
      if (valid[u]) {
        load8(yp + u * step, v[u]);
        if (!FUSED_POOL) load8(dp + u * step, d[u]);
      }
    }
    if (!FUSED_POOL) dp += WROWS * step;
#pragma unroll
    for (int u = 0; u < WROWS; ++u) {
      if (!in[u]) continue;
      float o[WV];
      if (valid[u]) {
#pragma unroll
        for (int j = 0; j < WV; ++j) {
          const float z = fmaf(v[u][j], sc[j], sh[j]);
          const float dd = FUSED_POOL ? fmaf(pc.cb[j], actf<ACT>(z, al[j]), pc.ca[j]) : d[u][j];
          const float g = dd * actg<ACT>(z, al[j]);
          o[j] = fmaf(sc[j], g, fmaf(ka[j], v[u][j], kb[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < WV; ++j) o[j] = 0.f;
      }
      store8(op + u * step, o);
    }
  }
}

// ReLU specialisation of the fused tdnn5 backward (the pooling gradient evaluated on the fly): the upstream gradient is
// affine in the activation, da = ca + cb*a per (segment, channel), and with ReLU a = z = scale*y + shift on the positive
// set, so dy = scale*g + ka*y + kb collapses to ONE fma with per-(segment, channel) constants chosen by the sign of z:
//   z > 0 :  dy = (scale^2 cb + ka) y + (scale (ca + cb shift) + kb)          z <= 0 :  dy = ka y + kb.
// Rows stay packed in registers (16 bytes each) until they are consumed:
// 5 instructions per element instead of 7.
__global__ void __launch_bounds__(256, 2) bn_act_bwd_apply_pool_relu_kernel(
    const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dy, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
    const float* __restrict__ dgamma, const float* __restrict__ dbeta, float inv_count, RowGrid rg, int C, long long ld,
    int seg_len, int seg_valid, const int* __restrict__ lengths, PoolGradSrc ps) {
  pdl_entry();
  constexpr int WROWS = XV_WROWS_FUSED;
  const WideBlock wb = wide_block(rg, seg_len, seg_valid, lengths);
  const int c0 = wb.c0;
  if (c0 >= C) return;
  const RowChunk rc = wb.rc;
  float sc[WV], sh[WV], a1[WV], b1[WV], a0[WV], b0[WV];
  load8f(scale + c0, sc); load8f(shift + c0, sh);
  {
    float mu[WV], rs[WV], dg[WV], db[WV];
    load8f(save_mean + c0, mu); load8f(save_rstd + c0, rs); load8f(dgamma + c0, dg); load8f(dbeta + c0, db);
    PoolCoef8 pc;
    pool_coef_load8(pc, ps, rc.b, c0, rc.L);
#pragma unroll
    for (int j = 0; j < WV; ++j) {
      a0[j] = -sc[j] * rs[j] * dg[j] * inv_count;
      b0[j] = -sc[j] * db[j] * inv_count - a0[j] * mu[j];
      a1[j] = fmaf(sc[j] * sc[j], pc.cb[j], a0[j]);
      b1[j] = fmaf(sc[j], fmaf(pc.cb[j], sh[j], pc.ca[j]), b0[j]);
    }
  }
  const long long step = 8 * ld;
  const long long off = (rc.m_base + rc.t0 + wb.w) * ld + c0;
  const __nv_bfloat16* yp = y + off;
  __nv_bfloat16* op = dy + off;
  for (int t = rc.t0 + wb.w; t < rc.t1; t += 8 * WROWS, yp += WROWS * step, op += WROWS * step) {
    uint4 raw[WROWS];
    bool in[WROWS], valid[WROWS];
#pragma unroll
    for (int u = 0; u < WROWS; ++u) {
      const int tt = t + 8 * u;
      in[u] = tt < rc.t1;
      valid[u] = tt < rc.t1 && tt < rc.L;
      if (valid[u]) raw[u] = *reinterpret_cast<const uint4*>(yp + u * step);
    }
#pragma unroll
    for (int u = 0; u < WROWS; ++u) {
      if (!in[u]) continue;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (valid[u]) {
        const uint32_t wds[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
        uint32_t ow[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float y0 = __uint_as_float(wds[q] << 16), y1 = __uint_as_float(wds[q] & 0xffff0000u);
          const bool p0 = fmaf(y0, sc[2 * q], sh[2 * q]) > 0.f, p1 = fmaf(y1, sc[2 * q + 1], sh[2 * q + 1]) > 0.f;
          const float r0 = fmaf(p0 ? a1[2 * q] : a0[2 * q], y0, p0 ? b1[2 * q] : b0[2 * q]);
          const float r1 = fmaf(p1 ? a1[2 * q + 1] : a0[2 * q + 1], y1, p1 ? b1[2 * q + 1] : b0[2 * q + 1]);
          const __nv_bfloat162 pk = __floats2bfloat162_rn(r0, r1);
          ow[q] = *reinterpret_cast<const uint32_t*>(&pk);
        }
        o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      *reinterpret_cast<uint4*>(op + u * step) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Statistics pooling (model/pooling.py:22-32; masked form multitask_v1/pooling.py:22-38).
// grid = (Cpad/256, B); block = 256 = 8 warps striding over time; shifted one-pass moments
// (shift = first frame) so that var = E[(x-x0)^2] - E[x-x0]^2 does not cancel catastrophically.
//
// Training with the fused tdnn5 BN + activation also emits, per (segment, channel), the four sums the BN backward
// needs.  The pooling gradient is affine in the activation, da_t = ca + cb * a_t (PoolCoef), so with g' = act'(z):
//   dbeta  = sum_{b,t} da g'          = sum_b ca S1 + cb S2,   S1 = sum_t g',      S2 = sum_t g' a
//   dgamma = sum_{b,t} da g' yhat     = sum_b ca S3 + cb S4,   S3 = sum_t g' yhat, S4 = sum_t g' a yhat
// which turns the backward column-reduction pass over the largest activation of the network (78 MB) into a
// [B, C]-sized kernel (pool_bn_bwd_reduce_kernel).
#ifndef XV_POOL_ROWS
#define XV_POOL_ROWS 1
#endif
// frames in flight per warp of the training variant (backward sums): 4 -> 36.4 us, 1 -> 47.8 us, 2 -> 51.6 us isolated at
// config 2 (the kernel is issue-bound: deeper unrolling amortises the loop / predicate overhead); the plain variant is the
// other way round (27.0 us at 1, 50.5 us at 4)
#ifndef XV_POOL_ROWS_BWD
#define XV_POOL_ROWS_BWD 4
#endif
#ifndef XV_POOL_CPT
#define XV_POOL_CPT 4
#endif
template <int N>
__device__ __forceinline__ void loadN(const __nv_bfloat16* p, float (&f)[N]);
template <>
__device__ __forceinline__ void loadN<8>(const __nv_bfloat16* p, float (&f)[8]) { load8(p, f); }
template <>
__device__ __forceinline__ void loadN<4>(const __nv_bfloat16* p, float (&f)[4]) { load4(p, f); }
template <int N>
__device__ __forceinline__ void loadNf(const float* p, float (&f)[N]);
template <>
__device__ __forceinline__ void loadNf<8>(const float* p, float (&f)[8]) { load8f(p, f); }
template <>
__device__ __forceinline__ void loadNf<4>(const float* p, float (&f)[4]) { load4f(p, f); }

// CPT = channels per thread (8: 16-byte loads; 4 when the extra sums would push the kernel past 128 registers).
template <bool BWD_SUMS, int CPT, int ACT>
__global__ void __launch_bounds__(256) stats_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                             __nv_bfloat16* __restrict__ out3, int seg_len,
                                                             int seg_valid, const int* __restrict__ lengths,
                                                             int c_real, int cpad, long long ld,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             const float* __restrict__ alpha,
                                                             const float* __restrict__ save_mean,
                                                             const float* __restrict__ save_rstd,
                                                             float* __restrict__ bwd_sums,
                                                             const int* __restrict__ starts) {
  pdl_entry();
  // scale != nullptr: x is the PRE-BN tensor and the pooled quantity is act(x*scale + shift) (fused tdnn5 BN+ReLU)
  constexpr int BC = 32 * CPT;     // channels per block
  __shared__ float red[8][2][BC];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * BC + lane * CPT;
  const int L = lengths ? lengths[b] : seg_valid;
  // starts != nullptr: ragged layout -- segment b occupies rows [starts[b], starts[b] + L) of one flat row space
  const __nv_bfloat16* xb = x + (starts ? static_cast<long long>(starts[b]) : static_cast<long long>(b) * seg_len) * ld;
  float s1[CPT], s2[CPT], x0[CPT];
  float q1[CPT], q2[CPT], q3[CPT], q4[CPT];     // dead code (and registers) when !BWD_SUMS
#pragma unroll
  for (int j = 0; j < CPT; ++j) s1[j] = s2[j] = x0[j] = 0.f;
  if (BWD_SUMS) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) q1[j] = q2[j] = q3[j] = q4[j] = 0.f;
  }
  if (c0 < cpad && L > 0) {
    float sc[CPT], sh[CPT], al[CPT], mu[CPT], rs[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { sc[j] = 1.f; sh[j] = 0.f; al[j] = 0.f; }
    if (scale) {
      loadNf<CPT>(scale + c0, sc);
      loadNf<CPT>(shift + c0, sh);
      if (ACT == ACT_PRELU) loadNf<CPT>(alpha + c0, al);
    }
    if (BWD_SUMS) { loadNf<CPT>(save_mean + c0, mu); loadNf<CPT>(save_rstd + c0, rs); }
    loadN<CPT>(xb + c0, x0);
    if (scale) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) x0[j] = actf<ACT>(fmaf(x0[j], sc[j], sh[j]), al[j]);
    }
    constexpr int PR = BWD_SUMS ? XV_POOL_ROWS_BWD : XV_POOL_ROWS;       // frames in flight per warp
    const long long step = 8 * ld;
    const __nv_bfloat16* xp = xb + static_cast<long long>(w) * ld + c0;
    for (int t = w; t < L; t += 8 * PR, xp += PR * step) {
      float v[PR][CPT];
      bool ok[PR];
#pragma unroll
      for (int u = 0; u < PR; ++u) {
        ok[u] = t + 8 * u < L;
        if (ok[u]) loadN<CPT>(xp + u * step, v[u]);
      }
#pragma unroll
      for (int u = 0; u < PR; ++u) {
        if (!ok[u]) continue;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float z = fmaf(v[u][j], sc[j], sh[j]);
          const float a = scale ? actf<ACT>(z, al[j]) : v[u][j];
          const float d = a - x0[j];
          s1[j] += d;
          s2[j] = fmaf(d, d, s2[j]);
          if (BWD_SUMS) {       // yhat = (y - mean) * rstd: the rstd factor is applied once, at the end
            const float gp = actg<ACT>(z, al[j]);
            const float ym = v[u][j] - mu[j];
            q1[j] += gp;
            q3[j] = fmaf(gp, ym, q3[j]);
            if (ACT == ACT_RELU) {       // g' a = a for relu: S2 = sum_t a comes from the pooling sum itself (below)
              q4[j] = fmaf(a, ym, q4[j]);
            } else {
              const float ga = gp * a;
              q2[j] += ga;
              q4[j] = fmaf(ga, ym, q4[j]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j) { red[w][0][lane * CPT + j] = s1[j]; red[w][1][lane * CPT + j] = s2[j]; }
  __syncthreads();
  const int c = blockIdx.x * BC + threadIdx.x;
  const bool c_own = threadIdx.x < BC && c < cpad;
  float sum_a = 0.f;        // sum over the valid frames of the pooled activation (= S2 of the backward sums for relu)
  if (c_own) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += red[k][0][threadIdx.x]; q += red[k][1][threadIdx.x]; }
    float mean = 0.f, sd = 0.f;
    if (c < c_real && L > 0) {
      float first = __bfloat162float(xb[c]);
      if (scale) first = actf<ACT>(fmaf(first, scale[c], shift[c]), ACT == ACT_PRELU ? alpha[c] : 0.f);
      sum_a = fmaf(static_cast<float>(L), first, a);
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
  if (BWD_SUMS) {
    float* sb = bwd_sums + static_cast<long long>(b) * 4 * cpad;
#pragma unroll
    for (int round = 0; round < 2; ++round) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        red[w][0][lane * CPT + j] = round ? q3[j] : q1[j];
        red[w][1][lane * CPT + j] = round ? q4[j] : q2[j];
      }
      __syncthreads();
      if (c_own) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += red[k][0][threadIdx.x]; q += red[k][1][threadIdx.x]; }
        if (ACT == ACT_RELU && round == 0) q = sum_a;
        const float f = round ? save_rstd[c] : 1.0f;
        sb[(2 * round) * cpad + c] = (c < c_real) ? a * f : 0.f;
        sb[(2 * round + 1) * cpad + c] = (c < c_real) ? q * f : 0.f;
      }
    }
  }
}

// ReLU specialisation of the training variant (fused tdnn5 BN + ReLU, pooled moments AND the four backward sums).
// With u = y - y0 (y0 = a reference point near the segment's first frame, pool_ref_point), kappa = scale, z0 = kappa*y0 + shift, the activation on
// the positive set P = {t : z_t > 0} is a = kappa*u + z0 and zero elsewhere, so THREE accumulators per channel,
//   B0 = |P|,   B1 = sum_P u,   B2 = sum_P u^2,
// give everything: sum a = kappa B1 + z0 B0; sum (a - m)^2 = kappa^2 B2 + 2 kappa d B1 + d^2 B0 + (L - B0) m^2 with d = z0 - m
// (every term is O(L var): u is measured from a typical frame, no cancellation); S1 = B0, S2 = sum a,
// S3 = rstd (B1 + e B0), S4 = rstd (kappa B2 + (kappa e + z0) B1 + z0 e B0) with e = y0 - mean.
// 7 instructions per element instead of ~12 and three accumulators instead of six, which lets a thread own 8 channels
// (16-byte loads, 512 contiguous bytes per warp) under 96 registers.
// Reference point of the shifted moments: the segment's first frame if it is active, else the point on the ReLU kink
// (z = 0) -- deviations are then measured from a TYPICAL activation value (the first frame's, or 0), which keeps the
// variance free of cancellation exactly like the generic kernel's shift by the first frame's activation.
__device__ __forceinline__ float pool_ref_point(float y_first, float kappa, float shift) {
  const float z = fmaf(y_first, kappa, shift);
  return (z > 0.f || kappa == 0.f) ? y_first : y_first - z / kappa;
}

__global__ void __launch_bounds__(256, 3) stats_pool_fwd_relu_sums_kernel(
    const __nv_bfloat16* __restrict__ x, float* __restrict__ out, __nv_bfloat16* __restrict__ out3, int seg_len, int seg_valid,
    const int* __restrict__ lengths, int c_real, int cpad, long long ld, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
    float* __restrict__ bwd_sums) {
  pdl_entry();
  constexpr int CPT = 8, BC = 32 * CPT, PR = 4;
  __shared__ float red[8][3][BC];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * BC + lane * CPT;
  const int L = lengths ? lengths[b] : seg_valid;
  const __nv_bfloat16* xb = x + static_cast<long long>(b) * seg_len * ld;
  float b0[CPT], b1[CPT], b2[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) b0[j] = b1[j] = b2[j] = 0.f;
  if (c0 < cpad && L > 0) {
    float sc[CPT], sh[CPT], y0[CPT];
    load8f(scale + c0, sc);
    load8f(shift + c0, sh);
    load8(xb + c0, y0);
#pragma unroll
    for (int j = 0; j < CPT; ++j) y0[j] = pool_ref_point(y0[j], sc[j], sh[j]);
    const long long step = 8 * ld;
    const __nv_bfloat16* xp = xb + static_cast<long long>(w) * ld + c0;
    for (int t = w; t < L; t += 8 * PR, xp += PR * step) {
      uint4 raw[PR];               // rows stay packed (4 registers each) until they are consumed
      bool ok[PR];
#pragma unroll
      for (int u = 0; u < PR; ++u) {
        ok[u] = t + 8 * u < L;
        if (ok[u]) raw[u] = *reinterpret_cast<const uint4*>(xp + u * step);
      }
#pragma unroll
      for (int u = 0; u < PR; ++u) {
        if (!ok[u]) continue;
        const uint32_t wds[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float yv = __uint_as_float((j & 1) ? (wds[j >> 1] & 0xffff0000u) : (wds[j >> 1] << 16));
          const float d = yv - y0[j];
          if (fmaf(yv, sc[j], sh[j]) > 0.f) {
            b0[j] += 1.0f;
            b1[j] += d;
            b2[j] = fmaf(d, d, b2[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    red[w][0][lane * CPT + j] = b0[j];
    red[w][1][lane * CPT + j] = b1[j];
    red[w][2][lane * CPT + j] = b2[j];
  }
  __syncthreads();
  const int c = blockIdx.x * BC + threadIdx.x;
  if (c >= cpad) return;
  float n0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { n0 += red[k][0][threadIdx.x]; s1 += red[k][1][threadIdx.x]; s2 += red[k][2][threadIdx.x]; }
  float mean = 0.f, sd = 0.f, S1 = 0.f, S2 = 0.f, S3 = 0.f, S4 = 0.f;
  if (c < c_real && L > 0) {
    const float kap = scale[c];
    const float y0c = pool_ref_point(__bfloat162float(xb[c]), kap, shift[c]);
    const float z0 = fmaf(y0c, kap, shift[c]);
    const float fl = static_cast<float>(L);
    const float invl = 1.0f / (fl + 1e-16f);
    const float sum_a = fmaf(kap, s1, z0 * n0);
    mean = sum_a * invl;
    const float d = z0 - mean;
    const float ssq = kap * kap * s2 + 2.f * kap * d * s1 + d * d * n0 + (fl - n0) * mean * mean;
    float var = ssq * invl;
    var = (var <= 1e-12f) ? 1e-12f : var;       // VAR2STD_EPSILON floor (mask blend, pooling.py:28-29)
    sd = sqrtf(var);
    const float e = y0c - save_mean[c], rs = save_rstd[c];
    S1 = n0;
    S2 = sum_a;
    S3 = rs * fmaf(e, n0, s1);
    S4 = rs * (kap * s2 + fmaf(kap, e, z0) * s1 + z0 * e * n0);
  }
  float* ob = out + static_cast<long long>(b) * 2 * cpad;
  ob[c] = mean;
  ob[cpad + c] = sd;
  if (out3) {   // [hi | hi | lo] split copy: operand of the tdnn6 GEMM (K = 3 * 2*cpad)
    __nv_bfloat16* o3 = out3 + static_cast<long long>(b) * 6 * cpad;
    const __nv_bfloat16 mh = __float2bfloat16(mean), shh = __float2bfloat16(sd);
    o3[c] = mh; o3[cpad + c] = shh;
    o3[2 * cpad + c] = mh; o3[3 * cpad + c] = shh;
    o3[4 * cpad + c] = __float2bfloat16(mean - __bfloat162float(mh));
    o3[5 * cpad + c] = __float2bfloat16(sd - __bfloat162float(shh));
  }
  float* sb = bwd_sums + static_cast<long long>(b) * 4 * cpad;
  sb[c] = S1;
  sb[cpad + c] = S2;
  sb[2 * cpad + c] = S3;
  sb[3 * cpad + c] = S4;
}

// BN backward reductions of the layer feeding the statistics pooling, from the per-(segment, channel) sums above:
// grid = (cpad/256, segment groups); one thread per channel walks its group's segments, one atomic pair per thread.
__global__ void __launch_bounds__(256) pool_bn_bwd_reduce_kernel(const float* __restrict__ pooled,
                                                                 const float* __restrict__ dpooled,
                                                                 const float* __restrict__ sums, int B, int seg_valid,
                                                                 const int* __restrict__ lengths, int c_real, int cpad,
                                                                 float* dgamma, float* dbeta) {
  pdl_entry();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= c_real) return;
  const int per = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(b0 + per, B);
  float db = 0.f, dg = 0.f;
  for (int b = b0; b < b1; ++b) {
    const float invl = 1.0f / (static_cast<float>(lengths ? lengths[b] : seg_valid) + 1e-16f);
    const float* pb = pooled + static_cast<long long>(b) * 2 * cpad;
    const float* gb = dpooled + static_cast<long long>(b) * 2 * cpad;
    const float* sb = sums + static_cast<long long>(b) * 4 * cpad;
    const float mu = pb[c], sd = pb[cpad + c];
    float ca = gb[c] * invl, cb = 0.f;
    if (sd * sd > 1.0000001e-12f) {       // same rule as pool_coef_load
      cb = gb[cpad + c] * invl / sd;
      ca -= cb * mu;
    }
    db += ca * sb[c] + cb * sb[cpad + c];
    dg += ca * sb[2 * cpad + c] + cb * sb[3 * cpad + c];
  }
  atomicAdd(dbeta + c, db);
  atomicAdd(dgamma + c, dg);
}

// dx_t = gmean/L + 1[var>floor] * gstd * (x_t - mean) / (L * std) on valid frames, 0 elsewhere.
__global__ void stats_pool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pooled,
                                      const float* __restrict__ dpooled, __nv_bfloat16* __restrict__ dx, int B,
                                      int seg_len, int seg_valid, const int* __restrict__ lengths, int c_real, int cpad,
                                      long long ld) {
  pdl_entry();
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

#define XV_ACT_DISPATCH(ACTV, EXPR)                               \
  switch (ACTV) {                                                  \
    case ACT_NONE: { constexpr int A_ = ACT_NONE; EXPR; } break;   \
    case ACT_RELU: { constexpr int A_ = ACT_RELU; EXPR; } break;   \
    case ACT_LRELU: { constexpr int A_ = ACT_LRELU; EXPR; } break; \
    case ACT_PRELU: { constexpr int A_ = ACT_PRELU; EXPR; } break; \
    default: { constexpr int A_ = ACT_TANH; EXPR; } break;         \
  }

static inline int grid_for(long long work_items, int block, int sms) {
  long long g = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(sms) * 16;   // grid-stride beyond 16 resident-ish blocks per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}


// ------------------------------------------------------------------------------------------------
// "Flat" one-wave variants of the BN apply / BN backward-apply kernels.  The chunked kernels above launch
// (channel groups x segments x row chunks) blocks -- 1024 blocks of bn_act_apply at C = 512 against ~740 resident
// (5 per SM): 1.38 waves, the second one 38 % full -- and every block re-loads its per-channel constants for 6-8 rows
// per thread.  Here the grid is exactly the resident capacity (occupancy query x SM count): block j owns the
// contiguous row range [j * rows_per_block, ...) of the flat-time matrix for one 256-channel group, every block has the
// same number of rows (+-8), constants are loaded once, and the (segment, frame) position of a row is walked
// incrementally (one integer division per warp).
#ifndef XV_BWD_FLAT_MINBLOCKS
#define XV_BWD_FLAT_MINBLOCKS 1       // 4 (<= 64 registers, 28 bytes spilled, 4 blocks per SM): 21.3 vs 21.1 us -- no gain
#endif
struct FlatWalk {
  int b, t, L;
};
__device__ __forceinline__ FlatWalk flat_start(int m, int seg_len, int seg_valid, const int* lengths, int rows) {
  FlatWalk f;
  if (seg_len <= 0) { f.b = 0; f.t = m; f.L = rows; return f; }
  f.b = m / seg_len;
  f.t = m - f.b * seg_len;
  f.L = lengths ? lengths[f.b] : seg_valid;
  return f;
}
__device__ __forceinline__ void flat_advance(FlatWalk& f, int step, int seg_len, int seg_valid, const int* lengths) {
  f.t += step;
  if (seg_len > 0) {
    while (f.t >= seg_len) {
      f.t -= seg_len;
      ++f.b;
      f.L = lengths ? lengths[f.b] : seg_valid;      // rows beyond the matrix are never touched (m < m1 checked first)
    }
  }
}

template <int ACT, int NR>
__global__ void __launch_bounds__(256) bn_act_apply_flat_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ a,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ alpha, int rows, int rows_per_block, int col_groups, int C,
                                    long long ld, int seg_len, int seg_valid, const int* __restrict__ lengths, BnTrainSrc bt) {
  pdl_entry();
  __shared__ float s_sc[WCH], s_sh[WCH];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cgroup = blockIdx.x % col_groups, rb = blockIdx.x / col_groups;
  const int c0 = cgroup * WCH + lane * WV;
  if (bt.col_sum != nullptr) {      // training-mode finalisation folded in (see bn_act_apply_kernel)
    const int c = cgroup * WCH + threadIdx.x;
    if (c < C) {
      const float mean_nb = bt.col_sum[c] / bt.count;                       // mean of the stored (bias-free) tensor
      const float var = fmaxf(bt.col_sumsq[c] / bt.count - mean_nb * mean_nb, 0.f);
      const float mean = mean_nb + (bt.bias ? bt.bias[c] : 0.f);            // moving mean only (see bn_finalize_train_kernel)
      const float rstd = rsqrtf(var + bt.eps);
      const float scv = bt.gamma[c] * rstd;
      const float shv = bt.beta[c] - mean_nb * scv;
      s_sc[threadIdx.x] = scv;
      s_sh[threadIdx.x] = shv;
      if (rb == 0) {
        bt.scale_out[c] = scv;
        bt.shift_out[c] = shv;
        bt.save_mean[c] = mean_nb;
        bt.save_rstd[c] = rstd;
        if (bt.moving_mean) {
          const float mv = bt.unbiased ? var * (bt.count / fmaxf(bt.count - 1.f, 1.f)) : var;
          bt.moving_mean[c] = bt.moving_mean[c] * bt.momentum + mean * (1.f - bt.momentum);
          bt.moving_var[c] = bt.moving_var[c] * bt.momentum + mv * (1.f - bt.momentum);
        }
      }
    }
    __syncthreads();
  }
  if (c0 >= C) return;
  float sc[WV], sh[WV], al[WV];
  if (bt.col_sum != nullptr) {
#pragma unroll
    for (int j = 0; j < WV; ++j) { sc[j] = s_sc[lane * WV + j]; sh[j] = s_sh[lane * WV + j]; }
  } else {
    load8f(scale + c0, sc);
    load8f(shift + c0, sh);
  }
#pragma unroll
  for (int j = 0; j < WV; ++j) al[j] = (ACT == ACT_PRELU) ? alpha[c0 + j] : 0.f;
  const int m0 = rb * rows_per_block, m1 = min(m0 + rows_per_block, rows);
  FlatWalk fw[NR];
#pragma unroll
  for (int u = 0; u < NR; ++u) fw[u] = flat_start(min(m0 + w + 8 * u, rows - 1), seg_len, seg_valid, lengths, rows);
  const long long step = 8 * ld;
  const __nv_bfloat16* yp = y + static_cast<long long>(m0 + w) * ld + c0;
  __nv_bfloat16* ap = a + static_cast<long long>(m0 + w) * ld + c0;
  for (int m = m0 + w; m < m1; m += 8 * NR, yp += NR * step, ap += NR * step) {
    float v[NR][WV];
    bool in[NR], valid[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      in[u] = m + 8 * u < m1;
      valid[u] = in[u] && fw[u].t < fw[u].L;
      if (valid[u]) load8(yp + u * step, v[u]);
    }
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      if (in[u]) {
        float o[WV];
#pragma unroll
        for (int j = 0; j < WV; ++j) o[j] = valid[u] ? actf<ACT>(fmaf(v[u][j], sc[j], sh[j]), al[j]) : 0.f;
        store8(ap + u * step, o);
      }
      if (m + 8 * (u + NR) < m1) flat_advance(fw[u], 8 * NR, seg_len, seg_valid, lengths);
    }
  }
}

// ncu (profiles/r01_hbm_kernels_ncu.md): 72 registers = three blocks per SM (34 % achieved occupancy); forcing 64
// registers / four blocks (XV_BWD_FLAT_MINBLOCKS=4) changed nothing, so occupancy is not what bounds it.
template <bool FUSED_POOL, int ACT, int NR>
__global__ void __launch_bounds__(256, (!FUSED_POOL && NR == 1) ? XV_BWD_FLAT_MINBLOCKS : 1) bn_act_bwd_apply_flat_kernel(
    const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ dy,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ save_mean,
    const float* __restrict__ save_rstd, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
    float inv_count, const float* __restrict__ alpha, int rows, int rows_per_block, int col_groups, int C, long long ld,
    int seg_len, int seg_valid, const int* __restrict__ lengths, PoolGradSrc ps) {
  pdl_entry();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cgroup = blockIdx.x % col_groups, rb = blockIdx.x / col_groups;
  const int c0 = cgroup * WCH + lane * WV;
  if (c0 >= C) return;
  float sc[WV], sh[WV], ka[WV], kb[WV], al[WV];
  load8f(scale + c0, sc); load8f(shift + c0, sh);
  {
    float mu[WV], rs[WV], dg[WV], db[WV];
    load8f(save_mean + c0, mu); load8f(save_rstd + c0, rs); load8f(dgamma + c0, dg); load8f(dbeta + c0, db);
#pragma unroll
    for (int j = 0; j < WV; ++j) {
      ka[j] = -sc[j] * rs[j] * dg[j] * inv_count;
      kb[j] = -sc[j] * db[j] * inv_count - ka[j] * mu[j];
      al[j] = (ACT == ACT_PRELU) ? alpha[c0 + j] : 0.f;
    }
  }
  const int m0 = rb * rows_per_block, m1 = min(m0 + rows_per_block, rows);
  FlatWalk fw[NR];
#pragma unroll
  for (int u = 0; u < NR; ++u) fw[u] = flat_start(min(m0 + w + 8 * u, rows - 1), seg_len, seg_valid, lengths, rows);
  PoolCoef8 pc;
  int pc_b = -1;
  const long long step = 8 * ld;
  const long long off = static_cast<long long>(m0 + w) * ld + c0;
  const __nv_bfloat16* yp = y + off;
  const __nv_bfloat16* dp = FUSED_POOL ? nullptr : da + off;
  __nv_bfloat16* op = dy + off;
  for (int m = m0 + w; m < m1; m += 8 * NR, yp += NR * step, op += NR * step) {
    float v[NR][WV], d[NR][WV];
    bool in[NR], valid[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      in[u] = m + 8 * u < m1;
      valid[u] = in[u] && fw[u].t < fw[u].L;
      if (valid[u]) {
        load8(yp + u * step, v[u]);
        if (!FUSED_POOL) load8(dp + u * step, d[u]);
      }
    }
    if (!FUSED_POOL) dp += NR * step;
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      if (in[u]) {
        float o[WV];
        if (valid[u]) {
          if (FUSED_POOL && fw[u].b != pc_b) {       // warp-uniform: a warp's rows cross a segment boundary rarely
            pc_b = fw[u].b;
            pool_coef_load8(pc, ps, pc_b, c0, fw[u].L);
          }
#pragma unroll
          for (int j = 0; j < WV; ++j) {
            const float z = fmaf(v[u][j], sc[j], sh[j]);
            const float dd = FUSED_POOL ? fmaf(pc.cb[j], actf<ACT>(z, al[j]), pc.ca[j]) : d[u][j];
            const float g = dd * actg<ACT>(z, al[j]);
            o[j] = fmaf(sc[j], g, fmaf(ka[j], v[u][j], kb[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < WV; ++j) o[j] = 0.f;
        }
        store8(op + u * step, o);
      }
      if (m + 8 * (u + NR) < m1) flat_advance(fw[u], 8 * NR, seg_len, seg_valid, lengths);
    }
  }
}

// Launch geometry of the flat kernels: grid = resident capacity (a multiple of the channel groups), equal row ranges.
#ifndef XV_FLAT_APPLY_DEFAULT
#define XV_FLAT_APPLY_DEFAULT 1
#endif
#ifndef XV_FLAT_BWD_DEFAULT
#define XV_FLAT_BWD_DEFAULT 1
#endif
#ifndef XV_FLAT_FUSED_DEFAULT
#define XV_FLAT_FUSED_DEFAULT 0       // 59.8 us flat (1 row) vs 49.6 us chunked with 4 rows in flight: coefficient reloads dominate
#endif
struct FlatGrid {
  int grid, rows_per_block, col_groups;
};
template <typename K>
static int flat_grid(K kernel, int* cache, long long rows, int C, int nr, FlatGrid* g) {
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  if (*cache <= 0) {
    int per_sm = 0;
    XV_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0));
    *cache = per_sm > 0 ? per_sm : 1;
  }
  g->col_groups = ceil_div(C, WCH);
  int row_blocks = (*cache * sms) / g->col_groups;
  if (row_blocks < 1) row_blocks = 1;
  const int quantum = 8 * nr;
  long long rpb = (rows + row_blocks - 1) / row_blocks;
  rpb = (rpb + quantum - 1) / quantum * quantum;
  row_blocks = static_cast<int>((rows + rpb - 1) / rpb);
  g->rows_per_block = static_cast<int>(rpb);
  g->grid = row_blocks * g->col_groups;
  return XV_OK;
}
// 0: chunked kernels; 1 / 2: flat kernels with 1 / 2 rows in flight per warp.  Per kernel family: XV_FLAT_APPLY (BN apply),
// XV_FLAT_BWD (BN backward apply), XV_FLAT_FUSED (tdnn5 backward apply with the on-the-fly pooling gradient); XV_FLAT
// sets all three.  Defaults from the on-box sweep (tools/layers_bench.py, profiles/).
enum { FLAT_APPLY = 0, FLAT_BWD = 1, FLAT_FUSED = 2 };
static int flat_mode(int which) {
  static int mode[3] = {-1, -1, -1};
  if (mode[which] < 0) {
    static const char* names[3] = {"XV_FLAT_APPLY", "XV_FLAT_BWD", "XV_FLAT_FUSED"};
    static const int defaults[3] = {XV_FLAT_APPLY_DEFAULT, XV_FLAT_BWD_DEFAULT, XV_FLAT_FUSED_DEFAULT};
    const char* e = getenv(names[which]);
    if (!e) e = getenv("XV_FLAT");
    int m = e ? atoi(e) : defaults[which];
    if (m < 0 || m > 2) m = defaults[which];
    mode[which] = m;
  }
  return mode[which];
}

// XV_POOL_RELU_FAST=0 selects the generic training kernel for the ReLU case too (A/B measurements)
static bool pool_relu_fast() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("XV_POOL_RELU_FAST");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

}  // namespace xv

using namespace xv;

extern "C" int xv_pack_input(const float* x, void* out, int B, int T, int D, int k, int dpad, int64_t ldo, void* stream) {
  if (!x || !out || B <= 0 || T <= 0 || D <= 0 || D > dpad || k * dpad > ldo || dpad % 8 || ldo % 8)
    return set_error(XV_ERR_INVALID, "xv_pack_input: bad arguments (dpad and ldo must be multiples of 8)");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  const long long total = static_cast<long long>(B) * T * (ldo / 8);
  ::xv::launch_pdl((pack_input_kernel), grid_for(total, 256, sms), 256, 0, static_cast<cudaStream_t>(stream), 
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
  ::xv::launch_pdl((bn_finalize_train_kernel), ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream), 
      col_sum, col_sumsq, bias, count, gamma, beta, moving_mean, moving_var, momentum, eps, unbiased, scale, shift,
      save_mean, save_rstd, C);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_finalize_infer(const float* gamma, const float* beta, const float* moving_mean,
                                    const float* moving_var, const float* bias, float eps, float* scale, float* shift,
                                    int C, void* stream) {
  if (!gamma || !beta || !moving_mean || !moving_var || !scale || !shift || C <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_finalize_infer: bad arguments");
  ::xv::launch_pdl((bn_finalize_infer_kernel), ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream), gamma, beta, moving_mean,
                   moving_var, bias, eps, scale, shift, C);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

static int check_act_layout(const char* who, int C, int64_t ld, int act, const float* alpha) {
  if (C <= 0 || C % 8 || ld % 8 || ld < C) return set_error(XV_ERR_INVALID, "%s: C and ld must be multiples of 8, ld >= C", who);
  if (act < ACT_NONE || act > ACT_TANH) return set_error(XV_ERR_INVALID, "%s: unknown activation %d", who, act);
  if (act == ACT_PRELU && !alpha) return set_error(XV_ERR_INVALID, "%s: prelu needs alpha", who);
  return XV_OK;
}

static int launch_bn_act_apply(const void* y, void* a, const float* scale, const float* shift, const float* alpha, int act,
                               int64_t rows, int C, int64_t ld, int seg_len, int seg_valid, const int32_t* lengths,
                               const BnTrainSrc& bt, void* stream) {
  int rc = check_act_layout("xv_bn_act_apply", C, ld, act, alpha); if (rc) return rc;
  if (rows > 0x7fffffffLL) return set_error(XV_ERR_INVALID, "xv_bn_act_apply: rows must fit in int32");
  if (seg_len > 0 && rows % seg_len) return set_error(XV_ERR_INVALID, "rows must be a multiple of seg_len");
  if (flat_mode(FLAT_APPLY) > 0) {
    static int occ[2][8];
    const int nr = flat_mode(FLAT_APPLY);
    FlatGrid fg;
    if (nr == 1) {
      XV_ACT_DISPATCH(act, {
        rc = flat_grid(bn_act_apply_flat_kernel<A_, 1>, &occ[0][A_], rows, C, 1, &fg); if (rc) return rc;
        ::xv::launch_pdl((bn_act_apply_flat_kernel<A_, 1>), fg.grid, 256, 0, static_cast<cudaStream_t>(stream),
            static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(a), scale, shift, alpha,
            static_cast<int>(rows), fg.rows_per_block, fg.col_groups, C, ld, seg_len, seg_valid, lengths, bt); });
    } else {
      XV_ACT_DISPATCH(act, {
        rc = flat_grid(bn_act_apply_flat_kernel<A_, 2>, &occ[1][A_], rows, C, 2, &fg); if (rc) return rc;
        ::xv::launch_pdl((bn_act_apply_flat_kernel<A_, 2>), fg.grid, 256, 0, static_cast<cudaStream_t>(stream),
            static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(a), scale, shift, alpha,
            static_cast<int>(rows), fg.rows_per_block, fg.col_groups, C, ld, seg_len, seg_valid, lengths, bt); });
    }
    XV_CUDA_CHECK(cudaGetLastError());
    return XV_OK;
  }
  const RowGrid rg = make_row_grid(rows, seg_len, C, WCH);
  const unsigned grid = static_cast<unsigned>(ceil_div(C, WCH) * (rows / rg.seg_len) * rg.chunks);   // row chunk major, channel group minor
  XV_ACT_DISPATCH(act, (::xv::launch_pdl((bn_act_apply_kernel<A_>), grid, 256, 0, static_cast<cudaStream_t>(stream),
      static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(a), scale, shift, alpha,
      rg, C, ld, seg_len, seg_valid, lengths, bt)));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_act_apply(const void* y, void* a, const float* scale, const float* shift, const float* alpha,
                               int act, int64_t rows, int C, int64_t ld, int seg_len, int seg_valid,
                               const int32_t* lengths, void* stream) {
  if (!y || !a || !scale || !shift || rows <= 0) return set_error(XV_ERR_INVALID, "xv_bn_act_apply: bad arguments");
  BnTrainSrc bt;
  memset(&bt, 0, sizeof(bt));
  return launch_bn_act_apply(y, a, scale, shift, alpha, act, rows, C, ld, seg_len, seg_valid, lengths, bt, stream);
}

extern "C" int xv_bn_train_apply(const void* y, void* a, const float* col_sum, const float* col_sumsq, const float* bias,
                                 float count, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                                 float momentum, float eps, int unbiased_moving_var, float* scale, float* shift,
                                 float* save_mean, float* save_rstd, const float* alpha, int act, int64_t rows, int C,
                                 int64_t ld, int seg_len, int seg_valid, const int32_t* lengths, void* stream) {
  if (!y || !a || !col_sum || !col_sumsq || !gamma || !beta || !scale || !shift || !save_mean || !save_rstd || rows <= 0 ||
      count <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_train_apply: bad arguments");
  BnTrainSrc bt{col_sum, col_sumsq, bias, gamma, beta, moving_mean, moving_var, scale, shift, save_mean, save_rstd,
                count, momentum, eps, unbiased_moving_var};
  return launch_bn_act_apply(y, a, scale, shift, alpha, act, rows, C, ld, seg_len, seg_valid, lengths, bt, stream);
}

extern "C" int xv_col_stats(const void* y, const float* bias, int64_t rows, int C, int64_t ld, int seg_len, int seg_valid,
                            const int32_t* lengths, float* col_sum, float* col_sumsq, void* stream) {
  if (!y || !col_sum || !col_sumsq || rows <= 0 || rows > 0x7fffffffLL || C <= 0 || C % 8 || ld % 8 || ld < C)
    return set_error(XV_ERR_INVALID, "xv_col_stats: bad arguments");
  if (seg_len > 0 && rows % seg_len) return set_error(XV_ERR_INVALID, "rows must be a multiple of seg_len");
  const RowGrid rg = make_row_grid(rows, seg_len, C);
  const unsigned grid = static_cast<unsigned>(ceil_div(C, SCH) * (rows / rg.seg_len) * rg.chunks);   // row chunk major, channel group minor
  ::xv::launch_pdl((col_stats_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(y), bias, rg, C, ld, seg_len, seg_valid, lengths, col_sum,
      col_sumsq);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

static int check_pool_src(const char* who, const void* da, const float* pooled, const float* dpooled, int pool_cpad,
                          int C, int seg_len) {
  if (!da && !pooled) return set_error(XV_ERR_INVALID, "%s: need either the upstream gradient or the pooled statistics", who);
  if (pooled && (!dpooled || pool_cpad != C || seg_len <= 0))
    return set_error(XV_ERR_INVALID, "%s: fused pooling gradient needs dpooled, pool_cpad == C and seg_len > 0", who);
  return XV_OK;
}

extern "C" int xv_bn_act_bwd_reduce(const void* y, const void* da, const float* scale, const float* shift,
                                    const float* save_mean, const float* save_rstd, const float* alpha, int act,
                                    int64_t rows, int C, int64_t ld, int seg_len, int seg_valid, const int32_t* lengths,
                                    float* dgamma, float* dbeta, float* dalpha, const float* pooled,
                                    const float* dpooled, int pool_cpad, int pool_c_real, void* stream) {
  { int rcp = check_pool_src("xv_bn_act_bwd_reduce", da, pooled, dpooled, pool_cpad, C, seg_len); if (rcp) return rcp; }
  if (!y || !scale || !shift || !save_mean || !save_rstd || !dgamma || !dbeta || rows <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_act_bwd_reduce: bad arguments");
  int rc = check_act_layout("xv_bn_act_bwd_reduce", C, ld, act, alpha); if (rc) return rc;
  if (rows > 0x7fffffffLL) return set_error(XV_ERR_INVALID, "xv_bn_act_bwd_reduce: rows must fit in int32");
  if (seg_len > 0 && rows % seg_len) return set_error(XV_ERR_INVALID, "rows must be a multiple of seg_len");
  const RowGrid rg = make_row_grid(rows, seg_len, C, WCH, pooled ? FUSED_STREAM_ROWS : STREAM_ROWS);
  const unsigned grid = static_cast<unsigned>(ceil_div(C, WCH) * (rows / rg.seg_len) * rg.chunks);   // row chunk major, channel group minor
  PoolGradSrc ps{pooled, dpooled, pool_cpad, pool_c_real};
  if (pooled) {
    XV_ACT_DISPATCH(act, (::xv::launch_pdl((bn_act_bwd_reduce_kernel<true, A_>), grid, 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(y), nullptr, scale, shift, save_mean, save_rstd, alpha,
        rg, C, ld, seg_len, seg_valid, lengths, dgamma, dbeta, dalpha, ps)));
  } else {
    XV_ACT_DISPATCH(act, (::xv::launch_pdl((bn_act_bwd_reduce_kernel<false, A_>), grid, 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(da), scale, shift, save_mean, save_rstd,
        alpha, rg, C, ld, seg_len, seg_valid, lengths, dgamma, dbeta, dalpha, ps)));
  }
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_bn_act_bwd_apply(const void* y, const void* da, void* dy, const float* scale, const float* shift,
                                   const float* save_mean, const float* save_rstd, const float* dgamma,
                                   const float* dbeta, float count, const float* alpha, int act, int64_t rows, int C,
                                   int64_t ld, int seg_len, int seg_valid, const int32_t* lengths, const float* pooled,
                                   const float* dpooled, int pool_cpad, int pool_c_real, void* stream) {
  { int rcp = check_pool_src("xv_bn_act_bwd_apply", da, pooled, dpooled, pool_cpad, C, seg_len); if (rcp) return rcp; }
  if (!y || !dy || !scale || !shift || !save_mean || !save_rstd || !dgamma || !dbeta || rows <= 0 || count <= 0)
    return set_error(XV_ERR_INVALID, "xv_bn_act_bwd_apply: bad arguments");
  int rc = check_act_layout("xv_bn_act_bwd_apply", C, ld, act, alpha); if (rc) return rc;
  if (rows > 0x7fffffffLL) return set_error(XV_ERR_INVALID, "xv_bn_act_bwd_apply: rows must fit in int32");
  PoolGradSrc ps{pooled, dpooled, pool_cpad, pool_c_real};
  if (seg_len > 0 && rows % seg_len) return set_error(XV_ERR_INVALID, "rows must be a multiple of seg_len");
  if (flat_mode(pooled ? FLAT_FUSED : FLAT_BWD) > 0) {
    static int occ[2][2][8];
    const int nr = flat_mode(pooled ? FLAT_FUSED : FLAT_BWD);
    const cudaStream_t s_ = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(y);
    const __nv_bfloat16* dab = static_cast<const __nv_bfloat16*>(da);
    __nv_bfloat16* dyb = static_cast<__nv_bfloat16*>(dy);
    const int rows_i = static_cast<int>(rows);
    const float invc = 1.0f / count;
    FlatGrid fg;
#define XV_FLAT_BWD(FUSED, NRV)                                                                                          \
    XV_ACT_DISPATCH(act, {                                                                                               \
      rc = flat_grid(bn_act_bwd_apply_flat_kernel<FUSED, A_, NRV>, &occ[FUSED ? 1 : 0][NRV - 1][A_], rows, C, NRV, &fg); \
      if (rc) return rc;                                                                                                 \
      ::xv::launch_pdl((bn_act_bwd_apply_flat_kernel<FUSED, A_, NRV>), fg.grid, 256, 0, s_, yb, FUSED ? nullptr : dab, dyb,  \
                       scale, shift, save_mean, save_rstd, dgamma, dbeta, invc, alpha, rows_i, fg.rows_per_block,        \
                       fg.col_groups, C, ld, seg_len, seg_valid, lengths, ps); })
    if (pooled) { if (nr == 1) { XV_FLAT_BWD(true, 1); } else { XV_FLAT_BWD(true, 2); } }
    else { if (nr == 1) { XV_FLAT_BWD(false, 1); } else { XV_FLAT_BWD(false, 2); } }
#undef XV_FLAT_BWD
    XV_CUDA_CHECK(cudaGetLastError());
    return XV_OK;
  }
  const RowGrid rg = make_row_grid(rows, seg_len, C, WCH, pooled ? FUSED_STREAM_ROWS : STREAM_ROWS);
  const unsigned grid = static_cast<unsigned>(ceil_div(C, WCH) * (rows / rg.seg_len) * rg.chunks);   // row chunk major, channel group minor
  if (pooled && act == ACT_RELU && pool_relu_fast()) {
    ::xv::launch_pdl((bn_act_bwd_apply_pool_relu_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream),
                     static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(dy), scale, shift, save_mean, save_rstd,
                     dgamma, dbeta, 1.0f / count, rg, C, static_cast<long long>(ld), seg_len, seg_valid, lengths, ps);
  } else if (pooled) {
    XV_ACT_DISPATCH(act, (::xv::launch_pdl((bn_act_bwd_apply_kernel<true, A_>), grid, 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(y), nullptr, static_cast<__nv_bfloat16*>(dy), scale, shift, save_mean,
        save_rstd, dgamma, dbeta, 1.0f / count, alpha, rg, C, ld, seg_len, seg_valid, lengths, ps)));
  } else {
    XV_ACT_DISPATCH(act, (::xv::launch_pdl((bn_act_bwd_apply_kernel<false, A_>), grid, 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(da), static_cast<__nv_bfloat16*>(dy),
        scale, shift, save_mean, save_rstd, dgamma, dbeta, 1.0f / count, alpha, rg, C, ld,
        seg_len, seg_valid, lengths, ps)));
  }
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_stats_pool_fwd(const void* x, float* out, void* out_split, int B, int seg_len, int seg_valid,
                                 const int32_t* lengths, int c_real, int cpad, int64_t ld, const float* scale,
                                 const float* shift, const float* alpha, int act, const float* save_mean,
                                 const float* save_rstd, float* bwd_sums, void* stream) {
  if (scale && (!shift || (act == ACT_PRELU && !alpha))) return set_error(XV_ERR_INVALID, "xv_stats_pool_fwd: fused BN needs shift (and alpha for prelu)");
  if (!x || !out || B <= 0 || seg_len <= 0 || cpad % 8 || c_real > cpad || ld % 8 || ld < cpad)
    return set_error(XV_ERR_INVALID, "xv_stats_pool_fwd: bad arguments");
  if (bwd_sums && (!scale || !save_mean || !save_rstd || act == ACT_PRELU))
    return set_error(XV_ERR_INVALID, "xv_stats_pool_fwd: backward sums need the fused BN (scale, shift, saved mean / rstd) and a non-prelu activation");
  if (act < ACT_NONE || act > ACT_TANH) return set_error(XV_ERR_INVALID, "xv_stats_pool_fwd: unknown activation %d", act);
  const cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* o3 = static_cast<__nv_bfloat16*>(out_split);
  if (bwd_sums && act == ACT_RELU && pool_relu_fast()) {
    dim3 grid(ceil_div(cpad, 256), B);
    ::xv::launch_pdl((stats_pool_fwd_relu_sums_kernel), grid, 256, 0, s_, xb, out, o3, seg_len, seg_valid, lengths, c_real, cpad,
                     static_cast<long long>(ld), scale, shift, save_mean, save_rstd, bwd_sums);
  } else if (bwd_sums) {
    dim3 grid(ceil_div(cpad, 32 * XV_POOL_CPT), B);
    XV_ACT_DISPATCH(act, (::xv::launch_pdl((stats_pool_fwd_kernel<true, XV_POOL_CPT, A_>), grid, 256, 0, s_, xb, out, o3, seg_len,
                                           seg_valid, lengths, c_real, cpad, ld, scale, shift, alpha, save_mean, save_rstd,
                                           bwd_sums, static_cast<const int*>(nullptr))));
  } else {
    dim3 grid(ceil_div(cpad, 256), B);
    const float* nf = nullptr;
    float* nfm = nullptr;
    XV_ACT_DISPATCH(act, (::xv::launch_pdl((stats_pool_fwd_kernel<false, 8, A_>), grid, 256, 0, s_, xb, out, o3, seg_len,
                                           seg_valid, lengths, c_real, cpad, ld, scale, shift, alpha, nf, nf, nfm,
                                           static_cast<const int*>(nullptr))));
  }
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_stats_pool_ragged(const void* x, float* out, void* out_split, int num_segments, const int32_t* starts,
                                    const int32_t* lengths, int c_real, int cpad, int64_t ld, const float* scale,
                                    const float* shift, const float* alpha, int act, void* stream) {
  if (scale && (!shift || (act == ACT_PRELU && !alpha))) return set_error(XV_ERR_INVALID, "xv_stats_pool_ragged: fused BN needs shift (and alpha for prelu)");
  if (!x || !out || !starts || !lengths || num_segments <= 0 || cpad % 8 || c_real > cpad || ld % 8 || ld < cpad)
    return set_error(XV_ERR_INVALID, "xv_stats_pool_ragged: bad arguments");
  if (act < ACT_NONE || act > ACT_TANH) return set_error(XV_ERR_INVALID, "xv_stats_pool_ragged: unknown activation %d", act);
  dim3 grid(ceil_div(cpad, 256), num_segments);
  const float* nf = nullptr;
  float* nfm = nullptr;
  XV_ACT_DISPATCH(act, (::xv::launch_pdl((stats_pool_fwd_kernel<false, 8, A_>), grid, 256, 0, static_cast<cudaStream_t>(stream),
                                         static_cast<const __nv_bfloat16*>(x), out, static_cast<__nv_bfloat16*>(out_split), 1, 0,
                                         lengths, c_real, cpad, ld, scale, shift, alpha, nf, nf, nfm, starts)));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_pool_bn_bwd_reduce(const float* pooled, const float* dpooled, const float* bwd_sums, int B, int seg_valid,
                                     const int32_t* lengths, int c_real, int cpad, float* dgamma, float* dbeta, void* stream) {
  if (!pooled || !dpooled || !bwd_sums || !dgamma || !dbeta || B <= 0 || c_real <= 0 || c_real > cpad)
    return set_error(XV_ERR_INVALID, "xv_pool_bn_bwd_reduce: bad arguments");
  dim3 grid(ceil_div(c_real, 256), B < 64 ? B : 64);
  ::xv::launch_pdl((pool_bn_bwd_reduce_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream), pooled, dpooled, bwd_sums, B, seg_valid,
                                                                                 lengths, c_real, cpad, dgamma, dbeta);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_stats_pool_bwd(const void* x, const float* pooled, const float* dpooled, void* dx, int B, int seg_len,
                                 int seg_valid, const int32_t* lengths, int c_real, int cpad, int64_t ld, void* stream) {
  if (!x || !pooled || !dpooled || !dx || B <= 0 || seg_len <= 0 || cpad % 8 || c_real > cpad || ld % 8 || ld < cpad)
    return set_error(XV_ERR_INVALID, "xv_stats_pool_bwd: bad arguments");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  const long long total = static_cast<long long>(B) * seg_len * (cpad / 8);
  ::xv::launch_pdl((stats_pool_bwd_kernel), grid_for(total, 256, sms), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), pooled, dpooled, static_cast<__nv_bfloat16*>(dx), B, seg_len, seg_valid,
      lengths, c_real, cpad, ld);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
