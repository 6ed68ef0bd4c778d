// Metric-learning losses on the embeddings: semi-hard triplet (model/loss.py:358-498), angular triplet (loss.py:501-634),
// the softmax GE2E validation loss (loss.py:637-705) and the pairwise matrices they are built on (model/common.py:61-110).
//
// Everything is a function of the Gram matrix G = X X^T of the [B, E] embeddings (B = speakers x segments per batch, a few
// hundred rows): one fp32 CUDA-core GEMM produces G, the mining kernels run one block per anchor on a row of the pairwise
// matrix held in shared memory and emit dLoss/d(pairwise entry) for that row (rows are block-exclusive: no global atomics),
// a "finish" kernel turns that into a symmetric coefficient matrix M and a row scale d with
//        dLoss/dX = M X + d o X,
// and a second fp32 GEMM evaluates it.  fp32 throughout: the mining decisions (hardest / semi-hard negative, violated
// triplets) are comparisons between pairwise entries, so the products are not rounded to bf16.
#include <cuda_bf16.h>
#include <math.h>

#include "xv_internal.h"

namespace xv {

constexpr int MET_MAX_B = 2048;          // rows of the pairwise matrix (shared-memory rows, [B, B] work matrices)
constexpr float MET_EPS = 1e-12f;

__device__ __forceinline__ float m_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float m_warp_min(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float m_warp_max(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reductions for 256-thread blocks; `red` is 8 floats of shared memory; every thread gets the result
__device__ __forceinline__ float m_block_sum(float v, float* red) {
  v = m_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ float m_block_max(float v, float* red) {
  v = m_warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) s = fmaxf(s, red[w]);
  return s;
}
__device__ __forceinline__ float m_block_min(float v, float* red) {
  v = m_warp_min(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) s = fminf(s, red[w]);
  return s;
}

// ------------------------------------------------------------------------------------------------
// fp32 GEMM, 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread.
//   NT: C[M,N] = A[M,K] B[N,K]^T            (the Gram matrix: A = B = X)
//   NN: C[M,N] = A[M,K] B[K,N] + rs[m] B2[m,n]   (dX = M X + d o X)
template <bool NT>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Cm,
                                                    int M, int N, int K, long long lda, long long ldb, long long ldc,
                                                    const float* __restrict__ rs, const float* __restrict__ B2, long long ldb2) {
  pdl_entry();
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = threadIdx.x + r * 256;         // 1024 elements of each tile
      {
        const int mm = idx >> 4, kk = idx & 15;      // A: rows m, K contiguous
        const int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < M && gk < K) ? A[static_cast<long long>(gm) * lda + gk] : 0.f;
      }
      if (NT) {
        const int nn = idx >> 4, kk = idx & 15;
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < N && gk < K) ? Bm[static_cast<long long>(gn) * ldb + gk] : 0.f;
      } else {
        const int kk = idx >> 6, nn = idx & 63;      // B: rows k, N contiguous
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < N && gk < K) ? Bm[static_cast<long long>(gk) * ldb + gn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
    const float r = rs ? rs[gm] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (rs) v = fmaf(r, B2[static_cast<long long>(gm) * ldb2 + gn], v);
      Cm[static_cast<long long>(gm) * ldc + gn] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pairwise Euclidean distances from the Gram matrix (model/common.py:61-93).
__global__ void __launch_bounds__(256) euclid_from_gram_kernel(const float* __restrict__ G, float* __restrict__ D, int B, int squared) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= static_cast<long long>(B) * B) return;
  const int r = static_cast<int>(i / B), c = static_cast<int>(i % B);
  float d = G[static_cast<long long>(r) * B + r] - 2.0f * G[i] + G[static_cast<long long>(c) * B + c];
  d = fmaxf(d, 0.f);
  if (!squared) d = (d == 0.f) ? 0.f : sqrtf(d);
  D[i] = d;
}

// Positive pairs of the batch (label equal, i != j): depends on the labels only.
__global__ void __launch_bounds__(256) count_positive_pairs_kernel(const int* __restrict__ labels, int B, float* __restrict__ counter) {
  pdl_entry();
  __shared__ float red[8];
  const int x = blockIdx.x;
  const int lx = labels[x];
  float c = 0.f;
  for (int y = threadIdx.x; y < B; y += 256) c += (y != x && labels[y] == lx) ? 1.f : 0.f;
  c = m_block_sum(c, red);
  if (threadIdx.x == 0 && c > 0.f) atomicAdd(counter, c);
}

// Semi-hard mining, one block per anchor x; warps take the positives i of the anchor in turn.
//   semi(x,i) = min{D_xy : y negative, D_xy > D_xi}  if that set is non-empty, else max{D_xy : y negative}
//   loss     += scale / num_pos * max(margin + D_xi - semi, 0)
//   dD[x,i]  += w,  dD[x,y*] -= w / ties   (w = scale / num_pos; the argmin / argmax shares the gradient among ties like
//   tf.reduce_min / tf.reduce_max)
__global__ void __launch_bounds__(256) semihard_kernel(const float* __restrict__ D, const int* __restrict__ labels, int B,
                                                       float margin, float scale, const float* __restrict__ num_pos,
                                                       float* __restrict__ loss, float* __restrict__ dD) {
  pdl_entry();
  extern __shared__ float sm[];             // [3][B]: D row | dD row | labels (as int)
  __shared__ float red[8];
  float* drow = sm;
  float* grow = sm + B;
  int* lab = reinterpret_cast<int*>(sm + 2 * B);
  const int x = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int y = threadIdx.x; y < B; y += 256) {
    drow[y] = D[static_cast<long long>(x) * B + y];
    grow[y] = 0.f;
    lab[y] = labels[y];
  }
  __syncthreads();
  const int lx = lab[x];
  float mx = -INFINITY, mn = INFINITY;
  for (int y = threadIdx.x; y < B; y += 256) {
    mn = fminf(mn, drow[y]);
    if (lab[y] != lx) mx = fmaxf(mx, drow[y]);
  }
  const float inside_raw = m_block_max(mx, red);
  const float rowmin = m_block_min(mn, red);
  const bool any_neg = inside_raw > -INFINITY;
  const float inside = any_neg ? inside_raw : rowmin;         // _masked_maximum over an empty mask = the row minimum
  const float np = fmaxf(num_pos[0], 1e-16f);
  const float w = scale / np;
  float lsum = 0.f;
  for (int i = wp; i < B; i += 8) {
    if (i == x || lab[i] != lx) continue;                     // uniform over the warp
    const float dxi = drow[i];
    float m = INFINITY;
    for (int y = lane; y < B; y += 32)
      if (lab[y] != lx && drow[y] > dxi) m = fminf(m, drow[y]);
    m = m_warp_min(m);
    const bool has = m < INFINITY;
    const float semi = has ? m : inside;
    const float lm = margin + dxi - semi;
    if (lm >= 0.f) {
      if (lane == 0) {
        lsum += lm;
        atomicAdd(&grow[i], w);
      }
      if (any_neg) {
        float nt = 0.f;
        for (int y = lane; y < B; y += 32)
          if (lab[y] != lx && drow[y] == semi && (!has || drow[y] > dxi)) nt += 1.f;
        nt = m_warp_sum(nt);
        const float share = -w / nt;
        for (int y = lane; y < B; y += 32)
          if (lab[y] != lx && drow[y] == semi && (!has || drow[y] > dxi)) atomicAdd(&grow[y], share);
      }
    }
  }
  lsum = m_block_sum(lsum, red);
  if (threadIdx.x == 0 && lsum != 0.f) atomicAdd(loss, w * lsum);
  __syncthreads();
  for (int y = threadIdx.x; y < B; y += 256) dD[static_cast<long long>(x) * B + y] = grow[y];
}

// dLoss/dD -> (M, d):  dd2_ij = dD_ij * dD/d(d^2);  E = dd2 + dd2^T;  M = -2 E;  d_i = 2 sum_j E_ij.
__device__ __forceinline__ float euclid_chain(float g, float Gii, float Gij, float Gjj, int squared) {
  const float raw = Gii - 2.0f * Gij + Gjj;
  if (squared) return (raw >= 0.f) ? g : 0.f;
  return (raw > 0.f) ? g * 0.5f * rsqrtf(raw) : 0.f;
}
__global__ void __launch_bounds__(256) euclid_finish_kernel(const float* __restrict__ G, const float* __restrict__ dD,
                                                            float* __restrict__ Mc, float* __restrict__ diag, int B, int squared) {
  pdl_entry();
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float Gii = G[static_cast<long long>(i) * B + i];
  float s = 0.f;
  for (int j = threadIdx.x; j < B; j += 256) {
    const float Gij = G[static_cast<long long>(i) * B + j], Gjj = G[static_cast<long long>(j) * B + j];
    const float e = euclid_chain(dD[static_cast<long long>(i) * B + j], Gii, Gij, Gjj, squared) +
                    euclid_chain(dD[static_cast<long long>(j) * B + i], Gjj, G[static_cast<long long>(j) * B + i], Gii, squared);
    Mc[static_cast<long long>(i) * B + j] = -2.0f * e;
    s += e;
  }
  s = m_block_sum(s, red);
  if (threadIdx.x == 0) diag[i] = 2.0f * s;
}

// ------------------------------------------------------------------------------------------------
// Angular triplet loss on the pairwise cosines (model/loss.py:501-634).
enum { ANG_ASOFTMAX = 0, ANG_AM = 1, ANG_ARC = 2 };

struct AngCfg {
  int kind, m_int;
  float margin, cos_m, sin_m, threshold;
};

__device__ __forceinline__ float ang_cos(float Gij, float inv_i, float inv_j) { return fminf(fmaxf(Gij * inv_i * inv_j, -1.f), 1.f); }

// d_p(c) and its derivative (loss.py:535-560); the sqrt of the arc form is floored at 1e-12 (finite slope at |c| = 1)
__device__ __forceinline__ float ang_positive(float c, const AngCfg& a, float* dpdc) {
  if (a.kind == ANG_AM) { *dpdc = 1.f; return c - a.margin; }
  if (a.kind == ANG_ASOFTMAX) {
    if (a.m_int == 1) { *dpdc = 1.f; return c; }
    const float s0 = (c > 0.f) ? 1.f : ((c < 0.f) ? -1.f : 0.f);
    if (a.m_int == 2) { *dpdc = 4.f * s0 * c; return 2.f * s0 * c * c - 1.f; }
    const float c2 = c * c;
    const float t = 2.f * c2 - 1.f;
    const float s3 = ((t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f)) * s0;
    const float s4 = 2.f * s0 + s3 - 3.f;
    *dpdc = s3 * (32.f * c2 * c - 16.f * c);
    return s3 * (8.f * c2 * c2 - 8.f * c2 + 1.f) + s4;
  }
  const float sq = sqrtf(fmaxf(1.f - c * c, 1e-12f));
  const float nw = c * a.cos_m - sq * a.sin_m;
  const float dn = a.cos_m + ((1.f - c * c > 1e-12f) ? (c / sq) * a.sin_m : 0.f);
  if (c <= a.threshold) { *dpdc = -dn; return -nw - 2.f; }
  *dpdc = dn;
  return nw;
}

// shared-memory rows of one anchor: cosine | d_p | d d_p / dc | gradient accumulator | labels
struct AngRows {
  float *c, *p, *dp, *g;
  int* lab;
};
__device__ __forceinline__ AngRows ang_load_rows(float* sm, const float* __restrict__ G, const int* __restrict__ labels, int B, int i,
                                                 const AngCfg& a) {
  AngRows r;
  r.c = sm; r.p = sm + B; r.dp = sm + 2 * B; r.g = sm + 3 * B;
  r.lab = reinterpret_cast<int*>(sm + 4 * B);
  const float inv_i = rsqrtf(fmaxf(G[static_cast<long long>(i) * B + i], MET_EPS));
  for (int j = threadIdx.x; j < B; j += 256) {
    const float inv_j = rsqrtf(fmaxf(G[static_cast<long long>(j) * B + j], MET_EPS));
    const float c = ang_cos(G[static_cast<long long>(i) * B + j], inv_i, inv_j);
    float d;
    r.c[j] = c;
    r.p[j] = ang_positive(c, a, &d);
    r.dp[j] = d;
    r.g[j] = 0.f;
    r.lab[j] = labels[j];
  }
  __syncthreads();
  return r;
}

// "all": counters[0] += sum over the valid triplets of max(c_ik - p_ij, 0); counters[1] += #(that > 1e-12)
__global__ void __launch_bounds__(256) angular_all_count_kernel(const float* __restrict__ G, const int* __restrict__ labels, int B,
                                                                AngCfg a, float* __restrict__ counters) {
  pdl_entry();
  extern __shared__ float sm[];
  __shared__ float red[8];
  const int i = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const AngRows r = ang_load_rows(sm, G, labels, B, i, a);
  const int li = r.lab[i];
  float sum = 0.f, cnt = 0.f;
  for (int j = wp; j < B; j += 8) {
    if (j == i || r.lab[j] != li) continue;
    const float pj = r.p[j];
    for (int k = lane; k < B; k += 32) {
      if (r.lab[k] == li) continue;
      const float t = fmaxf(r.c[k] - pj, 0.f);
      sum += t;
      cnt += (t > 1e-12f) ? 1.f : 0.f;
    }
  }
  sum = m_block_sum(sum, red);
  cnt = m_block_sum(cnt, red);
  if (threadIdx.x == 0 && (sum != 0.f || cnt != 0.f)) {
    atomicAdd(counters, sum);
    atomicAdd(counters + 1, cnt);
  }
}

// "all": loss += scale * sum / (count + 1e-16); dS[i, :] = dLoss/dc[i, :] with w = scale / (count + 1e-16)
__global__ void __launch_bounds__(256) angular_all_grad_kernel(const float* __restrict__ G, const int* __restrict__ labels, int B,
                                                               AngCfg a, float scale, const float* __restrict__ counters,
                                                               float* __restrict__ loss, float* __restrict__ dS) {
  pdl_entry();
  extern __shared__ float sm[];
  const int i = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const AngRows r = ang_load_rows(sm, G, labels, B, i, a);
  const int li = r.lab[i];
  const float w = scale / (counters[1] + 1e-16f);
  if (i == 0 && threadIdx.x == 0) atomicAdd(loss, w * counters[0]);
  for (int j = wp; j < B; j += 8) {
    if (j == i || r.lab[j] != li) continue;
    const float pj = r.p[j];
    float nj = 0.f;
    for (int k = lane; k < B; k += 32) {
      if (r.lab[k] == li) continue;
      if (r.c[k] - pj >= 0.f) {           // tf.maximum passes the gradient to its first argument on ties
        nj += 1.f;
        atomicAdd(&r.g[k], w);
      }
    }
    nj = m_warp_sum(nj);
    if (lane == 0 && nj > 0.f) atomicAdd(&r.g[j], -w * nj * r.dp[j]);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < B; j += 256) dS[static_cast<long long>(i) * B + j] = r.g[j];
}

// "hard": hardest positive (min d_p over the positives) against the hardest negative (max cos over the negatives)
__global__ void __launch_bounds__(256) angular_hard_kernel(const float* __restrict__ G, const int* __restrict__ labels, int B, AngCfg a,
                                                           float scale, float* __restrict__ loss, float* __restrict__ dS) {
  pdl_entry();
  extern __shared__ float sm[];
  __shared__ float red[8];
  const int i = blockIdx.x;
  const AngRows r = ang_load_rows(sm, G, labels, B, i, a);
  const int li = r.lab[i];
  float pmin = INFINITY, nmax = -INFINITY, rowmax = -INFINITY, rowmin = INFINITY;
  for (int j = threadIdx.x; j < B; j += 256) {
    rowmax = fmaxf(rowmax, r.p[j]);
    rowmin = fminf(rowmin, r.p[j]);
    if (r.lab[j] != li) nmax = fmaxf(nmax, r.c[j]);
    else if (j != i) pmin = fminf(pmin, r.p[j]);
  }
  pmin = m_block_min(pmin, red);
  nmax = m_block_max(nmax, red);
  rowmax = m_block_max(rowmax, red);
  rowmin = m_block_min(rowmin, red);
  const bool has_p = pmin < INFINITY, has_n = nmax > -INFINITY;
  // loss.py:615-623: entries outside the mask are filled with the row maximum of d_p (positives) / the row MINIMUM of d_p
  // (negatives), so an anchor with positives takes min(pmin, rowmax) = pmin, one without takes rowmax, etc.
  // (the diagonal is never in either mask, so the fill value always takes part).  An anchor without positives / negatives
  // contributes its loss value, but no gradient flows through the fill.
  const float hp = has_p ? fminf(pmin, rowmax) : rowmax;
  const float hn = has_n ? fmaxf(nmax, rowmin) : rowmin;
  const float t = hn - hp;
  const float w = scale / static_cast<float>(B);
  if (threadIdx.x == 0 && t > 0.f) atomicAdd(loss, w * t);
  float ntp = 0.f, ntn = 0.f;
  if (t >= 0.f) {
    for (int j = threadIdx.x; j < B; j += 256) {
      if (r.lab[j] != li) ntn += (has_n && r.c[j] == hn) ? 1.f : 0.f;
      else if (j != i) ntp += (has_p && r.p[j] == hp) ? 1.f : 0.f;
    }
  }
  ntp = m_block_sum(ntp, red);
  ntn = m_block_sum(ntn, red);
  for (int j = threadIdx.x; j < B; j += 256) {
    float g = 0.f;
    if (t >= 0.f) {
      if (r.lab[j] != li) { if (ntn > 0.f && r.c[j] == hn) g = w / ntn; }
      else if (j != i) { if (ntp > 0.f && r.p[j] == hp) g = -w / ntp * r.dp[j]; }
    }
    dS[static_cast<long long>(i) * B + j] = g;
  }
}

// dLoss/dcos -> (M, d):  H = dS where the clip passes;  Hs = H + H^T;  M_ij = Hs_ij inv_i inv_j;
// d_i = -inv_i^3 sum_j Hs_ij G_ij inv_j  (0 where |x_i|^2 sits on the 1e-12 floor)
__global__ void __launch_bounds__(256) cos_finish_kernel(const float* __restrict__ G, const float* __restrict__ dS, float* __restrict__ Mc,
                                                         float* __restrict__ diag, int B) {
  pdl_entry();
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float Gii = G[static_cast<long long>(i) * B + i];
  const float inv_i = rsqrtf(fmaxf(Gii, MET_EPS));
  float s = 0.f;
  for (int j = threadIdx.x; j < B; j += 256) {
    const float Gij = G[static_cast<long long>(i) * B + j];
    const float inv_j = rsqrtf(fmaxf(G[static_cast<long long>(j) * B + j], MET_EPS));
    const float raw = Gij * inv_i * inv_j;
    const float pass = (raw >= -1.f && raw <= 1.f) ? 1.f : 0.f;
    // G and the clip are symmetric: raw_ji == raw_ij (the GEMM sums both entries in the same order)
    const float hs = pass * (dS[static_cast<long long>(i) * B + j] + dS[static_cast<long long>(j) * B + i]);
    Mc[static_cast<long long>(i) * B + j] = hs * inv_i * inv_j;
    s = fmaf(hs * Gij, inv_j, s);
  }
  s = m_block_sum(s, red);
  if (threadIdx.x == 0) diag[i] = (Gii > MET_EPS) ? -inv_i * inv_i * inv_i * s : 0.f;
}

// ------------------------------------------------------------------------------------------------
// Softmax GE2E validation loss (loss.py:637-705), forward only.  Batches are speaker-ordered: rows [s*m, (s+1)*m) belong to
// speaker s.  f = l2-normalised rows; S_s = sum of the speaker's rows.
__global__ void __launch_bounds__(256) e2e_prepare_kernel(const float* __restrict__ x, float* __restrict__ f, float* __restrict__ S,
                                                          float* __restrict__ Snorm2, int m, int E, long long ldx) {
  pdl_entry();
  __shared__ float red[8];
  __shared__ float inv[64];
  const int s = blockIdx.x;
  for (int r0 = 0; r0 < m; r0 += 64) {         // row norms, 64 rows at a time
    for (int r = 0; r < 64 && r0 + r < m; ++r) {
      const float* xr = x + static_cast<long long>(s * m + r0 + r) * ldx;
      float q = 0.f;
      for (int e = threadIdx.x; e < E; e += 256) q = fmaf(xr[e], xr[e], q);
      q = m_block_sum(q, red);
      if (threadIdx.x == 0) inv[r] = rsqrtf(fmaxf(q, MET_EPS));
    }
    __syncthreads();
    for (int r = 0; r < 64 && r0 + r < m; ++r) {
      const long long row = s * m + r0 + r;
      for (int e = threadIdx.x; e < E; e += 256) f[row * E + e] = x[row * ldx + e] * inv[r];
    }
    __syncthreads();
  }
  float q = 0.f;
  for (int e = threadIdx.x; e < E; e += 256) {
    float a = 0.f;
    for (int r = 0; r < m; ++r) a += f[static_cast<long long>(s * m + r) * E + e];
    S[static_cast<long long>(s) * E + e] = a;
    q = fmaf(a, a, q);
  }
  q = m_block_sum(q, red);
  if (threadIdx.x == 0) Snorm2[s] = q;
}

__global__ void __launch_bounds__(256) e2e_loss_kernel(const float* __restrict__ f, const float* __restrict__ S,
                                                       const float* __restrict__ Snorm2, int n, int m, int E, float scale,
                                                       float* __restrict__ loss) {
  pdl_entry();
  extern __shared__ float sim[];       // [n]
  __shared__ float red[8];
  const int i = blockIdx.x, own = i / m, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float* fi = f + static_cast<long long>(i) * E;
  for (int j = wp; j < n; j += 8) {
    const float* Sj = S + static_cast<long long>(j) * E;
    float d = 0.f, q = 0.f;
    for (int e = lane; e < E; e += 32) {
      d = fmaf(fi[e], Sj[e], d);
      q = fmaf(fi[e], fi[e], q);
    }
    d = m_warp_sum(d);
    q = m_warp_sum(q);
    if (lane == 0) {
      float v;
      if (j == own) {
        // centre of the speaker's OTHER segments: (S - f_i) / |S - f_i|  (a positive rescaling of the mean leaves it unchanged)
        const float n2 = Snorm2[j] - 2.f * d + q;
        v = (d - q) * rsqrtf(fmaxf(n2, MET_EPS));
      } else {
        // l2_scaling(mean): S / m normalised -- the 1e-12 floor acts on |S / m|^2
        const float mm = static_cast<float>(m);
        v = (d / mm) * rsqrtf(fmaxf(Snorm2[j] / (mm * mm), MET_EPS));
      }
      sim[j] = 20.0f * v;
    }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, sim[j]);
  mx = m_block_max(mx, red);
  float se = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) se += expf(sim[j] - mx);
  se = m_block_sum(se, red);
  if (threadIdx.x == 0) atomicAdd(loss, scale * (mx + logf(se) - sim[own]));
}

// ------------------------------------------------------------------------------------------------
// Generalized angular triplet loss against class centres (model/loss.py:708-901, loss_compute = "raw").  cos [B, ldc] =
// <f_i / |f_i|, w_j / |w_j|> comes from the tcgen05 GEMM; dist = 2 - 2 cos (unit vectors).  One block per sample:
//   target_i = dist[i, y_i], active_i = target_i > target_margin,
//   top-n = 1 / k: the n smallest non-target distances, n = 0: every non-target class;
//   triplet  = sum active_i max(margin + target_i - dist_ij, 1e-16) / (#(that > 1e-12) + 1e-12)
//   centre   = sum active_i target_i / (sum active_i + 1e-12)
// GRAD = false accumulates the four sums (counters[0..3]); GRAD = true reads them, adds the loss and writes d = dLoss/dcos.
struct GtCfg {
  float margin, target_margin, w_triplet, w_center;
  int topn;
};

template <bool GRAD>
__global__ void __launch_bounds__(256) gtriplet_rows_kernel(const float* __restrict__ cosm, const int* __restrict__ labels, int Cn,
                                                            long long ldc, GtCfg g, float scale, float* __restrict__ counters,
                                                            float* __restrict__ loss, __nv_bfloat16* __restrict__ d) {
  pdl_entry();
  extern __shared__ float sd[];          // [Cn] distances; selected entries are overwritten with +inf during the top-n search
  __shared__ float red[8];
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ float selv;
  __shared__ int seli;
  const int i = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int y = labels[i];
  const float* cr = cosm + static_cast<long long>(i) * ldc;
  for (int j = threadIdx.x; j < Cn; j += 256) sd[j] = 2.0f - 2.0f * cr[j];
  __syncthreads();
  const float td = sd[y];
  const float tm = (td > g.target_margin) ? 1.f : 0.f;
  const float eps = 1e-12f;
  float a = 0.f, cg = 0.f;
  __nv_bfloat16* dr = nullptr;
  if (GRAD) {
    a = scale * g.w_triplet / (counters[1] + eps);
    cg = scale * g.w_center / (counters[3] + eps);
    if (i == 0 && threadIdx.x == 0)
      atomicAdd(loss, scale * (g.w_triplet * counters[0] / (counters[1] + eps) + g.w_center * counters[2] / (counters[3] + eps)));
    if (d) {
      dr = d + static_cast<long long>(i) * ldc;
      for (int j = threadIdx.x; j < ldc; j += 256) dr[j] = __float2bfloat16(0.f);
    }
    __syncthreads();
  }
  float sum = 0.f, cnt = 0.f, nsel = 0.f;
  if (g.topn == 0) {
    for (int j = threadIdx.x; j < Cn; j += 256) {
      if (j == y) continue;
      const float t = g.margin + td - sd[j];
      const float tl = fmaxf(t, 1e-16f) * tm;
      sum += tl;
      cnt += (tl > eps) ? 1.f : 0.f;
      if (GRAD && tm > 0.f && t >= 1e-16f) {
        nsel += 1.f;
        if (dr) dr[j] = __float2bfloat16(2.0f * a);       // dLoss/ddist = -a, dist = 2 - 2 cos
      }
    }
  } else {
    __syncthreads();                                      // every thread has read its copy of the target distance
    if (threadIdx.x == 0) sd[y] = INFINITY;               // the target is pushed above the row maximum (loss.py:790)
    __syncthreads();
    for (int k = 0; k < g.topn; ++k) {
      float bv = INFINITY;
      int bi = 0x7fffffff;
      for (int j = threadIdx.x; j < Cn; j += 256)
        if (sd[j] < bv) { bv = sd[j]; bi = j; }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { rv[wp] = bv; ri[wp] = bi; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
          if (rv[w] < bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
        selv = bv;
        seli = bi;
        if (bi < Cn) sd[bi] = INFINITY;
      }
      __syncthreads();
      if (seli >= Cn) break;                               // fewer than top-n non-target classes
      if (threadIdx.x == 0) {
        const float t = g.margin + td - selv;
        const float tl = fmaxf(t, 1e-16f) * tm;
        sum += tl;
        cnt += (tl > eps) ? 1.f : 0.f;
        if (GRAD && tm > 0.f && t >= 1e-16f) {
          nsel += 1.f;
          if (dr) dr[seli] = __float2bfloat16(2.0f * a);
        }
      }
      __syncthreads();
    }
  }
  if (!GRAD) {
    sum = m_block_sum(sum, red);
    cnt = m_block_sum(cnt, red);
    if (threadIdx.x == 0) {
      if (sum != 0.f) atomicAdd(counters, sum);
      if (cnt != 0.f) atomicAdd(counters + 1, cnt);
      if (tm > 0.f) {
        atomicAdd(counters + 2, td);
        atomicAdd(counters + 3, 1.f);
      }
    }
  } else {
    nsel = m_block_sum(nsel, red);
    // dLoss/dtarget = a * (#selected negatives) + centre term; d cos = -2 d dist
    if (threadIdx.x == 0 && dr) dr[y] = __float2bfloat16(-2.0f * (a * nsel + cg * tm));
  }
}

// t[e] = sum_j w[e, j] inv_norm[j]  (sum of the normalised centres), one block per row e
__global__ void __launch_bounds__(256) centre_colsum_kernel(const float* __restrict__ w, const float* __restrict__ inv_norm,
                                                            float* __restrict__ t, int Cn, long long ldw) {
  pdl_entry();
  __shared__ float red[8];
  const int e = blockIdx.x;
  float s = 0.f;
  for (int j = threadIdx.x; j < Cn; j += 256) s = fmaf(w[static_cast<long long>(e) * ldw + j], inv_norm[j], s);
  s = m_block_sum(s, red);
  if (threadIdx.x == 0) t[e] = s;
}
// between = -mean_{i != j} |w_i - w_j|^2 = -2 + 2 (|t|^2 - C) / (C (C - 1))  for unit centres (loss.py:821-822)
__global__ void __launch_bounds__(256) centre_between_loss_kernel(const float* __restrict__ t, int E, int Cn, float coef,
                                                                  float* __restrict__ loss) {
  pdl_entry();
  __shared__ float red[8];
  float s = 0.f;
  for (int e = threadIdx.x; e < E; e += 256) s = fmaf(t[e], t[e], s);
  s = m_block_sum(s, red);
  const float cc = static_cast<float>(Cn) * static_cast<float>(Cn - 1);
  if (threadIdx.x == 0) atomicAdd(loss, coef * (-2.0f + 2.0f * (s - static_cast<float>(Cn)) / cc));
}
// dLoss/dwn[e, j] += coef * 4 / (C (C - 1)) * (t[e] - wn[e, j])
__global__ void __launch_bounds__(256) centre_between_bwd_kernel(float* __restrict__ dwn, const float* __restrict__ w,
                                                                 const float* __restrict__ inv_norm, const float* __restrict__ t,
                                                                 int E, int Cn, long long ldw, float coef4) {
  pdl_entry();
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(E) * Cn) return;
  const int e = static_cast<int>(idx / Cn), j = static_cast<int>(idx % Cn);
  const long long o = static_cast<long long>(e) * ldw + j;
  dwn[o] += coef4 * (t[e] - w[o] * inv_norm[j]);
}
// Moving-average centres (loss.py:766-783): delta_i = (w[:, y_i] - f_i) * decay from the OLD centres, then
// w[:, y_i] -= delta_i summed over the samples of a class (tf.scatter_nd adds repeated indices).
__global__ void __launch_bounds__(128) centre_delta_kernel(const float* __restrict__ w, const float* __restrict__ f,
                                                           const int* __restrict__ labels, float* __restrict__ delta, int E,
                                                           long long ldw, float decay) {
  pdl_entry();
  const int i = blockIdx.x, y = labels[i];
  for (int e = threadIdx.x; e < E; e += 128)
    delta[static_cast<long long>(i) * E + e] = (w[static_cast<long long>(e) * ldw + y] - f[static_cast<long long>(i) * E + e]) * decay;
}
__global__ void __launch_bounds__(128) centre_scatter_kernel(float* __restrict__ w, const float* __restrict__ delta,
                                                             const int* __restrict__ labels, int E, long long ldw) {
  pdl_entry();
  const int i = blockIdx.x, y = labels[i];
  for (int e = threadIdx.x; e < E; e += 128) atomicAdd(&w[static_cast<long long>(e) * ldw + y], -delta[static_cast<long long>(i) * E + e]);
}

}  // namespace xv

using namespace xv;

static int met_check(const char* who, int B) {
  if (B < 1 || B > MET_MAX_B) return set_error(XV_ERR_UNSUPPORTED, "%s: the batch must have 1..%d rows", who, MET_MAX_B);
  return XV_OK;
}

template <typename Kern>
static int met_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) XV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * MET_MAX_B * 4));
  return XV_OK;
}

extern "C" int xv_gram_f32(const float* x, float* gram, int B, int E, int64_t ldx, void* stream) {
  if (!x || !gram || E < 1 || ldx < E) return set_error(XV_ERR_INVALID, "xv_gram_f32: bad arguments");
  int rc = met_check("xv_gram_f32", B); if (rc) return rc;
  dim3 grid(ceil_div(B, 64), ceil_div(B, 64));
  ::xv::launch_pdl((sgemm_kernel<true>), grid, 256, 0, static_cast<cudaStream_t>(stream), x, x, gram, B, B, E,
                   static_cast<long long>(ldx), static_cast<long long>(ldx), static_cast<long long>(B),
                   static_cast<const float*>(nullptr), static_cast<const float*>(nullptr), 0LL);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_pairwise_bwd(const float* coef, const float* diag, const float* x, float* dx, int B, int E, int64_t ldx,
                               void* stream) {
  if (!coef || !diag || !x || !dx || E < 1 || ldx < E) return set_error(XV_ERR_INVALID, "xv_pairwise_bwd: bad arguments");
  int rc = met_check("xv_pairwise_bwd", B); if (rc) return rc;
  dim3 grid(ceil_div(E, 64), ceil_div(B, 64));
  ::xv::launch_pdl((sgemm_kernel<false>), grid, 256, 0, static_cast<cudaStream_t>(stream), coef, x, dx, B, E, B,
                   static_cast<long long>(B), static_cast<long long>(ldx), static_cast<long long>(ldx), diag, x,
                   static_cast<long long>(ldx));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_semihard_triplet(const float* gram, const int32_t* labels, int B, float margin, int squared, float scale,
                                   float* loss, float* coef, float* diag, float* work, void* stream) {
  if (!gram || !labels || !loss || !coef || !diag || !work) return set_error(XV_ERR_INVALID, "xv_semihard_triplet: bad arguments");
  int rc = met_check("xv_semihard_triplet", B); if (rc) return rc;
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  const long long BB = static_cast<long long>(B) * B;
  float* D = work;                 // [B, B] distances
  float* dD = work + BB;           // [B, B] dLoss/dD
  float* counter = work + 2 * BB;  // [1] positive pairs
  XV_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(float), s_));
  ::xv::launch_pdl((euclid_from_gram_kernel), ceil_div(BB, 256), 256, 0, s_, gram, D, B, squared);
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((count_positive_pairs_kernel), B, 256, 0, s_, labels, B, counter);
  XV_CUDA_CHECK(cudaGetLastError());
  const size_t smem = static_cast<size_t>(3) * B * sizeof(float);
  rc = met_smem(semihard_kernel, smem); if (rc) return rc;
  ::xv::launch_pdl((semihard_kernel), B, 256, smem, s_, static_cast<const float*>(D), labels, B, margin, scale,
                   static_cast<const float*>(counter), loss, dD);
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((euclid_finish_kernel), B, 256, 0, s_, gram, static_cast<const float*>(dD), coef, diag, B, squared);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_angular_triplet(const float* gram, const int32_t* labels, int B, int kind, float margin, int hard, float scale,
                                  float* loss, float* coef, float* diag, float* work, void* stream) {
  if (!gram || !labels || !loss || !coef || !diag || !work || kind < 0 || kind > 2)
    return set_error(XV_ERR_INVALID, "xv_angular_triplet: bad arguments");
  int rc = met_check("xv_angular_triplet", B); if (rc) return rc;
  AngCfg a;
  a.kind = kind;
  a.margin = margin;
  a.m_int = static_cast<int>(margin);
  a.cos_m = cosf(margin);
  a.sin_m = sinf(margin);
  a.threshold = cosf(3.14159265358979323846f - margin);
  if (kind == ANG_ASOFTMAX && a.m_int != 1 && a.m_int != 2 && a.m_int != 4)
    return set_error(XV_ERR_UNSUPPORTED, "xv_angular_triplet: asoftmax margin must be 1, 2 or 4 (loss.py:537-550)");
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  const long long BB = static_cast<long long>(B) * B;
  float* dS = work;                // [B, B] dLoss/dcos
  float* counters = work + BB;     // [2] sum, count
  const size_t smem = static_cast<size_t>(5) * B * sizeof(float);
  if (hard) {
    rc = met_smem(angular_hard_kernel, smem); if (rc) return rc;
    ::xv::launch_pdl((angular_hard_kernel), B, 256, smem, s_, gram, labels, B, a, scale, loss, dS);
    XV_CUDA_CHECK(cudaGetLastError());
  } else {
    XV_CUDA_CHECK(cudaMemsetAsync(counters, 0, 2 * sizeof(float), s_));
    rc = met_smem(angular_all_count_kernel, smem); if (rc) return rc;
    rc = met_smem(angular_all_grad_kernel, smem); if (rc) return rc;
    ::xv::launch_pdl((angular_all_count_kernel), B, 256, smem, s_, gram, labels, B, a, counters);
    XV_CUDA_CHECK(cudaGetLastError());
    ::xv::launch_pdl((angular_all_grad_kernel), B, 256, smem, s_, gram, labels, B, a, scale, static_cast<const float*>(counters),
                     loss, dS);
    XV_CUDA_CHECK(cudaGetLastError());
  }
  ::xv::launch_pdl((cos_finish_kernel), B, 256, 0, s_, gram, static_cast<const float*>(dS), coef, diag, B);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_e2e_valid_loss(const float* x, int num_speakers, int num_segments, int E, int64_t ldx, float scale, float* loss,
                                 float* work, void* stream) {
  if (!x || !loss || !work || num_speakers < 1 || num_segments < 2 || E < 1 || ldx < E)
    return set_error(XV_ERR_INVALID, "xv_e2e_valid_loss: bad arguments (>= 2 segments per speaker)");
  if (num_speakers > 8192) return set_error(XV_ERR_UNSUPPORTED, "xv_e2e_valid_loss: at most 8192 speakers per batch");
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  const long long Bn = static_cast<long long>(num_speakers) * num_segments;
  float* f = work;                                   // [B, E]
  float* S = work + Bn * E;                          // [n, E]
  float* Sn = S + static_cast<long long>(num_speakers) * E;    // [n]
  ::xv::launch_pdl((e2e_prepare_kernel), num_speakers, 256, 0, s_, x, f, S, Sn, num_segments, E, static_cast<long long>(ldx));
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((e2e_loss_kernel), static_cast<int>(Bn), 256, static_cast<size_t>(num_speakers) * sizeof(float), s_,
                   static_cast<const float*>(f), static_cast<const float*>(S), static_cast<const float*>(Sn), num_speakers,
                   num_segments, E, scale / static_cast<float>(Bn), loss);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_center_triplet(const float* cosm, const int32_t* labels, int B, int C, int64_t ldc, float margin, float target_margin,
                                 int topn, float w_triplet, float w_center, float scale, float* loss, void* d, float* counters,
                                 void* stream) {
  if (!cosm || !labels || !loss || !counters || B < 1 || C < 2 || ldc < C || topn < 0 || topn >= C)
    return set_error(XV_ERR_INVALID, "xv_center_triplet: bad arguments (0 <= triplet_topn < classes)");
  const size_t smem = static_cast<size_t>(C) * sizeof(float);
  if (smem > 200 * 1024) return set_error(XV_ERR_UNSUPPORTED, "xv_center_triplet: at most 51200 classes");
  if (smem > 48 * 1024) {
    XV_CUDA_CHECK(cudaFuncSetAttribute(gtriplet_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    XV_CUDA_CHECK(cudaFuncSetAttribute(gtriplet_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  GtCfg g;
  g.margin = margin; g.target_margin = target_margin; g.w_triplet = w_triplet; g.w_center = w_center; g.topn = topn;
  XV_CUDA_CHECK(cudaMemsetAsync(counters, 0, 4 * sizeof(float), s_));
  ::xv::launch_pdl((gtriplet_rows_kernel<false>), B, 256, smem, s_, cosm, labels, C, static_cast<long long>(ldc), g, scale, counters,
                   loss, static_cast<__nv_bfloat16*>(nullptr));
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((gtriplet_rows_kernel<true>), B, 256, smem, s_, cosm, labels, C, static_cast<long long>(ldc), g, scale, counters,
                   loss, static_cast<__nv_bfloat16*>(d));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_center_between(const float* w, const float* inv_norm, int E, int C, int64_t ldw, float coef, float* t, float* loss,
                                 void* stream) {
  if (!w || !inv_norm || !t || !loss || E < 1 || C < 2 || ldw < C) return set_error(XV_ERR_INVALID, "xv_center_between: bad arguments");
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  ::xv::launch_pdl((centre_colsum_kernel), E, 256, 0, s_, w, inv_norm, t, C, static_cast<long long>(ldw));
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((centre_between_loss_kernel), 1, 256, 0, s_, static_cast<const float*>(t), E, C, coef, loss);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_center_between_bwd(float* dwn, const float* w, const float* inv_norm, const float* t, int E, int C, int64_t ldw,
                                     float coef, void* stream) {
  if (!dwn || !w || !inv_norm || !t || E < 1 || C < 2 || ldw < C) return set_error(XV_ERR_INVALID, "xv_center_between_bwd: bad arguments");
  const float coef4 = coef * 4.0f / (static_cast<float>(C) * static_cast<float>(C - 1));
  ::xv::launch_pdl((centre_between_bwd_kernel), ceil_div(static_cast<long long>(E) * C, 256), 256, 0, static_cast<cudaStream_t>(stream),
                   dwn, w, inv_norm, t, E, C, static_cast<long long>(ldw), coef4);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_center_update(float* w, const float* feats, const int32_t* labels, float* delta, int B, int E, int64_t ldw,
                                float decay, void* stream) {
  if (!w || !feats || !labels || !delta || B < 1 || E < 1 || ldw < 1) return set_error(XV_ERR_INVALID, "xv_center_update: bad arguments");
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  ::xv::launch_pdl((centre_delta_kernel), B, 128, 0, s_, static_cast<const float*>(w), feats, labels, delta, E,
                   static_cast<long long>(ldw), decay);
  XV_CUDA_CHECK(cudaGetLastError());
  ::xv::launch_pdl((centre_scatter_kernel), B, 128, 0, s_, w, static_cast<const float*>(delta), labels, E, static_cast<long long>(ldw));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
