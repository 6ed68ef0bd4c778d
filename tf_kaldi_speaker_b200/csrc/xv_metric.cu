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
