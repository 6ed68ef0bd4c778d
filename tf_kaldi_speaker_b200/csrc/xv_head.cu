// Small kernels around the fused margin-softmax GEMM epilogues (model/loss.py:97-159, 207-247, 293-345,
// model/common.py:45-58): weight/feature normalisation into bf16 GEMM operands, the log-sum-exp combine of
// the per-tile partials, and the chain rule back through the two normalisations.
#include <cuda_bf16.h>

#include "xv_internal.h"

namespace xv {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += sh[i];
  return t;
}

// w f32 [E, C] (TF layout, C contiguous) -> wn3 bf16 [3E, ldw] = [hi(wn); lo(wn); hi(wn)], wn = w * rsqrt(max(sum_e w^2, 1e-12))
// (tf.nn.l2_normalize(w, dim=0), loss.py:104,213,299).  normalize = 0 keeps w (plain softmax head).
// Block = 32 x HROWG threads: 64 columns (two per thread, 8-byte loads) x HROWG row groups; column sums go through smem.
// (8 row groups = 113 blocks x 8 warps for 7200 speakers left 90 % of the SM's warp slots empty: 28 us for 37 MB.)
constexpr int HROWG = 32;
__global__ void __launch_bounds__(32 * HROWG) head_prep_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wn3,
                                                                float* __restrict__ inv_norm, int E, int C, long long ldw,
                                                                int normalize) {
  pdl_entry();
  __shared__ float red[HROWG][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + tx * 2;
  const bool ok = c < C;                      // C is even (padded to a multiple of 8)
  float s0 = 0.f, s1 = 0.f;
  if (normalize && ok) {
    for (int e = ty; e < E; e += HROWG) {
      const float2 v = *reinterpret_cast<const float2*>(w + static_cast<long long>(e) * C + c);
      s0 += v.x * v.x;
      s1 += v.y * v.y;
    }
  }
  red[ty][tx * 2] = s0;
  red[ty][tx * 2 + 1] = s1;
  __syncthreads();
  float i0 = 1.f, i1 = 1.f;
  if (normalize) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < HROWG; ++k) { a0 += red[k][tx * 2]; a1 += red[k][tx * 2 + 1]; }
    i0 = rsqrtf(fmaxf(a0, 1e-12f));
    i1 = rsqrtf(fmaxf(a1, 1e-12f));
  }
  if (!ok) return;
  if (inv_norm && ty == 0) { inv_norm[c] = i0; inv_norm[c + 1] = i1; }
  for (int e = ty; e < E; e += HROWG) {
    const float2 v = *reinterpret_cast<const float2*>(w + static_cast<long long>(e) * C + c);
    const float v0 = v.x * i0, v1 = v.y * i1;
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
    *reinterpret_cast<__nv_bfloat162*>(wn3 + static_cast<long long>(e) * ldw + c) = h;
    *reinterpret_cast<__nv_bfloat162*>(wn3 + static_cast<long long>(E + e) * ldw + c) = l;
    *reinterpret_cast<__nv_bfloat162*>(wn3 + static_cast<long long>(2 * E + e) * ldw + c) = h;
  }
}

// One block per row: optional l2_scaling (trainer.py:183-186 / common.py:45-58), ||x||, bf16 [hi|hi|lo] split.
__global__ void head_prep_features_kernel(const float* __restrict__ u, float scaling, float* __restrict__ x,
                                          __nv_bfloat16* __restrict__ x3, float* __restrict__ xnorm,
                                          float* __restrict__ u_rinv, int E) {
  pdl_entry();
  __shared__ float sh[32];
  const int i = blockIdx.x;
  const float* ur = u + static_cast<long long>(i) * E;
  float ss = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) ss += ur[e] * ur[e];
  ss = block_sum(ss, sh);
  float mul = 1.f;
  if (scaling > 0.f) {
    const float rinv = rsqrtf(fmaxf(ss, 1e-12f));
    mul = rinv * scaling;
    if (threadIdx.x == 0 && u_rinv) u_rinv[i] = rinv;
  }
  float s2 = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float v = ur[e] * mul;
    s2 += v * v;
    if (x) x[static_cast<long long>(i) * E + e] = v;
    const __nv_bfloat16 h = __float2bfloat16(v);
    __nv_bfloat16* row = x3 + static_cast<long long>(i) * 3 * E;
    row[e] = h;
    row[E + e] = h;
    row[2 * E + e] = __float2bfloat16(v - __bfloat162float(h));
  }
  s2 = block_sum(s2, sh);
  if (threadIdx.x == 0 && xnorm) xnorm[i] = fmaxf(sqrtf(s2), 1e-12f);
}

// lse_i = log sum_j exp(z'_ij) from the per-N-tile (max, sum) partials; loss += sum_i (lse_i - z'_{i,y_i}) * inv_batch
// One warp per batch row (lanes stride over the partials; the first version walked them serially per thread).
__global__ void __launch_bounds__(128) head_combine_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                                           const float* __restrict__ target_logit, int nblk, int B,
                                                           float inv_batch, float* __restrict__ lse,
                                                           float* __restrict__ loss_rows, float* loss) {
  pdl_entry();
  __shared__ float sh[32];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float li = 0.f;
  if (i < B) {
    float gmax = -INFINITY;
    for (int k = lane; k < nblk; k += 32) gmax = fmaxf(gmax, part_max[static_cast<long long>(k) * B + i]);
    for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    float s = 0.f;
    for (int k = lane; k < nblk; k += 32)
      s += part_sum[static_cast<long long>(k) * B + i] * expf(part_max[static_cast<long long>(k) * B + i] - gmax);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float l = logf(s) + gmax;
    if (lane == 0) {
      lse[i] = l;
      li = l - target_logit[i];
      if (loss_rows) loss_rows[i] = li;
    }
  }
  const float tot = block_sum(li, sh);
  if (threadIdx.x == 0) atomicAdd(loss, tot * inv_batch);
}

// dx = dx_gemm + gnorm * x/||x||, then (optionally) back through x = s * u / ||u||.
__global__ void head_finish_dx_kernel(const float* __restrict__ dxg, const float* __restrict__ gnorm,
                                      const float* __restrict__ x, const float* __restrict__ xnorm,
                                      const float* __restrict__ u, const float* __restrict__ u_rinv, float scaling,
                                      float* __restrict__ du, int E) {
  pdl_entry();
  __shared__ float sh[32];
  const int i = blockIdx.x;
  const long long off = static_cast<long long>(i) * E;
  const float gn = gnorm ? gnorm[i] : 0.f;
  const float xn = xnorm ? xnorm[i] : 1.f;
  const float gcoef = (gnorm && xn > 1e-12f) ? gn / xn : 0.f;   // maximum(norm, eps): no gradient to x when clamped
  if (scaling <= 0.f) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) du[off + e] = dxg[off + e] + gcoef * x[off + e];
    return;
  }
  const float rinv = u_rinv[i];
  float dot = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float dxe = dxg[off + e] + gcoef * x[off + e];
    dot += dxe * u[off + e];
  }
  dot = block_sum(dot, sh);
  // x = s*u*rinv, rinv = (max(|u|^2, eps))^-1/2 : dx/du = s*rinv*(I - u u^T rinv^2) when |u|^2 > eps
  const bool clamped = (rinv >= 1e6f);   // rsqrt(1e-12)
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float dxe = dxg[off + e] + gcoef * x[off + e];
    float g = scaling * rinv * dxe;
    if (!clamped) g -= scaling * rinv * rinv * rinv * dot * u[off + e];
    du[off + e] = g;
  }
}

// dW_j = (dWn_j - wn_j <wn_j, dWn_j>) * inv_norm_j   (in place on the gradient buffer); same 64 x 8 blocking.
__global__ void __launch_bounds__(32 * HROWG) head_finish_dw_kernel(float* __restrict__ dw, const float* __restrict__ w,
                                                                    const float* __restrict__ inv_norm, int E, int C) {
  pdl_entry();
  __shared__ float red[HROWG][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + tx * 2;
  const bool ok = c < C;
  float i0 = 1.f, i1 = 1.f, d0 = 0.f, d1 = 0.f;
  if (ok) {
    i0 = inv_norm[c];
    i1 = inv_norm[c + 1];
    for (int e = ty; e < E; e += HROWG) {
      const long long idx = static_cast<long long>(e) * C + c;
      const float2 wv = *reinterpret_cast<const float2*>(w + idx);
      const float2 gv = *reinterpret_cast<const float2*>(dw + idx);
      d0 += wv.x * i0 * gv.x;
      d1 += wv.y * i1 * gv.y;
    }
  }
  red[ty][tx * 2] = d0;
  red[ty][tx * 2 + 1] = d1;
  __syncthreads();
  if (!ok) return;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int k = 0; k < HROWG; ++k) { a0 += red[k][tx * 2]; a1 += red[k][tx * 2 + 1]; }
  const bool cl0 = (i0 >= 1e6f), cl1 = (i1 >= 1e6f);     // rsqrt(1e-12): the norm was clamped, no projection term
  for (int e = ty; e < E; e += HROWG) {
    const long long idx = static_cast<long long>(e) * C + c;
    const float2 wv = *reinterpret_cast<const float2*>(w + idx);
    float2 gv = *reinterpret_cast<const float2*>(dw + idx);
    gv.x = gv.x * i0 - (cl0 ? 0.f : wv.x * i0 * i0 * a0);
    gv.y = gv.y * i1 - (cl1 ? 0.f : wv.y * i1 * i1 * a1);
    *reinterpret_cast<float2*>(dw + idx) = gv;
  }
}


// ---- class-sharded head (the speaker matrix split by columns over the data-parallel ranks) --------------------------
// Global labels -> labels local to the shard that owns columns [lo, lo + n_local): out-of-shard rows get -1, which no
// column of the fused GEMM epilogues matches (no margin, no target logit, no ||x|| gradient from this shard).
__global__ void head_local_labels_kernel(const int* __restrict__ labels, int lo, int n_local, int* __restrict__ out, int R) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const int l = labels[i] - lo;
  out[i] = (l >= 0 && l < n_local) ? l : -1;
}

// This shard's per-row (max, sum exp(z' - max), target logit or 0) from the per-tile partials of the head-forward GEMM:
// out f32 [3, R].  One warp per row, as head_combine_kernel.
__global__ void __launch_bounds__(128) head_shard_partials_kernel(const float* __restrict__ part_max,
                                                                  const float* __restrict__ part_sum,
                                                                  const float* __restrict__ target_logit, int nblk, int R,
                                                                  float* __restrict__ out) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= R) return;
  float gmax = -INFINITY;
  for (int k = lane; k < nblk; k += 32) gmax = fmaxf(gmax, part_max[static_cast<long long>(k) * R + i]);
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  float s = 0.f;
  for (int k = lane; k < nblk; k += 32)
    s += part_sum[static_cast<long long>(k) * R + i] * expf(part_max[static_cast<long long>(k) * R + i] - gmax);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    out[i] = gmax;
    out[R + i] = s;
    out[2 * R + i] = target_logit[i];     // zero unless this shard owns the row's label
  }
}

// The all-gathered shard partials [S, 3, R] -> lse_i = log sum_s sum_s exp(max_s - gmax) + gmax (the all-reduce(max) /
// all-reduce(sum) pair of a class-sharded softmax, done locally on the gathered (max, sum) pairs), target logit =
// sum over shards (one owner), loss += sum_i (lse_i - target_i) * inv_batch.  One thread per row.
__global__ void __launch_bounds__(128) head_combine_shards_kernel(const float* __restrict__ parts, int S, int R,
                                                                  float inv_batch, float* __restrict__ lse,
                                                                  float* __restrict__ loss_rows, float* loss) {
  pdl_entry();
  __shared__ float sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float li = 0.f;
  if (i < R) {
    float gmax = -INFINITY;
    for (int s = 0; s < S; ++s) gmax = fmaxf(gmax, parts[(static_cast<long long>(s) * 3) * R + i]);
    float acc = 0.f, tgt = 0.f;
    for (int s = 0; s < S; ++s) {
      const float* ps = parts + (static_cast<long long>(s) * 3) * R;
      acc += ps[R + i] * expf(ps[i] - gmax);
      tgt += ps[2 * R + i];
    }
    const float l = logf(acc) + gmax;
    lse[i] = l;
    li = l - tgt;
    if (loss_rows) loss_rows[i] = li;
  }
  const float tot = block_sum(li, sh);
  if (threadIdx.x == 0) atomicAdd(loss, tot * inv_batch);
}

}  // namespace xv

using namespace xv;

extern "C" int xv_head_prep_weights(const float* w, void* wn3, float* inv_norm, int E, int C, int64_t ldw, int normalize,
                                    void* stream) {
  if (!w || !wn3 || E <= 0 || C <= 0 || (C & 1) || ldw < C || ldw % 8) return set_error(XV_ERR_INVALID, "xv_head_prep_weights: bad arguments (C must be even, ldw a multiple of 8)");
  ::xv::launch_pdl((head_prep_weights_kernel), ceil_div(C, 64), 32 * HROWG, 0, static_cast<cudaStream_t>(stream), 
      w, static_cast<__nv_bfloat16*>(wn3), inv_norm, E, C, ldw, normalize);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_head_prep_features(const float* u, float scaling, float* x, void* x3, float* xnorm, float* u_rinv,
                                     int B, int E, void* stream) {
  if (!u || !x3 || B <= 0 || E <= 0) return set_error(XV_ERR_INVALID, "xv_head_prep_features: bad arguments");
  if (scaling > 0.f && !u_rinv) return set_error(XV_ERR_INVALID, "xv_head_prep_features: feature_norm needs u_rinv");
  ::xv::launch_pdl((head_prep_features_kernel), B, 128, 0, static_cast<cudaStream_t>(stream), u, scaling, x, static_cast<__nv_bfloat16*>(x3),
                                                                              xnorm, u_rinv, E);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_head_combine(const float* part_max, const float* part_sum, const float* target_logit, int nblk, int B,
                               float inv_batch, float* lse, float* loss_rows, float* loss, void* stream) {
  if (!part_max || !part_sum || !target_logit || !lse || !loss || nblk <= 0 || B <= 0)
    return set_error(XV_ERR_INVALID, "xv_head_combine: bad arguments");
  ::xv::launch_pdl((head_combine_kernel), ceil_div(B, 4), 128, 0, static_cast<cudaStream_t>(stream), part_max, part_sum, target_logit,
                                                                                      nblk, B, inv_batch, lse, loss_rows, loss);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_head_finish_dx(const float* dx_gemm, const float* gnorm, const float* x, const float* xnorm,
                                 const float* u, const float* u_rinv, float scaling, float* du, int B, int E, void* stream) {
  if (!dx_gemm || !du || B <= 0 || E <= 0) return set_error(XV_ERR_INVALID, "xv_head_finish_dx: bad arguments");
  if (gnorm && (!x || !xnorm)) return set_error(XV_ERR_INVALID, "xv_head_finish_dx: margin term needs x and xnorm");
  if (scaling > 0.f && (!u || !u_rinv || !x)) return set_error(XV_ERR_INVALID, "xv_head_finish_dx: feature_norm needs u, u_rinv, x");
  ::xv::launch_pdl((head_finish_dx_kernel), B, 128, 0, static_cast<cudaStream_t>(stream), dx_gemm, gnorm, x, xnorm, u, u_rinv, scaling, du, E);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_head_finish_dw(float* dw, const float* w, const float* inv_norm, int E, int C, void* stream) {
  if (!dw || !w || !inv_norm || E <= 0 || C <= 0 || (C & 1)) return set_error(XV_ERR_INVALID, "xv_head_finish_dw: bad arguments (C must be even)");
  ::xv::launch_pdl((head_finish_dw_kernel), ceil_div(C, 64), 32 * HROWG, 0, static_cast<cudaStream_t>(stream), dw, w, inv_norm, E, C);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_head_local_labels(const int32_t* labels, int lo, int n_local, int32_t* out, int R, void* stream) {
  if (!labels || !out || R <= 0 || n_local <= 0 || lo < 0) return set_error(XV_ERR_INVALID, "xv_head_local_labels: bad arguments");
  ::xv::launch_pdl((head_local_labels_kernel), ceil_div(R, 128), 128, 0, static_cast<cudaStream_t>(stream), labels, lo, n_local, out, R);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_head_shard_partials(const float* part_max, const float* part_sum, const float* target_logit, int nblk,
                                      int R, float* out, void* stream) {
  if (!part_max || !part_sum || !target_logit || !out || nblk <= 0 || R <= 0)
    return set_error(XV_ERR_INVALID, "xv_head_shard_partials: bad arguments");
  ::xv::launch_pdl((head_shard_partials_kernel), ceil_div(R, 4), 128, 0, static_cast<cudaStream_t>(stream), part_max, part_sum,
                   target_logit, nblk, R, out);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_head_combine_shards(const float* parts, int S, int R, float inv_batch, float* lse, float* loss_rows,
                                      float* loss, void* stream) {
  if (!parts || !lse || !loss || S <= 0 || R <= 0) return set_error(XV_ERR_INVALID, "xv_head_combine_shards: bad arguments");
  ::xv::launch_pdl((head_combine_shards_kernel), ceil_div(R, 128), 128, 0, static_cast<cudaStream_t>(stream), parts, S, R, inv_batch,
                   lse, loss_rows, loss);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
