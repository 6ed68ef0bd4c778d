// Auxiliary losses of the margin heads (model/loss.py:985-1037): ring loss and minimum hyperspherical energy (MHE).
// Both are tiny next to the head GEMMs and need no contraction of their own:
//   ring  L = lambda * mean_i (||x_i|| - r)^2 : the head's feature preparation already produced ||x_i||; the gradient
//         enters the existing dLoss/d||x_i|| path (gnorm, consumed by xv_head_finish_dx) and d/dr is a scalar.
//   MHE   L = lambda / (mean_{i,j} (2 - 2 <wn_{y_i}, wn_j>) + 1e-6) with wn the column-normalised speaker matrix.  The
//         double sum factorises, sum_{i,j} <wn_{y_i}, wn_j> = <S, t> with S = sum_i wn_{y_i} and t = sum_j wn_j, so the
//         reference's [B, C] matmul (loss.py:1029) becomes two [E] vectors; dL/dwn_j = kappa * (h_j t + S) (h = label
//         histogram) is a rank-2 update of the dWn buffer that xv_head_finish_dw then takes through the normalisation.
#include "xv_internal.h"

namespace xv {

__device__ __forceinline__ float aux_block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += sh[i];
  return t;
}

// One block.  loss != NULL: loss += scale_loss * sum_i (n_i - r)^2.  gnorm != NULL: gnorm[i] += 2 * scale_grad * (n_i - r)
// and dr += -2 * scale_grad * sum_i (n_i - r).   (scale = lambda / global batch.)
__global__ void __launch_bounds__(256) ring_loss_kernel(const float* __restrict__ xnorm, const float* __restrict__ r, int B,
                                                       float scale, float* loss, float* gnorm, float* dr) {
  pdl_entry();
  __shared__ float sh[32];
  const float rv = r[0];
  float sq = 0.f, sd = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float d = xnorm[i] - rv;
    sq += d * d;
    sd += d;
    if (gnorm) gnorm[i] += 2.f * scale * d;
  }
  sq = aux_block_sum(sq, sh);
  sd = aux_block_sum(sd, sh);
  if (threadIdx.x == 0) {
    if (loss) atomicAdd(loss, scale * sq);
    if (dr) atomicAdd(dr, -2.f * scale * sd);
  }
}

// One block per embedding row e: t[e] = sum_j w[e, j] * inv_norm[j], S[e] = sum_i w[e, y_i] * inv_norm[y_i];
// block 0 also builds the label histogram hist[j] (zero on entry).
__global__ void __launch_bounds__(256) mhe_rows_kernel(const float* __restrict__ w, const float* __restrict__ inv_norm,
                                                      const int* __restrict__ labels, int B, int C, long long ldw,
                                                      float* __restrict__ t, float* __restrict__ S, float* hist) {
  pdl_entry();
  __shared__ float sh[32];
  const int e = blockIdx.x;
  const float* row = w + static_cast<long long>(e) * ldw;
  float a = 0.f, b = 0.f;
  for (int j = threadIdx.x; j < C; j += blockDim.x) a = fmaf(row[j], inv_norm[j], a);
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int y = labels[i];
    if (y >= 0 && y < C) {
      b = fmaf(row[y], inv_norm[y], b);
      if (e == 0) atomicAdd(hist + y, 1.0f);
    }
  }
  a = aux_block_sum(a, sh);
  b = aux_block_sum(b, sh);
  if (threadIdx.x == 0) {
    t[e] = a;
    S[e] = b;
  }
}

// One block: m = 2 - 2 <S, t> / (B C); loss += scale * lambda / (m + 1e-6); kappa[0] = scale * 2 lambda / ((m+1e-6)^2 B C).
__global__ void __launch_bounds__(256) mhe_scalar_kernel(const float* __restrict__ t, const float* __restrict__ S, int E, int B,
                                                        int C, float lambda, float scale, float* loss, float* kappa) {
  pdl_entry();
  __shared__ float sh[32];
  float d = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) d = fmaf(S[e], t[e], d);
  d = aux_block_sum(d, sh);
  if (threadIdx.x == 0) {
    const float bc = static_cast<float>(B) * static_cast<float>(C);
    const float m = 2.f - 2.f * d / bc + 1e-6f;
    if (loss) atomicAdd(loss, scale * lambda / m);
    kappa[0] = scale * 2.f * lambda / (m * m * bc);
  }
}

// dWn[e, j] += kappa * (hist[j] * t[e] + S[e])
__global__ void __launch_bounds__(256) mhe_grad_kernel(float* __restrict__ dwn, const float* __restrict__ t,
                                                      const float* __restrict__ S, const float* __restrict__ hist,
                                                      const float* __restrict__ kappa, int E, int C, long long ldw) {
  pdl_entry();
  const float k = kappa[0];
  const long long total = static_cast<long long>(E) * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(idx / C), j = static_cast<int>(idx - static_cast<long long>(e) * C);
    dwn[static_cast<long long>(e) * ldw + j] += k * fmaf(hist[j], t[e], S[e]);
  }
}

}  // namespace xv

using namespace xv;

extern "C" int xv_ring_loss(const float* xnorm, const float* r, int B, float scale, float* loss, float* gnorm, float* dr,
                            void* stream) {
  if (!xnorm || !r || B <= 0) return set_error(XV_ERR_INVALID, "xv_ring_loss: bad arguments");
  ::xv::launch_pdl((ring_loss_kernel), 1, 256, 0, static_cast<cudaStream_t>(stream), xnorm, r, B, scale, loss, gnorm, dr);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_mhe_forward(const float* w, const float* inv_norm, const int32_t* labels, int B, int E, int C, int64_t ldw,
                              float lambda, float scale, float* t, float* S, float* hist, float* kappa, float* loss,
                              void* stream) {
  if (!w || !inv_norm || !labels || !t || !S || !hist || !kappa || B <= 0 || E <= 0 || C <= 0 || ldw < C)
    return set_error(XV_ERR_INVALID, "xv_mhe_forward: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ::xv::launch_pdl((mhe_rows_kernel), E, 256, 0, s, w, inv_norm, labels, B, C, static_cast<long long>(ldw), t, S, hist);
  ::xv::launch_pdl((mhe_scalar_kernel), 1, 256, 0, s, static_cast<const float*>(t), static_cast<const float*>(S), E, B, C, lambda,
                   scale, loss, kappa);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_mhe_backward(float* dwn, const float* t, const float* S, const float* hist, const float* kappa, int E, int C,
                               int64_t ldw, void* stream) {
  if (!dwn || !t || !S || !hist || !kappa || E <= 0 || C <= 0 || ldw < C)
    return set_error(XV_ERR_INVALID, "xv_mhe_backward: bad arguments");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  const long long total = static_cast<long long>(E) * C;
  long long g = (total + 255) / 256;
  if (g > 8LL * sms) g = 8LL * sms;
  ::xv::launch_pdl((mhe_grad_kernel), static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream), dwn, t, S, hist, kappa, E,
                   C, static_cast<long long>(ldw));
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
