// Host-feeder side of the path: Kaldi compressed-matrix ('CM ', format 1) segments are shipped to the GPU as raw uint8
// crops and dequantised + transposed here, instead of on the loader processes' CPUs
// (reference dataset/kaldi_io.py:780-797 uint16/uint8 -> float maps, :852-868 column-major crop, `mat.T`).
// Bit-exact with the reference reader: every operation is a single correctly-rounded float32 operation in the order the
// NumPy expressions evaluate (no FMA contraction: explicit __f*_rn intrinsics; IEEE division).
#include <cstdint>

#include "xv_internal.h"

namespace xv {

// data   u8  [B, D, ld_t]   segment b, feature column d: T consecutive frames (the stored column-major order)
// hdr    u16 [B, D, 4]      percentile_0 / 25 / 75 / 100 of column d
// glob   f32 [B, 2]         (min_value, range) of the matrix segment b was cut from
// out    f32 [B, T, ldo]    row-major frames (the reference's mat.T)
// grid = (ceil(T/32), ceil(D/32), B), block = (32, 8): a 32-frame x 32-column tile goes through shared memory so that
// both the byte reads (frames contiguous) and the float writes (columns contiguous) are coalesced.
__global__ void __launch_bounds__(256) cm_decode_kernel(const uint8_t* __restrict__ data, const uint16_t* __restrict__ hdr,
                                                        const float* __restrict__ glob, float* __restrict__ out, int T, int D,
                                                        long long ld_t, long long ldo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float gmin = glob[2 * b], grange = glob[2 * b + 1];
  const float step = __fmul_rn(grange, 1.52590218966964e-05f);            // range * 1.52590218966964e-05
  for (int dd = ty; dd < 32; dd += 8) {
    const int d = d0 + dd, t = t0 + tx;
    float r = 0.f;
    if (d < D && t < T) {
      const uint2 hw = *reinterpret_cast<const uint2*>(hdr + (static_cast<long long>(b) * D + d) * 4);
      const float p0 = __fadd_rn(gmin, __fmul_rn(step, static_cast<float>(hw.x & 0xffffu)));
      const float p25 = __fadd_rn(gmin, __fmul_rn(step, static_cast<float>(hw.x >> 16)));
      const float p75 = __fadd_rn(gmin, __fmul_rn(step, static_cast<float>(hw.y & 0xffffu)));
      const float p100 = __fadd_rn(gmin, __fmul_rn(step, static_cast<float>(hw.y >> 16)));
      const unsigned int q = data[(static_cast<long long>(b) * D + d) * ld_t + t];
      const float v = static_cast<float>(q);
      if (q <= 64u) r = __fadd_rn(p0, __fmul_rn(__fdiv_rn(__fsub_rn(p25, p0), 64.f), v));
      else if (q <= 192u) r = __fadd_rn(p25, __fmul_rn(__fdiv_rn(__fsub_rn(p75, p25), 128.f), __fsub_rn(v, 64.f)));
      else r = __fadd_rn(p75, __fmul_rn(__fdiv_rn(__fsub_rn(p100, p75), 63.f), __fsub_rn(v, 192.f)));
    }
    tile[dd][tx] = r;
  }
  __syncthreads();
  for (int tt = ty; tt < 32; tt += 8) {
    const int t = t0 + tt, d = d0 + tx;
    if (t < T && d < D) out[(static_cast<long long>(b) * T + t) * ldo + d] = tile[tx][tt];
  }
}

}  // namespace xv

using namespace xv;

extern "C" int xv_cm_decode(const void* data, const void* headers, const float* glob, float* out, int B, int T, int D,
                            int64_t ld_t, int64_t ldo, void* stream) {
  if (!data || !headers || !glob || !out || B <= 0 || T <= 0 || D <= 0 || ld_t < T || ldo < D || B > 65535)
    return set_error(XV_ERR_INVALID, "xv_cm_decode: bad arguments (B in [1, 65535], ld_t >= T, ldo >= D)");
  const dim3 grid(ceil_div(T, 32), ceil_div(D, 32), B), block(32, 8);
  cm_decode_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(data), static_cast<const uint16_t*>(headers), glob, out, T, D, ld_t, ldo);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
