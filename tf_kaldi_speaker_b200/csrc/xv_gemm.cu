// Host side of the tcgen05 implicit-GEMM: tensor-map encoding, argument checks, dispatch.
#include <stdlib.h>

#include "xv_gemm_kernel.cuh"

namespace xv {

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

static int make_tmap(CUtensorMap* map, const xv_operand& op, int box_rows_kmajor) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(XV_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if (op.ptr == nullptr || (reinterpret_cast<uintptr_t>(op.ptr) & 15)) return set_error(XV_ERR_INVALID, "operand pointer null or not 16B aligned");
  if (op.ld % 8 != 0 || op.ld < op.cols) return set_error(XV_ERR_INVALID, "operand ld must be a multiple of 8 elements and >= cols");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(op.cols), static_cast<cuuint64_t>(op.rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(op.ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(op.mn_major ? 64 : box_rows_kmajor)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(XV_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  return XV_OK;
}

// Output tensor map for the epilogue's TMA stores: box = 32 rows x 128 bytes (64 bf16 or 32 f32 columns), 128B swizzle.
static int make_out_tmap(CUtensorMap* map, void* out, int M, int N, long long ldc, bool f32) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(XV_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  const int esz = f32 ? 4 : 2;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldc) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(XV_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (CUresult %d)", static_cast<int>(r));
  return XV_OK;
}

}  // namespace xv

// Persistent GEMM grids normally take every SM (one 227 KB CTA each).  While a collective runs concurrently (the
// overlapped gradient all-reduce of the data-parallel step) its CTAs hold some SMs; a GEMM CTA that cannot be placed
// starts late and, because tiles are strided statically over the grid, delays the whole GEMM.  The host therefore caps
// the grid at (SMs - reserve) for the kernels it enqueues during the overlap.  0 = no cap.
static int g_cta_limit = 0;
extern "C" int xv_gemm_set_cta_limit(int max_ctas) {
  if (max_ctas < 0) return ::xv::set_error(XV_ERR_INVALID, "xv_gemm_set_cta_limit: max_ctas must be >= 0");
  g_cta_limit = max_ctas;
  return XV_OK;
}

extern "C" int xv_gemm_bf16(const xv_gemm_args* a, void* stream) {
  using namespace xv;
  if (!a) return set_error(XV_ERR_INVALID, "null args");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(XV_ERR_INVALID, "M, N, K must be positive");
  if (a->splits < 1) return set_error(XV_ERR_INVALID, "splits must be >= 1");
  if (a->splits > 1 && a->epilogue != XV_EPI_F32) return set_error(XV_ERR_INVALID, "split-K needs the f32 epilogue");
  if (a->out == nullptr) return set_error(XV_ERR_INVALID, "null output");
  if (a->epilogue == XV_EPI_BF16 || a->epilogue == XV_EPI_HEAD_BWD) {
    if (a->ldc % 8) return set_error(XV_ERR_INVALID, "bf16 output ldc must be a multiple of 8");
  } else if (a->epilogue == XV_EPI_F32) {
    if (a->ldc % 4) return set_error(XV_ERR_INVALID, "f32 output ldc must be a multiple of 4");
  }
  for (const xv_operand* op : {&a->a, &a->b}) {
    if (!op->mn_major && op->div && (op->div % BLOCK_K)) return set_error(XV_ERR_INVALID, "K-major tap div must be a multiple of 64");
    if (op->mn_major && op->div && (op->div % 64)) return set_error(XV_ERR_INVALID, "MN-major tap div must be a multiple of 64");
  }
  if (a->a.mn_major && a->a.div && (a->a.div % BLOCK_M)) return set_error(XV_ERR_INVALID, "MN-major A tap div must be a multiple of 128");
  if (a->b.mn_major && a->b.div && (a->b.div % MAX_BN)) return set_error(XV_ERR_INVALID, "MN-major B tap div must be a multiple of 256");
  if ((a->col_sum == nullptr) != (a->col_sumsq == nullptr) && a->epilogue == XV_EPI_BF16)
    return set_error(XV_ERR_INVALID, "col_sum and col_sumsq must be given together");

  // CTA pairs (cta_group::2, 256 x 256 tiles) for the matrix-output epilogues when there are enough tiles to fill
  // the 74 SM pairs; head epilogues and small problems run one CTA per 128 x 128 tile.  XV_GEMM_CG=1|2 forces it.
  int cg = 1;
  {
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("XV_GEMM_CG");
      forced = e ? atoi(e) : 0;
    }
    const bool can = (a->epilogue == XV_EPI_BF16 || a->epilogue == XV_EPI_F32);
    const long long pair_tiles = static_cast<long long>((a->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) *
                                 ((a->N + MAX_BN - 1) / MAX_BN) * (a->splits > 0 ? a->splits : 1);
    // (M <= 128: one row block -- a pair tile would be half empty, the 128 x 128 single-CTA tiles fit exactly)
    if (can && ((forced == 2) || (forced == 0 && pair_tiles >= 64 && a->M > BLOCK_M))) cg = 2;
  }
  GemmKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  int rc = make_tmap(&kp.tma_a, a->a, BLOCK_M);
  if (rc) return rc;
  const int bn = cg == 2 ? GemmCfg<2>::BN : GemmCfg<1>::BN;
  rc = make_tmap(&kp.tma_b, a->b, bn / cg);
  if (rc) return rc;
  kp.M = a->M; kp.N = a->N; kp.K = a->K;
  kp.a_div = a->a.div; kp.a_tap = a->a.tap_rows; kp.b_div = a->b.div; kp.b_tap = a->b.tap_rows;
  kp.num_m = (a->M + BLOCK_M * cg - 1) / (BLOCK_M * cg);
  kp.num_n = (a->N + bn - 1) / bn;
  kp.num_kb = (a->K + BLOCK_K - 1) / BLOCK_K;
  int splits = a->splits < kp.num_kb ? a->splits : kp.num_kb;
  kp.kb_per_split = (kp.num_kb + splits - 1) / splits;
  splits = (kp.num_kb + kp.kb_per_split - 1) / kp.kb_per_split;   // no empty split
  kp.splits = splits;
  kp.seg_len = a->seg_len; kp.seg_valid = a->seg_valid;
  kp.accumulate = (a->epilogue == XV_EPI_BF16) ? a->accumulate : 0;
  kp.out = a->out; kp.ldc = a->ldc; kp.bias = a->bias; kp.col_sum = a->col_sum; kp.col_sumsq = a->col_sumsq;
  kp.head = a->head;
  memset(&kp.bnb, 0, sizeof(kp.bnb));
  if (a->bn_bwd.y != nullptr) {
    if (a->epilogue != XV_EPI_BF16 || !a->col_sum || !a->col_sumsq || (a->N % 32) || !a->bn_bwd.scale || !a->bn_bwd.shift ||
        !a->bn_bwd.mean || !a->bn_bwd.rstd || a->bn_bwd.ldy % 8 || a->bn_bwd.ldy < a->N ||
        (reinterpret_cast<uintptr_t>(a->bn_bwd.y) & 15))
      return set_error(XV_ERR_INVALID, "bn_bwd fusion needs the bf16 epilogue, col_sum/col_sumsq, N %% 32 == 0 and a 16-byte aligned y with ldy %% 8 == 0");
    kp.bnb = a->bn_bwd;
  }
  if (a->affine_scale != nullptr) {
    if (a->epilogue != XV_EPI_BF16 || !a->affine_shift || a->accumulate || a->col_sum || a->bn_bwd.y || (a->N % 4) ||
        (reinterpret_cast<uintptr_t>(a->affine_scale) & 15) || (reinterpret_cast<uintptr_t>(a->affine_shift) & 15))
      return set_error(XV_ERR_INVALID, "affine epilogue: bf16 output without statistics / accumulate, 16-byte aligned scale / shift, N %% 4 == 0");
    kp.aff_scale = a->affine_scale; kp.aff_shift = a->affine_shift; kp.aff_neg_slope = a->affine_neg_slope;
  }
  // Matrix outputs leave through TMA stores (split-K partials through the TMA reduce-add unit) whenever the layout
  // allows a tensor map; the gradient fan-in mode (read-modify-write of bf16) keeps the direct path.
  kp.use_tma_out = 0;
  if (a->epilogue != XV_EPI_HEAD_FWD && !(a->epilogue == XV_EPI_BF16 && a->accumulate)) {
    const bool f32 = a->epilogue == XV_EPI_F32;
    const bool aligned = (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (a->ldc * (f32 ? 4 : 2)) % 16 == 0 && a->ldc >= a->N;
    if (aligned) {
      rc = make_out_tmap(&kp.tma_out, a->out, a->M, a->N, a->ldc, f32);
      if (rc) return rc;
      kp.use_tma_out = 1;
    }
  }

  int sms = 0;
  rc = device_sm_count(&sms);
  if (rc) return rc;
  if ((kp.bnb.y != nullptr || kp.aff_scale != nullptr || (a->epilogue == XV_EPI_BF16 && kp.col_sum != nullptr)) && !kp.use_tma_out)
    return set_error(XV_ERR_INVALID, "column statistics / bn_bwd fusion need a TMA-storable bf16 output (16-byte aligned, no accumulate)");
  if (a->epilogue == XV_EPI_BF16 && kp.col_sum != nullptr && (a->N & 1))
    return set_error(XV_ERR_INVALID, "column statistics need an even N");
  const long long tiles = static_cast<long long>(kp.num_m) * kp.num_n * kp.splits;
  if (g_cta_limit > 0 && g_cta_limit < sms) sms = g_cta_limit < 2 ? 2 : g_cta_limit;   // leave SMs to a concurrent collective
  const int units = sms / cg;                                   // CTAs (cg = 1) or CTA pairs (cg = 2) on the device
  const int grid = static_cast<int>(tiles < units ? tiles : units) * cg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  kp.a_mn = a->a.mn_major != 0; kp.b_mn = a->b.mn_major != 0;
  if (cg == 2) {
    switch (a->epilogue) {
      case XV_EPI_BF16: return launch_gemm<XV_EPI_BF16, 2>(kp, grid, s);
      default: return launch_gemm<XV_EPI_F32, 2>(kp, grid, s);
    }
  }
  switch (a->epilogue) {
    case XV_EPI_BF16: return launch_gemm<XV_EPI_BF16, 1>(kp, grid, s);
    case XV_EPI_F32: return launch_gemm<XV_EPI_F32, 1>(kp, grid, s);
    case XV_EPI_HEAD_FWD: return launch_gemm<XV_EPI_HEAD_FWD, 1>(kp, grid, s);
    case XV_EPI_HEAD_BWD: return launch_gemm<XV_EPI_HEAD_BWD, 1>(kp, grid, s);
    default: return set_error(XV_ERR_INVALID, "unknown epilogue %d", a->epilogue);
  }
}
