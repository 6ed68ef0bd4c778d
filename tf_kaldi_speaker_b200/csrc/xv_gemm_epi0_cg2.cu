// Instantiation of the tcgen05 GEMM kernel for epilogue 0, CTA pairs (cta_group::2, 256 x 256 tiles).
#include "xv_gemm_kernel.cuh"
namespace xv {
template int launch_gemm<0, 2>(const GemmKernelParams&, int, cudaStream_t);
}
