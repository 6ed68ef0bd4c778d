// Instantiation of the tcgen05 GEMM kernel for epilogue 2, one CTA per tile (see xv_gemm_kernel.cuh).
#include "xv_gemm_kernel.cuh"
namespace xv {
template int launch_gemm<2, 1>(const GemmKernelParams&, int, cudaStream_t);
}
