// Instantiation of the tcgen05 GEMM kernel for epilogue 0, one CTA per tile (see xv_gemm_kernel.cuh).
#include "xv_gemm_kernel.cuh"
namespace xv {
template int launch_gemm<0, 1>(const GemmKernelParams&, int, cudaStream_t);
}
