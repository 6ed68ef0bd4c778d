// Instantiation of the tcgen05 GEMM kernel for epilogue 2 (see xv_gemm_kernel.cuh).
#include "xv_gemm_kernel.cuh"
namespace xv {
template int launch_gemm<2>(const GemmKernelParams&, int, cudaStream_t);
}
