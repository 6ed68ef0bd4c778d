// Instantiation of the tcgen05 GEMM kernel for epilogue 3 (see xv_gemm_kernel.cuh).
#include "xv_gemm_kernel.cuh"
namespace xv {
template int launch_gemm<3>(const GemmKernelParams&, int, cudaStream_t);
}
