// Internal helpers shared by the translation units of libxvector_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/xvector_b200.h"

namespace xv {

int set_error(int code, const char* fmt, ...);
int device_sm_count(int* out);

#define XV_CUDA_CHECK(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::xv::set_error(XV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                  \
  } while (0)

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// Every kernel of the library is launched with programmaticStreamSerializationAllowed and begins with pdl_entry():
// "launch_dependents" lets the NEXT kernel of the stream be scheduled while this one is still running, "wait" blocks
// until the PREVIOUS kernel has completed and flushed its memory.  Memory semantics are those of plain stream order
// (nothing touches global memory before the wait); what is gained is the launch latency, CTA scheduling and -- in the
// GEMM -- barrier / TMEM / tensor-map setup of kernel N+1 overlapping the tail of kernel N.  The edges survive CUDA
// graph capture.  Opt-in with XV_PDL=1: inside the captured training step the kernels already run back to back, and
// the A/B measurement on B200 showed no gain (1.137 ms with PDL vs 1.131 ms without).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {     // thread-block cluster (CTA pair of the cta_group::2 GEMM): both CTAs land on one TPC
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  return launch_cluster(kernel, grid, block, smem, stream, 1, static_cast<Args&&>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() {
  pdl_launch_dependents();
  pdl_wait();
}
#endif

}  // namespace xv
