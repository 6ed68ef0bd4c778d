// Error reporting and device queries of the C ABI.
#include <stdlib.h>

#include "xv_internal.h"

namespace xv {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int device_sm_count(int* out) {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    XV_CUDA_CHECK(cudaGetDevice(&dev));
    int major = 0;
    XV_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return set_error(XV_ERR_CUDA, "xvector_b200 needs an sm_100 GPU (found compute capability %d.x); there is no fallback", major);
    XV_CUDA_CHECK(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
  }
  *out = cached;
  return XV_OK;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("XV_PDL");
    cached = (e && e[0] == '1') ? 1 : 0;     // measured on B200: no gain inside the captured step (1.137 vs 1.131 ms), so opt-in
  }
  return cached != 0;
}

}  // namespace xv

extern "C" const char* xv_last_error(void) { return xv::g_err; }
extern "C" int xv_version(void) { return 100; }
extern "C" int xv_device_info(int32_t out[3]) {
  int dev = 0;
  XV_CUDA_CHECK(cudaGetDevice(&dev));
  int v = 0;
  XV_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  out[0] = v;
  XV_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
  out[1] = v;
  XV_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
  out[2] = v;
  return XV_OK;
}
