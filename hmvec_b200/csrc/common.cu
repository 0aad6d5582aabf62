// common.cu -- error plumbing and device probes for the C ABI.
#include "common.cuh"

namespace hmv {
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace hmv

extern "C" int hmv_abi_version(void) { return 1; }
extern "C" const char* hmv_last_error(void) { return hmv::err_buf(); }
extern "C" int hmv_device_cc(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return hmv::fail(HMV_E_CUDA, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
  return p.major * 10 + p.minor;
}
