#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cp {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace cp

extern "C" const char* cp_last_error_string(void) { return cp::g_err; }
extern "C" int cp_version(void) { return 100; }
extern "C" int cp_device_arch(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cp::set_error("cp_device_arch: no CUDA device");
    (void)cudaGetLastError();
    return CP_E_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}
