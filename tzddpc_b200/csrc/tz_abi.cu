// Version / error plumbing of the C ABI (include/tzddpc.h).
#include "tz_common.cuh"

namespace tz {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace tz

extern "C" const char* tz_version(void) { return "tzddpc-b200 0.1.0 (sm_100a, fp64)"; }

extern "C" size_t tz_last_error(char* buf, size_t cap) {
  const char* m = tz::last_error_buf();
  const size_t n = strlen(m);
  if (buf && cap) {
    const size_t k = n < cap - 1 ? n : cap - 1;
    memcpy(buf, m, k);
    buf[k] = 0;
  }
  return n;
}

extern "C" int tz_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  TZ_CUDA(cudaGetDevice(&dev));
  TZ_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TZ_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}
