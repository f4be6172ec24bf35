// Library-wide entry points: version, error text, device info.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace edet {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace edet

extern "C" int edet_version(void) { return 100; }
extern "C" const char* edet_last_error(void) { return edet::g_err; }
extern "C" int edet_device_info(int* sm_count, int* cc) {
  int dev = 0, sms = 0, major = 0, minor = 0;
  EDET_CHECK_CUDA(cudaGetDevice(&dev));
  EDET_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  EDET_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  EDET_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc) *cc = major * 10 + minor;
  return EDET_OK;
}
