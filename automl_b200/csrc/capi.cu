// Library-wide entry points: version, error text, device info.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "tc_common.cuh"

namespace edet {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_opt_dw_impl{0};
int option_dw_impl() { return g_opt_dw_impl.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_stem_impl{0};
int option_stem_impl() { return g_opt_stem_impl.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_sepconv_impl{0};
int option_sepconv_impl() { return g_opt_sepconv_impl.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_pw_teams{0};
int option_pw_teams() { return g_opt_pw_teams.load(std::memory_order_relaxed); }

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    set_error("cudaGetDevice failed or device ordinal %d >= %d", dev, kMaxDevices);
    return -1;
  }
  return dev;
}

int device_sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  const int dev = current_device();
  if (dev < 0) return 0;
  int v = cached[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
      set_error("cudaDeviceGetAttribute(MultiProcessorCount) failed");
      return 0;
    }
    cached[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

namespace pwtc {
// zero-initialised in every context that loads the module; each slot resets itself
__device__ unsigned g_tile_sched[2 * kSchedSlots];

unsigned* next_sched_slot() {
  static std::atomic<unsigned*> base[kMaxDevices];
  static std::atomic<unsigned> next[kMaxDevices];
  const int dev = current_device();
  if (dev < 0) return nullptr;
  unsigned* b = base[dev].load(std::memory_order_acquire);
  if (b == nullptr) {
    void* addr = nullptr;
    if (cudaGetSymbolAddress(&addr, g_tile_sched) != cudaSuccess || addr == nullptr) {
      set_error("cudaGetSymbolAddress(g_tile_sched) failed");
      return nullptr;
    }
    b = static_cast<unsigned*>(addr);
    base[dev].store(b, std::memory_order_release);
  }
  const unsigned i = next[dev].fetch_add(1u, std::memory_order_relaxed) % kSchedSlots;
  return b + 2 * i;
}
}  // namespace pwtc
}  // namespace edet

extern "C" int edet_version(void) { return 200; }
extern "C" int edet_set_option(const char* name, int value) {
  using namespace edet;
  EDET_CHECK_ARG(name != nullptr, "set_option: null name");
  if (strcmp(name, "dw_impl") == 0) {
    EDET_CHECK_ARG(value >= 0 && value <= 2, "set_option: dw_impl must be 0, 1 or 2");
    g_opt_dw_impl.store(value);
    return EDET_OK;
  }
  if (strcmp(name, "sepconv_impl") == 0) {
    EDET_CHECK_ARG(value == 0 || value == 1, "set_option: sepconv_impl must be 0 or 1");
    g_opt_sepconv_impl.store(value);
    return EDET_OK;
  }
  if (strcmp(name, "stem_impl") == 0) {
    EDET_CHECK_ARG(value == 0 || value == 1, "set_option: stem_impl must be 0 or 1");
    g_opt_stem_impl.store(value);
    return EDET_OK;
  }
  if (strcmp(name, "pw_teams") == 0) {
    EDET_CHECK_ARG(value == 0 || value == 2 || value == 3, "set_option: pw_teams must be 0, 2 or 3");
    g_opt_pw_teams.store(value);
    return EDET_OK;
  }
  set_error("set_option: unknown option '%s'", name);
  return EDET_ERR_INVALID;
}
extern "C" int edet_get_option(const char* name, int* value) {
  using namespace edet;
  EDET_CHECK_ARG(name != nullptr && value != nullptr, "get_option: null pointer");
  if (strcmp(name, "dw_impl") == 0) {
    *value = option_dw_impl();
    return EDET_OK;
  }
  if (strcmp(name, "sepconv_impl") == 0) {
    *value = option_sepconv_impl();
    return EDET_OK;
  }
  if (strcmp(name, "stem_impl") == 0) {
    *value = option_stem_impl();
    return EDET_OK;
  }
  if (strcmp(name, "pw_teams") == 0) {
    *value = option_pw_teams();
    return EDET_OK;
  }
  set_error("get_option: unknown option '%s'", name);
  return EDET_ERR_INVALID;
}
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
