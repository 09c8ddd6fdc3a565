// Library-level entry points of the srgpt C-ABI: version, last error, device info.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "srgpt_b200.h"

namespace srgpt {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200
  }
  return cached;
}

bool env_flag(const char* name) {
  const char* v = getenv(name);
  return v != nullptr && v[0] != 0 && v[0] != '0';
}
bool pdl_enabled() {
  static const bool on = !env_flag("SRGPT_NO_PDL");
  return on;
}

static unsigned long long* g_trace_buf = nullptr;
static int g_trace_cap = 0, g_trace_n = 0;

unsigned long long* trace_next_slot() {
  if (g_trace_buf == nullptr || g_trace_n >= g_trace_cap) return nullptr;
  return g_trace_buf + 4 * (size_t)(g_trace_n++);
}

}  // namespace srgpt

using namespace srgpt;

extern "C" __attribute__((visibility("default"))) int srgpt_trace_begin(void* device_buf, int capacity_records) {
  SRGPT_CHECK_ARG(device_buf != nullptr && capacity_records > 0);
  g_trace_buf = reinterpret_cast<unsigned long long*>(device_buf);
  g_trace_cap = capacity_records;
  g_trace_n = 0;
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) int srgpt_trace_end(void) {
  const int n = g_trace_n;
  g_trace_buf = nullptr;
  g_trace_cap = g_trace_n = 0;
  return n;
}

extern "C" __attribute__((visibility("default"))) int srgpt_abi_version(void) { return SRGPT_ABI_VERSION; }

extern "C" __attribute__((visibility("default"))) int srgpt_elem_type(void) {
#ifdef SRGPT_ELEM_F16
  return 1;
#else
  return 0;
#endif
}

extern "C" __attribute__((visibility("default"))) const char* srgpt_last_error(void) { return g_last_error; }

extern "C" __attribute__((visibility("default"))) int srgpt_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0, n = 0, maj = 0, mnr = 0;
  SRGPT_CHECK_CUDA(cudaGetDevice(&dev));
  SRGPT_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  SRGPT_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  SRGPT_CHECK_CUDA(cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_out) *sm_count_out = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = mnr;
  if (maj != 10) {
    set_last_error("srgpt_b200 kernels are built for sm_100a only; device is sm_%d%d", maj, mnr);
    return SRGPT_ERR_UNSUPPORTED;
  }
  return SRGPT_OK;
}
