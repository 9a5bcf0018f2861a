// Library-level entry points of the C ABI: versioning, thread-local error string, device info.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "../../include/dasp_b200.h"

namespace dasp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached_dev = dev;
    cached_sms = n;
  }
  return cached_sms;
}

static int g_forced_warps = 0;
int debug_forced_warps() { return g_forced_warps; }
static int g_reverb_path = 0;
int debug_reverb_path() { return g_reverb_path; }

void reverb_shutdown();  // reverb.cu

}  // namespace dasp

extern "C" {

int dasp_abi_version(void) { return DASP_ABI_VERSION; }
const char* dasp_last_error(void) { return dasp::g_err; }
int dasp_compiled_arch(void) { return 1000; }
void dasp_shutdown(void) { dasp::reverb_shutdown(); }
void dasp_debug_force_warps(int warps) { dasp::g_forced_warps = (warps == 1 || warps == 2 || warps == 4 || warps == 8) ? warps : 0; }

void dasp_debug_reverb_path(int path) { dasp::g_reverb_path = (path == 1 || path == 2) ? path : 0; }

}  // extern "C"
