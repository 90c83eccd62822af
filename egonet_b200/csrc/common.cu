// Error reporting + device gate shared by every entry point.
#include "common.h"

#include <cstdlib>
#include <cstring>

namespace egn {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* get_error() { return g_err; }

bool pdl_enabled() {
  static const bool on = !(getenv("EGN_PDL") && atoi(getenv("EGN_PDL")) == 0);
  return on;
}

static int probe_device(std::string* why) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    *why = std::string("no CUDA device: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return 0;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    *why = "cudaGetDevice failed";
    return 0;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    char buf[128];
    snprintf(buf, sizeof(buf), "device %d is sm_%d%d; this library is built for sm_100a only", dev,
             major, minor);
    *why = buf;
    return 0;
  }
  return 1;
}

int require_device() {
  std::string why;
  if (!probe_device(&why)) {
    set_error("%s (egonet_b200 has no CPU fallback)", why.c_str());
    return EGN_ERR_NO_DEVICE;
  }
  return EGN_OK;
}

}  // namespace egn

extern "C" {

int egn_version(void) { return EGN_VERSION; }
const char* egn_last_error(void) { return egn::get_error(); }
int egn_device_ok(void) {
  std::string why;
  return egn::probe_device(&why);
}

}  // extern "C"
