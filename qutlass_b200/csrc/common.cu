// Error plumbing + device queries for libb200q.
#include "common.cuh"

#include <string.h>

namespace b200q {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_sm_major[64];
static int g_sm_count[64];
static bool g_dev_init[64];

static int query_device(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return B200Q_ECUDA;
  }
  if (dev < 0 || dev >= 64) {
    set_error("device ordinal %d out of range", dev);
    return B200Q_EINVAL;
  }
  if (!g_dev_init[dev]) {
    int major = 0, minor = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_sm_major[dev] = major * 10 + minor;
    g_sm_count[dev] = sms;
    g_dev_init[dev] = true;
  }
  *dev_out = dev;
  return 0;
}

int check_device_sm100() {
  int dev;
  int rc = query_device(&dev);
  if (rc) return rc;
  if (g_sm_major[dev] != 100) {
    set_error("libb200q is built for sm_100a only; current device is sm_%d", g_sm_major[dev]);
    return B200Q_EUNSUPPORTED;
  }
  return 0;
}

int current_device() {
  int dev;
  return query_device(&dev) ? 0 : dev;
}

int num_sms() {
  int dev;
  if (query_device(&dev)) return 148;
  return g_sm_count[dev] > 0 ? g_sm_count[dev] : 148;
}

}  // namespace b200q

extern "C" int b200q_abi_version(void) { return 1; }
extern "C" const char* b200q_last_error(void) { return b200q::g_err; }
