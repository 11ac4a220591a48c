// Error plumbing shared by every C-ABI entry point (include/dmc_b200.h).
// Entry points return 0 on success and a negative code otherwise; the message
// is retrievable with dmc_last_error().  Nothing here synchronises the device.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <atomic>

static thread_local char g_err[512] = "";

void dmc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};

int dmc_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    dmc_set_error("%s: %s", what, cudaGetErrorString(e));
    return DMC_ERR_CUDA;
  }
  return DMC_OK;
}

extern "C" const char* dmc_last_error(void) { return g_err; }

extern "C" int dmc_abi_version(void) { return 1; }

// Compute capability of the current device as major*10+minor (100 on B200), or
// a negative code when no device is usable.  The Python host refuses to run the
// product path on anything else.
extern "C" int dmc_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    dmc_set_error("no CUDA device");
    return DMC_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

// Number of kernel launches issued by this library since the last reset (every
// launch site goes through dmc_check_launch exactly once).
extern "C" long long dmc_launch_count(void) { return g_launches.load(); }
extern "C" void dmc_reset_launch_count(void) { g_launches.store(0); }

// 0: shared zero ring (row 0 / column 0 only), 1: private ring on every side (see common.cuh).
extern "C" int dmc_layout_pad_hi(void) { return DMC_PAD_HI; }
