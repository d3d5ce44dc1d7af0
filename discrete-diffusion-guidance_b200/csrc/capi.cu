// Error plumbing and device checks shared by every C-ABI entry point.
#include <cstdarg>
#include <cstdio>

#include "ddg_common.cuh"

namespace ddg {
namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return DDG_OK;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return DDG_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace ddg

extern "C" int ddg_abi_version(void) { return DDG_ABI_VERSION; }

extern "C" const char* ddg_last_error(void) { return ddg::g_err; }

extern "C" int ddg_sm_count(void) { return ddg::sm_count(); }

extern "C" int ddg_device_check(void) {
  // Idempotent attribute cache; the only global mutable state in the library.
  static int ok[64] = {0};  // 0 unknown, 1 good, -1 bad
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    ddg::set_error("no CUDA device: %s (ddg_b200 has no CPU fallback)", cudaGetErrorString(e));
    return DDG_ERR_UNSUPPORTED;
  }
  if (dev < 0 || dev >= 64) return DDG_OK;
  if (ok[dev] == 0) {
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    ok[dev] = (major == 10) ? 1 : -1;
    if (ok[dev] < 0)
      ddg::set_error("device %d is sm_%d%d; ddg_b200 is built for sm_100a only", dev, major, minor);
  }
  if (ok[dev] < 0) {
    return DDG_ERR_UNSUPPORTED;
  }
  return DDG_OK;
}
