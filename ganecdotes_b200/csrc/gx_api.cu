// Library-level entry points: version, error reporting, device probing.
#include "gx_common.cuh"

static thread_local int g_last_cuda_error = 0;

void gx_set_last_cuda_error(int e) { g_last_cuda_error = e; }

int gx_sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return GX_SM_COUNT_FALLBACK;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return GX_SM_COUNT_FALLBACK;
  cached = n;
  return n;
}

extern "C" int gx_version(void) { return 100; }

extern "C" int gx_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" const char* gx_error_string(int code) {
  switch (code) {
    case GX_OK: return "ok";
    case GX_ERR_ARG: return "invalid argument or unsupported shape";
    case GX_ERR_CUDA: return "CUDA error (see gx_last_cuda_error)";
    case GX_ERR_UNSUPPORTED: return "device is not sm_100";
    default: return "unknown error";
  }
}

extern "C" int gx_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
