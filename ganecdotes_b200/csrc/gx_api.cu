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

// SM budgets of the two kernel families (0 = all SMs).  They let a caller run tensor-bound
// contractions and HBM-bound streaming kernels side by side on disjoint SM sets.
static int g_umma_ctas = 0, g_stream_ctas = 0;
int gx_umma_cta_budget() { return g_umma_ctas > 0 ? g_umma_ctas : gx_sm_count(); }
int gx_stream_cta_budget() { return g_stream_ctas > 0 ? g_stream_ctas : gx_sm_count(); }

extern "C" int gx_set_sm_budget(int umma_ctas, int stream_ctas) {
  if (umma_ctas < 0 || stream_ctas < 0) return GX_ERR_ARG;
  g_umma_ctas = umma_ctas;
  g_stream_ctas = stream_ctas;
  return GX_OK;
}

// bumped whenever a descriptor struct or an exported signature changes
extern "C" int gx_version(void) { return GX_ABI_VERSION; }

// sizeof() of descriptor `which` (0 conv, 1 gemm, 2 gather, 3 ll): a binding compares it with its own layout at
// load time, so a stale library is an error and not a silent misread of the struct fields
extern "C" int gx_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(gx_conv_desc);
    case 1: return (int)sizeof(gx_gemm_desc);
    case 2: return (int)sizeof(gx_gather_desc);
    case 3: return (int)sizeof(gx_ll_desc);
    default: return -1;
  }
}

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
