// Exchange buffers of the distributed Sinkhorn: one cudaMalloc'ed buffer per rank, exported / mapped with CUDA IPC
// (one process per GPU on one box; the peers are reached over NVLink / NVSwitch by plain stores).
#include <string.h>

#include "gx_common.cuh"

extern "C" int gx_peer_alloc(long long bytes, void** ptr) {
  GX_CHECK_ARG(ptr && bytes > 0);
  void* p = nullptr;
  GX_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  GX_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  GX_CHECK_CUDA(cudaDeviceSynchronize());
  *ptr = p;
  return GX_OK;
}

extern "C" int gx_peer_free(void* ptr) {
  GX_CHECK_ARG(ptr);
  GX_CHECK_CUDA(cudaFree(ptr));
  return GX_OK;
}

extern "C" int gx_peer_export(void* ptr, void* handle64) {
  GX_CHECK_ARG(ptr && handle64);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  GX_CHECK_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return GX_OK;
}

extern "C" int gx_peer_open(const void* handle64, void** ptr) {
  GX_CHECK_ARG(ptr && handle64);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  GX_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return GX_OK;
}

extern "C" int gx_peer_close(void* ptr) {
  GX_CHECK_ARG(ptr);
  GX_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return GX_OK;
}
