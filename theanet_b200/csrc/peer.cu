// Peer-memory plumbing for the fused gradient all-reduce + update (update.cu): buffers that other
// ranks (one process per GPU) map over NVLink with CUDA IPC.  Not in the reference (it has no
// multi-device path); SURVEY.md 8(f1).
#include "common.cuh"

using namespace tn;

// cudaMalloc'ed (never from a caching allocator: the IPC handle must describe exactly this block),
// zero-filled
extern "C" int tn_peer_alloc(size_t bytes, void **ptr) {
  TN_REQUIRE(ptr && bytes > 0, TN_ERR_ARG, "tn_peer_alloc: bad argument");
  cudaError_t e = cudaMalloc(ptr, bytes);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_peer_alloc: %s", cudaGetErrorString(e));
  e = cudaMemset(*ptr, 0, bytes);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_peer_alloc: %s", cudaGetErrorString(e));
  return TN_OK;
}

extern "C" int tn_peer_free(void *ptr) {
  cudaError_t e = cudaFree(ptr);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_peer_free: %s", cudaGetErrorString(e));
  return TN_OK;
}

// handle_out: 64 bytes (cudaIpcMemHandle_t)
extern "C" int tn_ipc_get_handle(void *ptr, void *handle_out) {
  TN_REQUIRE(ptr && handle_out, TN_ERR_ARG, "tn_ipc_get_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out, ptr);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_ipc_get_handle: %s", cudaGetErrorString(e));
  return TN_OK;
}

extern "C" int tn_ipc_open_handle(const void *handle, void **ptr) {
  TN_REQUIRE(handle && ptr, TN_ERR_ARG, "tn_ipc_open_handle: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_ipc_open_handle: %s", cudaGetErrorString(e));
  return TN_OK;
}

extern "C" int tn_ipc_close_handle(void *ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_ipc_close_handle: %s", cudaGetErrorString(e));
  return TN_OK;
}
