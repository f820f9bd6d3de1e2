// peer.cu -- peer-memory plumbing for the fused Ulysses exchange: buffers that other ranks of the same NVSwitch box map
// into their address space (CUDA IPC), so that kernel 0 can READ its heads' rows straight out of the peers' projection
// outputs and kernel 4 can WRITE its output rows straight into the peers' result buffers over NVLink -- no NCCL data
// collective, no send/receive staging copies.  cudaMalloc memory (not a PyTorch caching-allocator sub-block) because an
// IPC handle names a whole allocation.
#include <string.h>

#include "rsa_common.cuh"

using namespace rsa;

extern "C" int rsa_peer_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) RSA_FAIL(RSA_ERR_ARG, "rsa_peer_alloc: bad arguments");
  RSA_CUDA_CHECK(cudaMalloc(ptr, bytes));
  RSA_CUDA_CHECK(cudaMemset(*ptr, 0, bytes));
  return RSA_OK;
}

extern "C" int rsa_peer_free(void* ptr) {
  if (ptr) RSA_CUDA_CHECK(cudaFree(ptr));
  return RSA_OK;
}

extern "C" int rsa_peer_export(const void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == RSA_PEER_HANDLE_BYTES, "IPC handle size");
  if (!ptr || !handle64) RSA_FAIL(RSA_ERR_ARG, "rsa_peer_export: null pointer");
  cudaIpcMemHandle_t h;
  RSA_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, sizeof(h));
  return RSA_OK;
}

extern "C" int rsa_peer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) RSA_FAIL(RSA_ERR_ARG, "rsa_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  RSA_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RSA_OK;
}

extern "C" int rsa_peer_close(void* ptr) {
  if (ptr) RSA_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return RSA_OK;
}
