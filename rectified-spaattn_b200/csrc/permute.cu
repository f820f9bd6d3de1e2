// permute.cu -- kernel 1: Gilbert-curve token permute / unpermute as a coalesced 16-byte row gather.
// Replaces hidden_states[:, hilbert_order] / hidden_states[:, linear_to_hilbert]
// (reference scripts/main_hunyuan.py:88-89, :183; ATen int64 advanced-index gather).
//
// HBM-bound: algorithmic bytes = 2 * batch * n_out * row_bytes.  Every thread moves kUnroll independent 16-B
// chunks (all loads issued before the first store); consecutive threads touch consecutive chunks of the same
// row, so both sides are fully coalesced; the int64 index is read once per chunk through L1.
#include "rsa_common.cuh"

namespace rsa {

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w));
}

__global__ void __launch_bounds__(kThreads) permute_rows_kernel(const char* __restrict__ src, char* __restrict__ dst,
                                                                const int64_t* __restrict__ index, int64_t n_out,
                                                                int64_t n_src, int chunks_per_row,
                                                                int64_t src_batch_stride, int64_t dst_batch_stride,
                                                                int64_t chunks_per_batch) {
  const int b = blockIdx.y;
  const int64_t base = ((int64_t)blockIdx.x * kUnroll) * kThreads + threadIdx.x;
  const char* sb = src + (int64_t)b * src_batch_stride;
  char* db = dst + (int64_t)b * dst_batch_stride;
  int4 val[kUnroll];
  int64_t doff[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t c = base + (int64_t)u * kThreads;
    doff[u] = -1;
    if (c < chunks_per_batch) {
      const int64_t row = c / chunks_per_row;
      const int col = (int)(c - row * chunks_per_row);
      const int64_t srow = __ldg(index + row);
      if (srow >= 0 && srow < n_src) {
        val[u] = ld_stream(reinterpret_cast<const int4*>(sb + (srow * chunks_per_row + col) * 16));
      } else {
        val[u] = make_int4(0, 0, 0, 0);  // out-of-range index: defined (zero) instead of a fault
      }
      doff[u] = c * 16;
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
    if (doff[u] >= 0) st_stream(reinterpret_cast<int4*>(db + doff[u]), val[u]);
}

}  // namespace
}  // namespace rsa

extern "C" int rsa_permute_rows(const void* src, void* dst, const int64_t* index, int batch, int64_t n_out,
                                int64_t n_src, int64_t row_bytes, int64_t src_batch_stride_bytes,
                                int64_t dst_batch_stride_bytes, void* stream) {
  using namespace rsa;
  if (!src || !dst || !index) RSA_FAIL(RSA_ERR_ARG, "rsa_permute_rows: null pointer");
  if (batch < 0 || n_out < 0 || n_src < 0 || row_bytes < 0) RSA_FAIL(RSA_ERR_ARG, "rsa_permute_rows: negative size");
  if (batch == 0 || n_out == 0 || row_bytes == 0) return RSA_OK;
  if (row_bytes % 16 || ((uintptr_t)src % 16) || ((uintptr_t)dst % 16) || src_batch_stride_bytes % 16 ||
      dst_batch_stride_bytes % 16)
    RSA_FAIL(RSA_ERR_UNSUPPORTED, "rsa_permute_rows: rows, pointers and batch strides must be 16-byte multiples");
  if (src == dst) RSA_FAIL(RSA_ERR_ARG, "rsa_permute_rows: in-place permutation is not supported");
  if (batch > 65535) RSA_FAIL(RSA_ERR_UNSUPPORTED, "rsa_permute_rows: batch > 65535");
  const int chunks_per_row = (int)(row_bytes / 16);
  const int64_t chunks = n_out * chunks_per_row;
  const int64_t per_block = (int64_t)kThreads * kUnroll;
  const int64_t blocks = (chunks + per_block - 1) / per_block;
  if (blocks > 0x7fffffffLL) RSA_FAIL(RSA_ERR_UNSUPPORTED, "rsa_permute_rows: too many rows");
  dim3 grid((unsigned)blocks, (unsigned)batch);
  permute_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      (const char*)src, (char*)dst, index, n_out, n_src, chunks_per_row, src_batch_stride_bytes,
      dst_batch_stride_bytes, chunks);
  RSA_CUDA_CHECK(cudaGetLastError());
  return RSA_OK;
}
