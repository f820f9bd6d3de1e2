// xu_rate.cu -- which pipe do the softmax's non-FMA instructions share?  One warp per SM sub-partition (128 threads,
// one CTA per SM) runs a dependent-free stream of (a) ex2.approx only, (b) cvt.rn.bf16x2.f32 only, (c) both in the
// 2:1 mix of kernel 4's softmax, (d) ex2 + a PRMT-based truncating pack instead of the cvt; cycles per warp-level
// instruction are printed.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate xu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
__device__ __forceinline__ uint32_t prmt_hi(float lo, float hi) { uint32_t d; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(d) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi))); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float a = x[i], b = x[i + 1];
      if (MODE == 0 || MODE == 2 || MODE == 3) { a = ex2(a); b = ex2(b); }
      if (MODE == 1 || MODE == 2) acc ^= cvt2(a, b);
      if (MODE == 3) acc ^= prmt_hi(a * 1.001953125f, b * 1.001953125f);
      if (MODE == 0) acc ^= __float_as_uint(a) ^ __float_as_uint(b);
      x[i] = a * 0.5f - 1.0f; x[i + 1] = b * 0.5f - 1.0f;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc) + x[3];
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[4] = {"ex2 x2", "cvt.bf16x2 x1", "ex2 x2 + cvt x1", "ex2 x2 + fmul x2 + prmt x1"};
  for (int threads : {128, 256}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
        if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%d warps/SMSP  %-28s %.2f cycles per group of (2 elements) per warp\n", threads / 128, names[mode], (double)c / iters / 8);
    }
  }
  return 0;
}
