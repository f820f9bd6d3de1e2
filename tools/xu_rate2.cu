// xu_rate2.cu -- does the MUFU pipe evaluate a PACKED pair of 16-bit exponentials per lane-slot?  Kernel 4's softmax
// needs 128 x 128 exponentials per kept pair = 1024 MUFU cycles per tile at 4 fp32 results per clock and sub-partition,
// exactly the tensor pipe's 1024 cycles; if ex2.approx.ftz.f16x2 / .bf16x2 retire two results per slot, P (which is
// rounded to 16 bits anyway) could be produced at twice the rate.  One warp per sub-partition (and two), dependent-free
// streams of (a) ex2.f32 x2, (b) cvt.f16x2 + ex2.f16x2, (c) cvt.bf16x2 + ex2.bf16x2, (d) ex2.f16x2 alone, (e) ex2.bf16x2
// alone; cycles per PAIR of exponentials per warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate2 xu_rate2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) { uint32_t d; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
__device__ __forceinline__ uint32_t cvt_b2(float lo, float hi) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = 0xb800b800u + threadIdx.x + i;  // two negative halves
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float a = x[i], b = x[i + 1];
      if (MODE == 0) {
        a = ex2(a); b = ex2(b);
        acc ^= __float_as_uint(a) ^ __float_as_uint(b);
      } else if (MODE == 1) {
        acc ^= ex2_h2(cvt_h2(a, b));
      } else if (MODE == 2) {
        acc ^= ex2_b2(cvt_b2(a, b));
      } else if (MODE == 3) {
        u[i / 2] = ex2_h2(u[i / 2]) | 0x80008000u;
      } else if (MODE == 4) {
        u[i / 2] = ex2_b2(u[i / 2]) | 0x80008000u;
      }
      if (MODE <= 2) { x[i] = a * 0.5f - 1.0f; x[i + 1] = b * 0.5f - 1.0f; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc) + x[3];
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[5] = {"ex2.f32 x2", "cvt.f16x2 + ex2.f16x2", "cvt.bf16x2 + ex2.bf16x2", "ex2.f16x2 alone", "ex2.bf16x2 alone"};
  for (int threads : {128, 256}) {
    for (int mode = 0; mode < 5; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
        if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
        if (mode == 4) k<4><<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%d warps/SMSP  %-26s %.2f cycles per pair of exponentials per warp\n", threads / 128, names[mode], (double)c / iters / 8);
    }
  }
  return 0;
}
