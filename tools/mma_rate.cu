// Micro-benchmark (bring-up tool, not product): issue rate of tcgen05.mma kind::f16 shapes, one CTA per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rectified-spaattn_b200/csrc -I include tools/mma_rate.cu -o tools/_build/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace rsa::ptx;

constexpr int kData = 12 * 16384;  // 4 A granules + 8 B granules

// MODE 0: SS (A, B K-major in smem)   1: TS (A in TMEM, B K-major)   2: TS (A in TMEM, B MN-major, like P V)
// DTILES: number of distinct accumulator tiles rotated over (1 = every MMA accumulates into the same D)
template <int MODE, int N, int DTILES>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* tmem_slot = (uint32_t*)(smem + kData);
  const uint32_t barw = smem_u32(smem + kData + 8);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kData / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  __syncthreads();
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(barw, 1); fence_barrier_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sb = smem_u32(smem);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, MODE == 2);
    uint64_t ad[8], bd[8];
    for (int j = 0; j < 8; ++j) {
      ad[j] = smem_desc_sw128(sb + (j >> 2) * 16384) + 2 * (j & 3);  // like Q: two head_dim halves x 4 k-steps
      bd[j] = MODE == 2 ? smem_desc_sw128(sb + 65536, 16384) + 128 * j
                        : smem_desc_sw128(sb + 65536 + (j >> 2) * 16384, 16384) + 2 * (j & 3);
    }
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + 256 + (DTILES > 1 ? (it % DTILES) * (N < 128 ? N : 128) : 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) umma_ss(N == 256 ? tmem : d, ad[j], bd[j], idesc, 1);
          else umma_ts(d, tmem + j * 8, bd[j], idesc, 1);
        }
      }
      umma_commit(barw);
      long long t1 = clock64();
      mbar_wait(barw, rep & 1);
      long long t2 = clock64();
      if (t2 - t0 < best) { best = t2 - t0; out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE, int N, int DTILES>
void run(int grid, long long* out) {
  const char* names[3] = {"SS", "TS(B K-major)", "TS(B MN-major)"};
  const int iters = 32, smem = kData + 64;
  cudaFuncSetAttribute(rate_kernel<MODE, N, DTILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<MODE, N, DTILES><<<grid, 128, smem>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  double issue = 0, total = 0;
  for (int b = 0; b < grid; ++b) { issue += out[2 * b]; total += out[2 * b + 1]; }
  printf("grid %3d %-15s M128 N%-3d K16 Dtiles %d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d)\n", grid,
         names[MODE], N, DTILES, issue / grid / (iters * 8), total / grid / (iters * 8), N / 2);
}

int main() {
  long long* out; cudaMallocManaged(&out, 148 * 2 * sizeof(long long));
  for (int grid : {1, 148}) {
    run<0, 64, 1>(grid, out);  run<0, 128, 1>(grid, out); run<0, 128, 2>(grid, out); run<0, 256, 1>(grid, out);
    run<1, 64, 1>(grid, out);  run<1, 128, 1>(grid, out); run<1, 128, 2>(grid, out);
    run<2, 64, 1>(grid, out);  run<2, 128, 1>(grid, out); run<2, 128, 2>(grid, out);
  }
  return 0;
}
