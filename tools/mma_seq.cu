// Micro-benchmark (bring-up tool): the exact MMA sequence of attn_tc5 (PV_s then QK_s, slots alternating) without
// any TMA or softmax, to separate tensor-pipe hazards from everything else.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace rsa::ptx;
constexpr int kData = 14 * 16384;

// ALIAS: P (A operand of PV) lives in the S columns QK overwrites next (as in the kernel) or in the O region of the
// other slot (no write-after-read between consecutive MMAs).  COMMITS: tcgen05.commit after every group.
// WAITS: a try_wait on an already-completed barrier before every group.
template <bool ALIAS, bool COMMITS, bool WAITS, bool QK_TS>
__global__ void __launch_bounds__(128, 1) seq_kernel(int rounds, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* tmem_slot = (uint32_t*)(smem + kData);
  const uint32_t barw = smem_u32(smem + kData + 8), bard = smem_u32(smem + kData + 16), barc = smem_u32(smem + kData + 24);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kData / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  __syncthreads();
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(barw, 1); mbar_init(bard, 1); mbar_init(barc, 1); fence_barrier_init(); mbar_arrive(barc); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t sb = smem_u32(smem);
    constexpr uint32_t idQK = umma_idesc_bf16(128, 128, false), idPV = umma_idesc_bf16(128, 128, true);
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      long long t0 = clock64();
      for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const uint32_t tS = tmem + 128 * s, tO = tmem + 256 + 128 * s;
          const uint32_t tP = ALIAS ? tS : tmem + 256 + 128 * (s ^ 1);
          const uint64_t vd = smem_desc_sw128(sb + 65536 + ((2 * r + s) % 5) * 32768, 16384);
          const uint64_t kd = smem_desc_sw128(sb + 65536 + ((2 * r + s + 2) % 5) * 32768);
          const uint64_t qd = smem_desc_sw128(sb + s * 32768);
          if (WAITS) { mbar_wait(barc, 0); mbar_wait(barc, 0); tc_fence_after(); }
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idPV, 1u);
          }
          if (WAITS) { mbar_wait(barc, 0); tc_fence_after(); }
          if (leader) {
#pragma unroll
            for (int ks = 4; ks < 8; ++ks) umma_ts(tO, tP + ks * 8, vd + 128 * ks, idPV, 1u);
            if (COMMITS) umma_commit(bard);
          }
          if (WAITS) { mbar_wait(barc, 0); tc_fence_after(); }
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t off = (ks >> 2) * 1024 + 2 * (ks & 3);
              if (QK_TS) umma_ts(tS, tO + ks * 8, kd + off, idQK, ks != 0);
              else umma_ss(tS, qd + off, kd + off, idQK, ks != 0);
            }
            if (COMMITS) { umma_commit(bard); umma_commit(bard); }
          }
          __syncwarp();
        }
      }
      if (leader) umma_commit(barw);
      __syncwarp();
      mbar_wait(barw, rep & 1);
      long long t2 = clock64();
      if (t2 - t0 < best) { best = t2 - t0; if (leader) out[blockIdx.x] = t2 - t0; }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <bool ALIAS, bool COMMITS, bool WAITS, bool QK_TS>
void run(int grid, long long* out) {
  const int rounds = 32, smem = kData + 64;
  cudaFuncSetAttribute(seq_kernel<ALIAS, COMMITS, WAITS, QK_TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  seq_kernel<ALIAS, COMMITS, WAITS, QK_TS><<<grid, 128, smem>>>(rounds, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  double tot = 0;
  for (int b = 0; b < grid; ++b) tot += out[b];
  printf("grid %3d alias %d commits %d waits %d qk_ts %d: %.0f cycles per round (2 slots x 16 MMAs; ideal 2048)\n", grid, ALIAS,
         COMMITS, WAITS, QK_TS, tot / grid / rounds);
}

int main() {
  long long* out; cudaMallocManaged(&out, 148 * sizeof(long long));
  for (int grid : {1, 148}) {
    run<false, false, false, false>(grid, out);
    run<true, false, false, false>(grid, out);
    run<true, true, false, false>(grid, out);
    run<true, false, true, false>(grid, out);
    run<true, true, true, false>(grid, out);
    run<false, true, true, false>(grid, out);
    run<true, true, true, true>(grid, out);
  }
  return 0;
}
