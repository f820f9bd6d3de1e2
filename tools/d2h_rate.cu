// d2h_rate.cu -- PCIe rate of cudaMemcpy2DAsync device -> pinned host for the host-buffer call's copy-out: a device chunk
// [S, n*128] bf16 into the token-major host result [S, 24*128] (rows of n*256 bytes, pitch 6144 bytes), and the
// contiguous copy for comparison.  nvcc -O2 -o d2h_rate d2h_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
int main() {
  const size_t S = 119056, H = 24, D = 128;
  char *host, *dev;
  cudaMallocHost(&host, S * H * D * 2);
  cudaMalloc(&dev, S * H * D * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int n : {1, 2, 3, 4, 8, 24}) {
    for (int mode = 0; mode < 2; ++mode) {  // 0: 2-D strided, 1: contiguous of the same size
      float ms = 0;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) cudaMemcpy2DAsync(host, H * D * 2, dev, n * D * 2, n * D * 2, S, cudaMemcpyDeviceToHost, 0);
        else cudaMemcpyAsync(host, dev, S * n * D * 2, cudaMemcpyDeviceToHost, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("n=%2d %s  %8.3f ms  %6.1f GB/s\n", n, mode ? "contiguous" : "2-D rows  ", ms, S * n * D * 2 / ms / 1e6);
    }
  }
  return 0;
}
