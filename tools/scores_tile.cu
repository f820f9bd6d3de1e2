// scores_tile.cu -- experiment for kernel 3a's inner loop (not part of the library): the three pooled products
//   A = Qp Kp^T, E1 = dQ Kp^T, E2 = Qp dK^T   (one fmaf chain over d = 0..127 per output, as in block_scores.cu)
// on synthetic TRANSPOSED operands [128][n], with
//   variant 0: the library's tiling -- 256 threads, 64 x 64 tile, 4 x 4 outputs per thread: 4 LDS.128 per 24 FFMA2
//   variant 1: 128 threads, 64 x 64 tile, 8 x 4 outputs per thread: 6 LDS.128 per 48 FFMA2
// ncu on the library kernel (session 6): LSU data pipe 83 % of peak (74 % from these LDS), FMA pipe 68 % -- the loop is
// bound by shared-memory wavefronts, so fewer loads per FMA should pay if the 130 registers of variant 1 leave enough
// warps.  Prints ms per launch of both and checks that the outputs are bit-identical.
// Build on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/scores_tile tools/scores_tile.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int DK = 32, T = 64;

__device__ __forceinline__ void stage(float (*dst)[T], const float* src, int ld, int tid, int nthreads) {
  for (int idx = tid; idx < DK * (T / 4); idx += nthreads) {
    const int d = idx >> 4, row = 4 * (idx & 15);
    *reinterpret_cast<float4*>(&dst[d][row]) = __ldg(reinterpret_cast<const float4*>(src + (size_t)d * ld + row));
  }
}

template <int RU>  // rows per thread: 4 (256 threads) or 8 (128 threads)
__global__ void __launch_bounds__(RU == 4 ? 256 : 128, RU == 4 ? 3 : 3)
scores_kernel(const float* qp, const float* dq, const float* kp, const float* dk, float* A, float* E1, float* E2, int n) {
  __shared__ __align__(16) float s_qp[DK][T], s_dq[DK][T], s_kp[DK][T], s_dk[DK][T];
  constexpr int NT = RU == 4 ? 256 : 128;
  const int bh = blockIdx.z, i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // variant 0: warp = 16 rows x 32 cols (lane = 8 * row group + col group); variant 1: warp = 32 rows x 32 cols, two
  // warps side by side cover 64 columns, two on top of each other 64 rows
  const int tx = 8 * (warp & 1) + (lane & 7);
  const int ty = (RU == 4 ? 4 : 4) * (warp >> 1) + (lane >> 3);  // row group of RU rows
  const size_t base = (size_t)bh * 128 * n;
  float2 a2[RU][2], e1[RU][2], e2[RU][2];
#pragma unroll
  for (int u = 0; u < RU; ++u)
#pragma unroll
    for (int w = 0; w < 2; ++w) a2[u][w] = e1[u][w] = e2[u][w] = make_float2(0.f, 0.f);
  for (int dc = 0; dc < 128; dc += DK) {
    __syncthreads();
    stage(s_qp, qp + base + (size_t)dc * n + i0, n, tid, NT);
    stage(s_dq, dq + base + (size_t)dc * n + i0, n, tid, NT);
    stage(s_kp, kp + base + (size_t)dc * n + j0, n, tid, NT);
    stage(s_dk, dk + base + (size_t)dc * n + j0, n, tid, NT);
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < DK; ++d) {
      float q[RU], g[RU];
#pragma unroll
      for (int u4 = 0; u4 < RU; u4 += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(&s_qp[d][RU * ty + u4]);
        const float4 g4 = *reinterpret_cast<const float4*>(&s_dq[d][RU * ty + u4]);
        q[u4] = q4.x, q[u4 + 1] = q4.y, q[u4 + 2] = q4.z, q[u4 + 3] = q4.w;
        g[u4] = g4.x, g[u4 + 1] = g4.y, g[u4 + 2] = g4.z, g[u4 + 3] = g4.w;
      }
      const float4 k4 = *reinterpret_cast<const float4*>(&s_kp[d][4 * tx]);
      const float4 h4 = *reinterpret_cast<const float4*>(&s_dk[d][4 * tx]);
      const float2 k2[2] = {make_float2(k4.x, k4.y), make_float2(k4.z, k4.w)};
      const float2 h2[2] = {make_float2(h4.x, h4.y), make_float2(h4.z, h4.w)};
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const float2 qq = make_float2(q[u], q[u]), gg = make_float2(g[u], g[u]);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          a2[u][w] = __ffma2_rn(qq, k2[w], a2[u][w]);
          e1[u][w] = __ffma2_rn(gg, k2[w], e1[u][w]);
          e2[u][w] = __ffma2_rn(qq, h2[w], e2[u][w]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < RU; ++u) {
    const size_t o = ((size_t)bh * n + i0 + RU * ty + u) * n + j0 + 4 * tx;
    *reinterpret_cast<float4*>(A + o) = make_float4(a2[u][0].x, a2[u][0].y, a2[u][1].x, a2[u][1].y);
    *reinterpret_cast<float4*>(E1 + o) = make_float4(e1[u][0].x, e1[u][0].y, e1[u][1].x, e1[u][1].y);
    *reinterpret_cast<float4*>(E2 + o) = make_float4(e2[u][0].x, e2[u][0].y, e2[u][1].x, e2[u][1].y);
  }
}

int main() {
  const int n = 960, bh = 24;  // 15 x 15 tiles per head, the HunyuanVideo 129-frame size rounded up to the tile
  const size_t in_elems = (size_t)bh * 128 * n, out_elems = (size_t)bh * n * n;
  std::vector<float> h(in_elems);
  float *in[4], *out[2][3];
  for (int a = 0; a < 4; ++a) {
    srand(17 + a);
    for (auto& x : h) x = (float)(rand() % 2001 - 1000) * 1e-3f;
    cudaMalloc(&in[a], in_elems * 4);
    cudaMemcpy(in[a], h.data(), in_elems * 4, cudaMemcpyHostToDevice);
  }
  for (int v = 0; v < 2; ++v)
    for (int a = 0; a < 3; ++a) cudaMalloc(&out[v][a], out_elems * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const dim3 grid(n / T, n / T, bh);
  for (int v = 0; v < 2; ++v) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      if (v == 0) scores_kernel<4><<<grid, 256>>>(in[0], in[1], in[2], in[3], out[v][0], out[v][1], out[v][2], n);
      else scores_kernel<8><<<grid, 128>>>(in[0], in[1], in[2], in[3], out[v][0], out[v][1], out[v][2], n);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("variant %d (%s): %.4f ms per launch, %s\n", v, v ? "8x4, 128 threads" : "4x4, 256 threads", best,
           cudaGetErrorString(cudaGetLastError()));
  }
  std::vector<float> a(out_elems), b(out_elems);
  bool same = true;
  for (int t = 0; t < 3; ++t) {
    cudaMemcpy(a.data(), out[0][t], out_elems * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), out[1][t], out_elems * 4, cudaMemcpyDeviceToHost);
    same = same && memcmp(a.data(), b.data(), out_elems * 4) == 0;
  }
  printf("outputs bit-identical: %s\n", same ? "yes" : "NO");
  return same ? 0 : 1;
}
