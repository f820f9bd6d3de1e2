// rect_c.cu -- kernel 3c: C = W_skip . Vp, the pooled contribution of the skipped blocks.
// Replaces torch.matmul(attn_pool_novalid, value_pool) (reference rectified_wan21_attn.py:336-338,
// rectified_hunyuan_attn.py:355-357); the reference then repeat_interleaves C to token granularity
// ([B,H,S,D]) -- here it stays [BH, NQ, 128] fp32 and is consumed by kernel 4's epilogue.
// Small fp32 SIMT GEMM ([NQ x n_ent] x [n_ent x 128] per head, ~5 GFLOP at the HunyuanVideo size).
// CTA: 32 query blocks x 128 channels, 256 threads, 4 rows x 4 channels per thread.
#include "rsa_common.cuh"

namespace rsa {
namespace {

constexpr int kThreads = 256;
constexpr int TM = 32, TK = 32;

__global__ void __launch_bounds__(kThreads) rect_c_kernel(const float* __restrict__ w, const float* __restrict__ vp,
                                                          float* __restrict__ c, int nq, int nqt, int n_ent,
                                                          int ent_ld, int nb) {
  __shared__ float s_w[TM][TK + 1];
  __shared__ __align__(16) float s_v[TK][128];
  const int bh = blockIdx.y, i0 = blockIdx.x * TM;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  float2 acc[4][2];  // packed fma.rn.f32x2: channels (4 tx, 4 tx + 1) and (4 tx + 2, 4 tx + 3)
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = make_float2(0.f, 0.f);
  const float* wb = w + ((int64_t)bh * nq + i0) * ent_ld;
  const float* vb = vp + (int64_t)bh * nb * 128;
  for (int j0 = 0; j0 < n_ent; j0 += TK) {
    __syncthreads();
    // W tile: 32 x 32
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = ty + 8 * p, jj = j0 + tx;
      s_w[r][tx] = (i0 + r < nq && jj < n_ent) ? wb[(int64_t)r * ent_ld + jj] : 0.f;
    }
    // Vp tile: 32 x 128
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = ty + 8 * p, jj = j0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (jj < n_ent) v = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)jj * 128 + 4 * tx));
      *reinterpret_cast<float4*>(&s_v[r][4 * tx]) = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < TK; ++j) {
      const float4 v4 = *reinterpret_cast<const float4*>(&s_v[j][4 * tx]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float wv = s_w[4 * ty + u][j];
        const float2 ww = make_float2(wv, wv);
        acc[u][0] = __ffma2_rn(ww, make_float2(v4.x, v4.y), acc[u][0]);
        acc[u][1] = __ffma2_rn(ww, make_float2(v4.z, v4.w), acc[u][1]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + 4 * ty + u;
    if (i < nq)
      *reinterpret_cast<float4*>(c + ((int64_t)bh * nqt + i) * 128 + 4 * tx) =
          make_float4(acc[u][0].x, acc[u][0].y, acc[u][1].x, acc[u][1].y);
  }
}

}  // namespace

int launch_rect_c(const rsa_attn_desc* d, char* ws, const WsLayout& L, cudaStream_t s) {
  (void)d;
  if (L.nq == 0) return RSA_OK;
  dim3 grid((L.nq + TM - 1) / TM, L.bh);
  rect_c_kernel<<<grid, kThreads, 0, s>>>((const float*)(ws + L.off_w), (const float*)(ws + L.off_v_pool),
                                          (float*)(ws + L.off_C), L.nq, L.nqt, L.n_entries, L.ent_ld, L.nb);
  RSA_CUDA_CHECK(cudaGetLastError());
  return RSA_OK;
}

}  // namespace rsa
