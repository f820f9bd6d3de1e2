// block_scores.cu -- kernel 3a: pooled score products and the GAPR gain/error test in one pass.
// Replaces (reference paths relative to the reference root):
//   attention_scores_flat = bmm(query_pool, [key_pool; key_text]^T)   rectified_wan21_attn.py:203, hunyuan :205
//   dot_q = bmm(mean|dQ|, key_pool^T); dot_k = bmm(query_pool, mean|dK|^T)      gapr_mask.py:26, :32
//   ~(IQ*JK*|A| > IQ*JK*(|dot_q| + |dot_k|))                                    gapr_mask.py:27-42
// (the IQ*JK = 2^14 factors are exact scalings and cancel bit-exactly in fp32.)
//
// fp32 SIMT on purpose: the selection downstream is a discrete decision, so every (i, j) product is ONE fmaf chain
// over d = 0..127 in that order -- the oracle replicates it bit for bit.  ~16 GFLOP at the HunyuanVideo size;
// operands (55 MB of pooled statistics) are L2-resident.
//
// CTA tile 64 (query blocks) x 64 (key entries), 256 threads, 4x4 micro-tile per thread, d staged through
// shared memory in chunks of 32 with a [d][row] layout so the inner loop is 4 LDS.128 per 48 FMA; the operands are read
// from the TRANSPOSED copies of the pooled statistics, so staging is a straight coalesced copy.  The FMAs are
// issued as packed fma.rn.f32x2 (two key columns per instruction; each half is an ordinary IEEE fmaf, so the
// per-element chain and its rounding are unchanged) -- the scalar FFMA issues at half rate on sm_100.
#include "rsa_common.cuh"

namespace rsa {
namespace {

constexpr int kThreads = 256;
constexpr int TI = 64, TJ = 64, DK = 32;

struct ScoreArgs {
  const float *qp, *dq, *kc, *dk;  // transposed pooled statistics [BH][128][ldq | ldk] (transpose_stats_kernel below):
                                   // q means, q deviations, keys (pooled + text), k deviations
  float* scores;                   // [BH,NQ,score_ld]
  uint8_t* nogapr;                 // [BH,NQ,nogapr_ld]
  int nq, nkc, score_ld, nogapr_ld, ldq, ldk;
};

// 32 d x 64 blocks of one transposed array -> dst[d][row]: thread -> (d = idx / 16, four consecutive blocks), two float4
// per thread; a warp reads two 256-byte runs.  (Staging from the [block][128] layout made every lane of a load touch
// its own 128-byte line -- 8.5 LSU wavefronts per request, 37 % of the LSU data pipe that bounds this kernel.)
__device__ __forceinline__ void stage(float (*dst)[TI], const float* src, int ld, int rows_avail, int tid) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int idx = tid + kThreads * p;
    const int d = idx >> 4, row = 4 * (idx & 15);
    const float* g = src + (int64_t)d * ld + row;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows_avail) {  // ld is a multiple of 4, so the whole quad lies inside the array; entries past the end of
      v = __ldg(reinterpret_cast<const float4*>(g));  // the ragged last tile were never written and count as zeros
      if (row + 1 >= rows_avail) v.y = 0.f;
      if (row + 2 >= rows_avail) v.z = 0.f;
      if (row + 3 >= rows_avail) v.w = 0.f;
    }
    *reinterpret_cast<float4*>(&dst[d][row]) = v;
  }
}

// [BH][rows][128] -> [BH][128][ld] for the four operand arrays (z = 4 bh + array), 32 x 32 tiles through shared memory,
// coalesced both ways; columns [rows, ld) are written as zeros.  55 MB in, 46 MB out at the HunyuanVideo size.  (Having
// kernel 2 store the transposed values itself -- 128 four-byte stores per CTA at a stride of ld floats -- doubled that
// kernel's time at C3b.)
struct TransposeArgs {
  const float* src[4];
  float* dst[4];
  int rows[4], ld[4];
};
__global__ void __launch_bounds__(256) transpose_stats_kernel(const TransposeArgs a) {
  __shared__ float t[32][33];
  const int arr = blockIdx.z & 3, bh = blockIdx.z >> 2;
  // (selected, not indexed: a dynamically indexed parameter array is copied to local memory)
  const int rows = arr == 2 ? a.rows[2] : a.rows[0], ld = arr == 2 ? a.ld[2] : a.ld[0];
  const float* src0 = arr == 0 ? a.src[0] : arr == 1 ? a.src[1] : arr == 2 ? a.src[2] : a.src[3];
  float* dst0 = arr == 0 ? a.dst[0] : arr == 1 ? a.dst[1] : arr == 2 ? a.dst[2] : a.dst[3];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (r0 >= ld) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = src0 + (int64_t)bh * rows * 128;
  float* dst = dst0 + (int64_t)bh * 128 * ld;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k;
    t[ty + 8 * k][tx] = r < rows ? __ldg(src + (int64_t)r * 128 + c0 + tx) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < ld) dst[(int64_t)c * ld + r] = t[tx][ty + 8 * k];
  }
}

__global__ void __launch_bounds__(kThreads, 3) block_scores_kernel(const ScoreArgs a) {
  __shared__ __align__(16) float s_qp[DK][TI];
  __shared__ __align__(16) float s_dq[DK][TI];
  __shared__ __align__(16) float s_kp[DK][TJ];
  __shared__ __align__(16) float s_dk[DK][TJ];

  const int bh = blockIdx.z;
  const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  // Warp w covers rows [16 (w >> 1), +16) x columns [32 (w & 1), +32) of the tile; inside it lane = 8 * (row group) +
  // (column group), so one LDS.128 of a warp touches 4 distinct query quads (broadcast) or 8 consecutive key quads
  // (128 B): 4 shared-memory wavefronts per d step instead of 6 with a 2 x 16 lane layout.
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = 8 * (warp & 1) + (lane & 7), ty = 4 * (warp >> 1) + (lane >> 3);
  const bool need_gapr = j0 < a.nq;  // tiles that only hold text-key columns skip the error products

  const float* qp = a.qp + (int64_t)bh * 128 * a.ldq + i0;
  const float* dq = a.dq + (int64_t)bh * 128 * a.ldq + i0;
  const float* kc = a.kc + (int64_t)bh * 128 * a.ldk + j0;
  const float* dk = a.dk + (int64_t)bh * 128 * a.ldq + j0;
  const int rows_i = a.nq - i0;
  const int rows_j = a.nkc - j0;
  const int rows_jg = a.nq - j0;  // columns that have deviation statistics

  float2 A2[4][2], E12[4][2], E22[4][2];  // [query row u][key-column pair w] = columns (2w, 2w+1)
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < 2; ++w) A2[u][w] = E12[u][w] = E22[u][w] = make_float2(0.f, 0.f);

  for (int dc = 0; dc < 128; dc += DK) {
    __syncthreads();
    stage(s_qp, qp + (int64_t)dc * a.ldq, a.ldq, rows_i, tid);
    stage(s_kp, kc + (int64_t)dc * a.ldk, a.ldk, rows_j, tid);
    if (need_gapr) {
      stage(s_dq, dq + (int64_t)dc * a.ldq, a.ldq, rows_i, tid);
      stage(s_dk, dk + (int64_t)dc * a.ldq, a.ldq, rows_jg, tid);
    }
    __syncthreads();
    if (need_gapr) {
#pragma unroll 8
      for (int d = 0; d < DK; ++d) {
        const float4 q4 = *reinterpret_cast<const float4*>(&s_qp[d][4 * ty]);
        const float4 g4 = *reinterpret_cast<const float4*>(&s_dq[d][4 * ty]);
        const float4 k4 = *reinterpret_cast<const float4*>(&s_kp[d][4 * tx]);
        const float4 h4 = *reinterpret_cast<const float4*>(&s_dk[d][4 * tx]);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
        const float2 k2[2] = {make_float2(k4.x, k4.y), make_float2(k4.z, k4.w)};
        const float2 h2[2] = {make_float2(h4.x, h4.y), make_float2(h4.z, h4.w)};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 qq = make_float2(q[u], q[u]), gg = make_float2(g[u], g[u]);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            A2[u][w] = __ffma2_rn(qq, k2[w], A2[u][w]);
            E12[u][w] = __ffma2_rn(gg, k2[w], E12[u][w]);
            E22[u][w] = __ffma2_rn(qq, h2[w], E22[u][w]);
          }
        }
      }
    } else {
#pragma unroll 8
      for (int d = 0; d < DK; ++d) {
        const float4 q4 = *reinterpret_cast<const float4*>(&s_qp[d][4 * ty]);
        const float4 k4 = *reinterpret_cast<const float4*>(&s_kp[d][4 * tx]);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        const float2 k2[2] = {make_float2(k4.x, k4.y), make_float2(k4.z, k4.w)};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 qq = make_float2(q[u], q[u]);
#pragma unroll
          for (int w = 0; w < 2; ++w) A2[u][w] = __ffma2_rn(qq, k2[w], A2[u][w]);
        }
      }
    }
  }

  float A[4][4], E1[4][4], E2[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      A[u][2 * w] = A2[u][w].x, A[u][2 * w + 1] = A2[u][w].y;
      E1[u][2 * w] = E12[u][w].x, E1[u][2 * w + 1] = E12[u][w].y;
      E2[u][2 * w] = E22[u][w].x, E2[u][2 * w + 1] = E22[u][w].y;
    }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + 4 * ty + u;
    if (i >= a.nq) continue;
    const int j = j0 + 4 * tx;
    float* srow = a.scores + ((int64_t)bh * a.nq + i) * a.score_ld + j;
    if (j + 3 < a.nkc) {
      *reinterpret_cast<float4*>(srow) = make_float4(A[u][0], A[u][1], A[u][2], A[u][3]);
    } else {
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (j + v < a.nkc) srow[v] = A[u][v];
    }
    if (need_gapr) {
      uint8_t* grow = a.nogapr + ((int64_t)bh * a.nq + i) * a.nogapr_ld + j;
      uint8_t g[4];
#pragma unroll
      for (int v = 0; v < 4; ++v)
        g[v] = (fabsf(A[u][v]) > __fadd_rn(fabsf(E1[u][v]), fabsf(E2[u][v]))) ? 0 : 1;
      if (j + 3 < a.nq) {
        *reinterpret_cast<uchar4*>(grow) = make_uchar4(g[0], g[1], g[2], g[3]);
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (j + v < a.nq) grow[v] = g[v];
      }
    }
  }
}

}  // namespace

int launch_block_scores(const rsa_attn_desc* d, char* ws, const WsLayout& L, cudaStream_t s) {
  (void)d;
  ScoreArgs a;
  a.qp = (const float*)(ws + L.off_q_pool_t);
  a.dq = (const float*)(ws + L.off_q_mad_t);
  a.kc = (const float*)(ws + L.off_k_cat_t);
  a.dk = (const float*)(ws + L.off_k_mad_t);
  a.ldq = L.ldq;
  a.ldk = L.ldk;
  a.scores = (float*)(ws + L.off_scores);
  a.nogapr = (uint8_t*)(ws + L.off_nogapr);
  a.nq = L.nq;
  a.nkc = L.nkc;
  a.score_ld = L.score_ld;
  a.nogapr_ld = L.nogapr_ld;
  if (L.nq == 0) return RSA_OK;
  TransposeArgs t;
  const size_t so[4] = {L.off_q_pool, L.off_q_mad, L.off_k_cat, L.off_k_mad};
  const size_t dof[4] = {L.off_q_pool_t, L.off_q_mad_t, L.off_k_cat_t, L.off_k_mad_t};
  for (int i = 0; i < 4; ++i) {
    t.src[i] = (const float*)(ws + so[i]);
    t.dst[i] = (float*)(ws + dof[i]);
    t.rows[i] = i == 2 ? L.nkc : L.nq;
    t.ld[i] = i == 2 ? L.ldk : L.ldq;
  }
  dim3 tgrid((L.ldk + 31) / 32, 4, 4 * L.bh);
  transpose_stats_kernel<<<tgrid, 256, 0, s>>>(t);
  RSA_CUDA_CHECK(cudaGetLastError());
  dim3 grid((L.nkc + TJ - 1) / TJ, (L.nq + TI - 1) / TI, L.bh);
  block_scores_kernel<<<grid, kThreads, 0, s>>>(a);
  RSA_CUDA_CHECK(cudaGetLastError());
  return RSA_OK;
}

}  // namespace rsa
