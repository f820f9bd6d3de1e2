// pool_stats.cu -- kernel 2: block mean pooling of Q, K, V and the GAPR deviation statistics in ONE read.
// Replaces, per 128-token block (reference paths relative to the reference root):
//   query_pool / key_pool / value_pool = X_blocks.mean(dim=-2)    rectified_wan21_attn.py:189-192, :337
//   delta = X_blocks - pool; delta.abs().mean(dim=-2)             rectified_spaattn/gapr_mask.py:19-30
//   key_text = key[:, :, NQ*128 : NQ*128 + attenable]             rectified_hunyuan_attn.py:193-194
// The reference re-reads Q and K four times and materialises two full-size temporaries; here every element of
// Q, K, V is read from HBM exactly once (algorithmic bytes = 3*S*H*D*2) and stays in registers for the
// second (deviation) sweep.
//
// One CTA (256 threads) per (block, head, tensor).  Thread (warp w, lane l) owns rows 16w + 2*it + (l>>4),
// it = 0..7, and the eight columns 8*(l&15) .. +7.  Summation order (the oracle replicates it bit for bit):
//   lane chain over it = 0..7 (sequential fp32) -> half-warp pair (shfl xor 16) -> warps 0..7 (sequential).
// Rows >= valid_rows contribute zeros; the divisor is always 128 (the reference zero-pads, wan21 :299-302).
// Rows are counted in the padded layout (rsa_common.cuh RowMap): a visual block reads memory rows below vis_len and
// zeros above, a text block reads memory rows shifted down by the gap.
#include "rsa_common.cuh"

namespace rsa {
namespace {

constexpr int kThreads = 256;

struct PoolArgs {
  const __nv_bfloat16* x[3];  // q, k, v
  int64_t stride[3][3];       // (batch, head, token) element strides
  int valid_rows[3];          // padded-layout rows >= this are zeros
  int vis_len, nq_vis, gap;   // RowMap
  int n_blk[3];               // blocks to pool per tensor (NQ, NQ, NB)
  float* mean[3];             // q_pool, k_cat, v_pool
  float* mad[3];              // q_mad, k_mad, nullptr
  int out_rows[3];            // rows per head of the mean arrays (NQ, NKC, NB)
  int heads;
  // text keys copied as fp32 rows behind the pooled keys
  int text_keys, text_from;   // a, memory row of the first text token (= vis_len)
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(kThreads) pool_stats_kernel(const PoolArgs a) {
  const int which = blockIdx.z;
  const int bh = blockIdx.y;
  const int blk = blockIdx.x;
  const int b = bh / a.heads, h = bh % a.heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (which == 3) {
    // text keys -> fp32 rows [NQ, NQ + a) of k_cat; one 16-lane group per row
    const int rows_per_cta = kThreads / 16;
    const int t = blk * rows_per_cta + (tid >> 4);
    if (t < a.text_keys) {
      const int tok = a.text_from + t;
      const __nv_bfloat16* src = a.x[1] + b * a.stride[1][0] + h * a.stride[1][1] + (int64_t)tok * a.stride[1][2];
      float f[8];
      uint4 u = make_uint4(0, 0, 0, 0);
      if (tok + a.gap < a.valid_rows[1]) u = *reinterpret_cast<const uint4*>(src + 8 * (tid & 15));
      unpack8(u, f);
      float* dst = a.mean[1] + ((int64_t)bh * a.out_rows[1] + a.n_blk[1] + t) * 128 + 8 * (tid & 15);
      reinterpret_cast<float4*>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    return;
  }
  if (blk >= a.n_blk[which]) return;

  __shared__ float s_part[8][128];
  __shared__ float s_mean[128];

  const __nv_bfloat16* base = a.x[which] + b * a.stride[which][0] + h * a.stride[which][1];
  const int64_t ts = a.stride[which][2];
  const int col = 8 * (lane & 15);
  const int row0 = blk * 128 + 16 * warp + (lane >> 4);
  const bool visual = blk < a.nq_vis;
  const int valid = visual ? min(a.valid_rows[which], a.vis_len) : a.valid_rows[which];
  const int shift = visual ? 0 : a.gap;  // padded row -> memory row

  uint4 raw[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = row0 + 2 * it;
    raw[it] = make_uint4(0, 0, 0, 0);
    if (r < valid) raw[it] = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(r - shift) * ts + col));
  }

  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    float f[8];
    unpack8(raw[it], f);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = __fadd_rn(acc[c], f[c]);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = __fadd_rn(acc[c], __shfl_xor_sync(0xffffffffu, acc[c], 16));
  if (lane < 16) {
#pragma unroll
    for (int c = 0; c < 8; ++c) s_part[warp][col + c] = acc[c];
  }
  __syncthreads();
  if (tid < 128) {
    float t = s_part[0][tid];
#pragma unroll
    for (int w = 1; w < 8; ++w) t = __fadd_rn(t, s_part[w][tid]);
    const float m = __fmul_rn(t, 1.0f / 128.0f);
    s_mean[tid] = m;
    a.mean[which][((int64_t)bh * a.out_rows[which] + blk) * 128 + tid] = m;
  }
  if (a.mad[which] == nullptr) return;
  __syncthreads();

  float mean[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) mean[c] = s_mean[col + c];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    float f[8];
    unpack8(raw[it], f);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = __fadd_rn(acc[c], fabsf(__fsub_rn(f[c], mean[c])));
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = __fadd_rn(acc[c], __shfl_xor_sync(0xffffffffu, acc[c], 16));
  __syncthreads();  // s_part reuse
  if (lane < 16) {
#pragma unroll
    for (int c = 0; c < 8; ++c) s_part[warp][col + c] = acc[c];
  }
  __syncthreads();
  if (tid < 128) {
    float t = s_part[0][tid];
#pragma unroll
    for (int w = 1; w < 8; ++w) t = __fadd_rn(t, s_part[w][tid]);
    a.mad[which][((int64_t)bh * a.n_blk[which] + blk) * 128 + tid] = __fmul_rn(t, 1.0f / 128.0f);
  }
}

}  // namespace

int launch_pool_stats(const rsa_attn_desc* d, const void* q, const void* k, const void* v, char* ws,
                      const WsLayout& L, cudaStream_t s) {
  PoolArgs a;
  a.x[0] = (const __nv_bfloat16*)q;
  a.x[1] = (const __nv_bfloat16*)k;
  a.x[2] = (const __nv_bfloat16*)v;
  for (int i = 0; i < 3; ++i) {
    a.stride[0][i] = d->q_stride[i];
    a.stride[1][i] = d->k_stride[i];
    a.stride[2][i] = d->v_stride[i];
  }
  const RowMap rm = row_map(d);
  const int seq_v = d->seq + rm.gap;
  const int kvz = d->kv_zero_from < seq_v ? d->kv_zero_from : seq_v;
  a.vis_len = rm.vis_len;
  a.nq_vis = rm.nq_vis;
  a.gap = rm.gap;
  a.valid_rows[0] = seq_v;
  a.valid_rows[1] = kvz;
  a.valid_rows[2] = kvz;
  a.n_blk[0] = L.nq;
  a.n_blk[1] = L.nq;
  a.n_blk[2] = L.nb;
  a.mean[0] = (float*)(ws + L.off_q_pool);
  a.mean[1] = (float*)(ws + L.off_k_cat);
  a.mean[2] = (float*)(ws + L.off_v_pool);
  a.mad[0] = (float*)(ws + L.off_q_mad);
  a.mad[1] = (float*)(ws + L.off_k_mad);
  a.mad[2] = nullptr;
  a.out_rows[0] = L.nq;
  a.out_rows[1] = L.nkc;
  a.out_rows[2] = L.nb;
  a.heads = d->heads;
  a.text_keys = L.a;
  a.text_from = rm.vis_len;
  const int text_ctas = (L.a + 15) / 16;
  const int gx = L.nb > text_ctas ? L.nb : text_ctas;
  dim3 grid(gx, L.bh, L.a > 0 ? 4 : 3);
  pool_stats_kernel<<<grid, kThreads, 0, s>>>(a);
  RSA_CUDA_CHECK(cudaGetLastError());
  return RSA_OK;
}

}  // namespace rsa
