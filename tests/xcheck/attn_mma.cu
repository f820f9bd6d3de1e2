// attn_mma.cu -- CROSS-CHECK implementation of kernel 4 on the legacy mma.sync (HMMA) path.
// TEST INFRASTRUCTURE (tests/xcheck/librsa_xcheck.so), not part of the product library: it exists so the tcgen05 kernel
// (csrc/attn_tc5.cu) can be compared against an independent CUDA implementation at full BASELINE.json sizes, where the
// CPU oracle is too slow.  Entry point: rsa_xcheck_attention (xcheck_api.cu), driven by tests/xcheck/__init__.py.
//
// Same contract as attn_tc5.cu: per 128-row query tile walk the ascending kept-block list, online softmax in
// fp32 with exp2 (reference rectified_wan21_attn.py:56-105), keys >= kv_len masked to -inf (:86-87), then the
// fused rectification epilogue O = Os * R + C (wan21 :346) written straight to [B, S, H, D].
#include <math.h>

#include "rsa_common.cuh"

namespace rsa {
namespace {

constexpr int kThreads = 256;
constexpr int kTileBytes = 128 * 256;  // 128 rows x 128 bf16
constexpr int kSmem = kTileBytes * 5;  // Q + 2 x (K, V)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// row r, 16-byte chunk c (0..15) -> byte offset inside a tile (XOR swizzle keeps ldmatrix conflict-free)
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 256 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void load_tile(uint32_t dst, const __nv_bfloat16* base, int64_t stride, int row0,
                                          int rows_total, int tid) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int idx = tid + p * kThreads;
    const int r = idx >> 4, c = idx & 15;
    const int gr = row0 + r;
    const bool ok = gr < rows_total;
    const __nv_bfloat16* src = base + (int64_t)(ok ? gr : 0) * stride + c * 8;
    cp_async16(dst + tile_off(r, c), src, ok ? 16 : 0);
  }
}

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1) attn_mma_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tile = blockIdx.x, bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;

  const __nv_bfloat16* qb = a.q + b * a.qs[0] + h * a.qs[1];
  const __nv_bfloat16* kb = a.k + b * a.ks[0] + h * a.ks[1];
  const __nv_bfloat16* vb = a.v + b * a.vs[0] + h * a.vs[1];
  __nv_bfloat16* ob = a.o + b * a.os[0] + h * a.os[1];

  const int64_t lrow = (int64_t)bh * a.nqt + tile;
  const int cnt = min(max(a.kept_cnt[lrow], 0), a.nb);
  const uint16_t* list = a.kept_idx + lrow * a.nb;
  const int row0 = tile * 128;
  const bool prescale = a.q_round != 0 && tile < a.nq_vis;  // as attn_tc5.cu: q~ rounding on visual tiles only
  const float sc = prescale ? 1.f : a.scale_log2;

  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK0 = sQ + kTileBytes, sV0 = sQ + 2 * kTileBytes;
  // stage s: K at sK0 + s*2*kTileBytes, V at sV0 + s*2*kTileBytes

  load_tile(sQ, qb, a.qs[2], row0, a.seq_q, tid);
  if (cnt > 0) {
    const int kv0 = (int)list[0] * 128;
    load_tile(sK0, kb, a.ks[2], kv0, a.seq_kv, tid);
    load_tile(sV0, vb, a.vs[2], kv0, a.seq_kv, tid);
  }
  cp_commit();

  float o[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  uint32_t qf[8][4];

  for (int it = 0; it < cnt; ++it) {
    const int st = it & 1;
    if (it + 1 < cnt) {
      const int kvn = (int)list[it + 1] * 128;
      load_tile(sK0 + (st ^ 1) * 2 * kTileBytes, kb, a.ks[2], kvn, a.seq_kv, tid);
      load_tile(sV0 + (st ^ 1) * 2 * kTileBytes, vb, a.vs[2], kvn, a.seq_kv, tid);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int r = warp * 16 + (lane & 15);
        ldsm4(sQ + tile_off(r, 2 * ks + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        // q~ = bf16(q * sm_scale * log2 e): the reference kernel's rounding of the pre-scaled query (wan21 :61-62)
        if (prescale) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            qf[ks][j] = pack_bf16(__uint_as_float(qf[ks][j] << 16) * a.scale_log2,
                                  __uint_as_float(qf[ks][j] & 0xffff0000u) * a.scale_log2);
        }
      }
    }
    const uint32_t sK = sK0 + st * 2 * kTileBytes, sV = sV0 + st * 2 * kTileBytes;

    // S = Q K^T   (16 x 128 per warp)
    float s[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        const int mid = lane >> 3;
        const int r = np * 16 + (lane & 7) + 8 * (mid >> 1);
        uint32_t b0, b1, b2, b3;
        ldsm4(sK + tile_off(r, 2 * ks + (mid & 1)), b0, b1, b2, b3);
        mma16816(s[2 * np], qf[ks], b0, b1);
        mma16816(s[2 * np + 1], qf[ks], b2, b3);
      }
    }
    // scale + key-validity mask
    const int kv0 = (int)list[it] * 128;
    const int lim = a.kv_len - kv0;  // columns >= lim are invalid
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int c = n * 8 + 2 * t4;
      s[n][0] = (c < lim) ? s[n][0] * sc : -INFINITY;
      s[n][1] = (c + 1 < lim) ? s[n][1] * sc : -INFINITY;
      s[n][2] = (c < lim) ? s[n][2] * sc : -INFINITY;
      s[n][3] = (c + 1 < lim) ? s[n][3] * sc : -INFINITY;
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
      mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float al0 = (m0 == -INFINITY) ? 0.f : exp2f(m0 - mx0);
    const float al1 = (m1 == -INFINITY) ? 0.f : exp2f(m1 - mx1);
    const float sub0 = (mx0 == -INFINITY) ? 0.f : mx0;
    const float sub1 = (mx1 == -INFINITY) ? 0.f : mx1;
    m0 = mx0;
    m1 = mx1;
    l0 *= al0;
    l1 *= al1;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      o[n][0] *= al0;
      o[n][1] *= al0;
      o[n][2] *= al1;
      o[n][3] *= al1;
    }
    uint32_t pf[8][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const float p0 = exp2f(s[n][0] - sub0), p1 = exp2f(s[n][1] - sub0);
      const float p2 = exp2f(s[n][2] - sub1), p3 = exp2f(s[n][3] - sub1);
      l0 += p0 + p1;
      l1 += p2 + p3;
      pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    // O += P V
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        const int mid = lane >> 3;
        const int r = ks * 16 + (lane & 7) + 8 * (mid & 1);
        uint32_t b0, b1, b2, b3;
        ldsm4t(sV + tile_off(r, 2 * np + (mid >> 1)), b0, b1, b2, b3);
        mma16816(o[2 * np], pf[ks], b0, b1);
        mma16816(o[2 * np + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();
  }
  if (cnt == 0) cp_wait<0>();

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float R = a.R ? a.R[lrow] : 1.f;
  const float i0 = l0 > 0.f ? R / l0 : 0.f, i1 = l1 > 0.f ? R / l1 : 0.f;
  const float* crow = a.C ? a.C + lrow * 128 : nullptr;
  const int r0 = row0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    const int c = n * 8 + 2 * t4;
    const float c0 = crow ? crow[c] : 0.f, c1 = crow ? crow[c + 1] : 0.f;
    if (r0 < a.seq_q) {
      const bool z = r0 >= a.q_valid;
      *reinterpret_cast<uint32_t*>(ob + (int64_t)r0 * a.os[2] + c) =
          z ? 0u : pack_bf16(fmaf(o[n][0], i0, c0), fmaf(o[n][1], i0, c1));
    }
    if (r1 < a.seq_q) {
      const bool z = r1 >= a.q_valid;
      *reinterpret_cast<uint32_t*>(ob + (int64_t)r1 * a.os[2] + c) =
          z ? 0u : pack_bf16(fmaf(o[n][2], i1, c0), fmaf(o[n][3], i1, c1));
    }
  }
}

}  // namespace

int launch_attention_mma(const AttnArgs& a, cudaStream_t s) {
  if (a.gap != 0) RSA_FAIL(RSA_ERR_UNSUPPORTED, "the mma.sync cross-check kernel handles block-aligned visual segments only");
  static bool configured = false;
  if (!configured) {
    RSA_CUDA_CHECK(cudaFuncSetAttribute(attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  if (a.nqt == 0) return RSA_OK;
  dim3 grid(a.nqt, a.batch * a.heads);
  attn_mma_kernel<<<grid, kThreads, kSmem, s>>>(a);
  RSA_CUDA_CHECK(cudaGetLastError());
  return RSA_OK;
}

}  // namespace rsa
