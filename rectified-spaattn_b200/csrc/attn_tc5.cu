// attn_tc5.cu -- kernel 4, the product path: block-sparse attention over per-row kept-block lists on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands fed by TMA, with the rectification
// epilogue O = Os * R + C fused into the output write.
//
// Replaces (reference paths): _triton_block_sparse_attention_onehot + kernel (rectified_wan21_attn.py:16-168),
// the epilogue `output_normal * R + C` (:346), the flash-attn call for text query rows
// (rectified_hunyuan_attn.py:371-380; text tiles simply carry a dense list with R = 1, C = 0) and the final
// cat / permute / reshape (:383-387; the output is written straight into [B, S, H, D]).
//
// One CTA = one 128-row query tile of one (batch, head); two CTAs are co-resident per SM (each owns 256 of the
// 512 TMEM columns and 112 KB of shared memory), so while one CTA's softmax warps work on S the tensor pipe runs
// the other CTA's MMAs -- the ping-pong FlashAttention-style kernels build by hand falls out of the hardware
// arbitration.  Per kept block j of the tile's list:
//     warp 0 (TMA)      K_j, V_j tiles -> 16 KB shared-memory granules (128 rows x 64 bf16, 128-byte swizzle)
//     warp 1 (MMA)      S = Q K_j^T            8 x tcgen05.mma 128x128x16, A and B from shared memory
//     warps 4-7         S (TMEM) -> registers, running max with lazy rescale of O, p = exp2(s*c - m), row sums,
//                       P (bf16) -> TMEM over the S columns
//     warp 1 (MMA)      O += P V_j             16 x tcgen05.mma 128x64x16, A = P from TMEM, B = V_j (MN-major)
// and at the end warps 4-7 read O from TMEM, apply 1/l, R and C and store bf16 rows.
#include <cuda.h>
#include <math.h>

#include "ptx_sm100.cuh"
#include "rsa_common.cuh"

namespace rsa {
namespace {

using namespace ptx;

constexpr int kThreads = 256;
constexpr int kGranule = 16384;  // 128 rows x 64 bf16
constexpr int kRing = 5;         // K/V granules in flight
constexpr int kOffQ = 0;
constexpr int kOffRing = 2 * kGranule;
constexpr int kOffBar = kOffRing + kRing * kGranule;
constexpr int kSmemBytes = kOffBar + 128;
constexpr uint32_t kTmemCols = 256;  // S / P at [0,128), O at [128,256)
constexpr float kRescaleThreshold = 8.f;  // log2 units: O and l are rescaled only when the max grows by > 2^8

// barrier slots (8 bytes each) inside the kOffBar region
enum { B_QFULL = 0, B_SFULL = 1, B_PFULL = 2, B_OFULL = 3, B_KVFULL = 4, B_KVEMPTY = 4 + kRing, B_COUNT = 4 + 2 * kRing };
static_assert(B_COUNT * 8 + 4 <= 128, "barrier region");

constexpr uint32_t kIdescQK = umma_idesc_bf16(128, 128, false);
constexpr uint32_t kIdescPV = umma_idesc_bf16(128, 64, true);

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

__global__ void __launch_bounds__(kThreads, 2)
attn_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = a.nqt - 1 - (int)blockIdx.x;  // dense (text) tiles are the longest: schedule them first
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int64_t lrow = (int64_t)bh * a.nqt + tile;
  const int cnt = a.kept_cnt[lrow];
  const uint16_t* __restrict__ list = a.kept_idx + lrow * a.nb;

  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kOffBar;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + B_COUNT * 8);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(bar(B_QFULL), 1);
    mbar_init(bar(B_SFULL), 1);
    mbar_init(bar(B_PFULL), 128);
    mbar_init(bar(B_OFULL), 1);
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar(B_KVFULL + i), 1);
      mbar_init(bar(B_KVEMPTY + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32((const void*)tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0 && cnt > 0) {
      // ------------------------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_arrive_expect_tx(bar(B_QFULL), 2 * kGranule);
        tma_load_4d(sbase + kOffQ, &tmQ, bar(B_QFULL), 0, tile * 128, h, b);
        tma_load_4d(sbase + kOffQ + kGranule, &tmQ, bar(B_QFULL), 64, tile * 128, h, b);
      }
      int n = 0;  // granule counter: ring slot n % kRing, use n / kRing
      for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int mine = (i0 + lane < cnt) ? (int)list[i0 + lane] : 0;
        const int nn = min(32, cnt - i0);
        for (int j = 0; j < nn; ++j) {
          const int kv0 = __shfl_sync(0xffffffffu, mine, j) * 128;
          if (lane == 0) {
#pragma unroll
            for (int t = 0; t < 4; ++t, ++n) {  // K half 0, K half 1, V half 0, V half 1
              const int slot = n % kRing;
              mbar_wait(bar(B_KVEMPTY + slot), ((n / kRing) & 1) ^ 1);
              mbar_arrive_expect_tx(bar(B_KVFULL + slot), kGranule);
              tma_load_4d(sbase + kOffRing + slot * kGranule, t < 2 ? (const void*)&tmK : (const void*)&tmV,
                          bar(B_KVFULL + slot), (t & 1) * 64, kv0, h, b);
            }
          }
          __syncwarp();
        }
      }
    } else if (warp == 1 && lane == 0 && cnt > 0) {
      // -------------------------------------------------------------------------------------- MMA issuer
      const uint64_t qd0 = smem_desc_sw128(sbase + kOffQ), qd1 = smem_desc_sw128(sbase + kOffQ + kGranule);
      const uint32_t tS = tmem, tO = tmem + 128;
      mbar_wait(bar(B_QFULL), 0);
      int n = 0;
      for (int i = 0; i < cnt; ++i) {
        // S = Q K^T : head_dim halves 0 and 1 live in consecutive granules
#pragma unroll
        for (int hf = 0; hf < 2; ++hf, ++n) {
          const int slot = n % kRing;
          mbar_wait(bar(B_KVFULL + slot), (n / kRing) & 1);
          tc_fence_after();
          const uint64_t kd = smem_desc_sw128(sbase + kOffRing + slot * kGranule);
          const uint64_t qd = hf ? qd1 : qd0;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // 16 head_dim elements = 32 bytes inside the swizzle atom
            umma_ss(tS, qd + 2 * ks, kd + 2 * ks, kIdescQK, (hf | ks) != 0);
          umma_commit(bar(B_KVEMPTY + slot));
        }
        umma_commit(bar(B_SFULL));
        // O += P V : output columns [0,64) from V half 0, [64,128) from V half 1
        mbar_wait(bar(B_PFULL), i & 1);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf, ++n) {
          const int slot = n % kRing;
          mbar_wait(bar(B_KVFULL + slot), (n / kRing) & 1);
          tc_fence_after();
          const uint64_t vd = smem_desc_sw128(sbase + kOffRing + slot * kGranule);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // 16 keys = 16 rows of 128 bytes; P: 16 bf16 = 8 TMEM columns
            umma_ts(tO + hf * 64, tS + ks * 8, vd + 128 * ks, kIdescPV, (i | ks) != 0);
          umma_commit(bar(B_KVEMPTY + slot));
        }
      }
      umma_commit(bar(B_OFULL));
    }
  } else {
    // ------------------------------------------------------------------- softmax + epilogue warpgroup
    setmaxnreg_inc<216>();
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) belong to this warp
    const int row = quarter * 32 + lane;
    const uint32_t tS = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tO = tS + 128;
    const float scale = a.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    const int lim_last = cnt > 0 ? a.kv_len - (int)list[cnt - 1] * 128 : 128;  // valid keys in the last block
    const bool dbg = a.dbg != nullptr && lrow == 0;

    for (int i = 0; i < cnt; ++i) {
      mbar_wait(bar(B_SFULL), i & 1);
      tc_fence_after();
      uint32_t s[128];
      tmem_ld32(tS + 0, s + 0);
      tmem_ld32(tS + 32, s + 32);
      tmem_ld32(tS + 64, s + 64);
      tmem_ld32(tS + 96, s + 96);
      tmem_wait_ld();
      if (dbg && i == 0) {
#pragma unroll
        for (int c = 0; c < 128; ++c) a.dbg[row * 128 + c] = u2f(s[c]);
      }
      if (i == cnt - 1 && lim_last < 128) {  // keys >= kv_len -> -inf (wan21 :75-87)
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= lim_last) s[c] = 0xff800000u;
      }
      float mx0 = u2f(s[0]), mx1 = u2f(s[1]), mx2 = u2f(s[2]), mx3 = u2f(s[3]);
#pragma unroll
      for (int c = 4; c < 128; c += 4) {
        mx0 = fmaxf(mx0, u2f(s[c]));
        mx1 = fmaxf(mx1, u2f(s[c + 1]));
        mx2 = fmaxf(mx2, u2f(s[c + 2]));
        mx3 = fmaxf(mx3, u2f(s[c + 3]));
      }
      const float mx_s = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale;
      if (i == 0) {
        m_ref = mx_s;
      } else {
        const bool need = mx_s > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {  // tcgen05.ld/st are warp-collective: rescale the warp's 32 rows
          const float alpha = need ? ex2(m_ref - mx_s) : 1.f;
          if (need) m_ref = mx_s;
          l *= alpha;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint32_t o[32];
            tmem_ld32(tO + c4 * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = f2u(u2f(o[j]) * alpha);
            tmem_st32(tO + c4 * 32, o);
          }
        }
      }
      const float neg_m = -m_ref;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {  // 32 S columns -> 16 packed P columns
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = c16 * 32 + j;
          const float p0 = ex2(fmaf(u2f(s[c]), scale, neg_m));
          const float p1 = ex2(fmaf(u2f(s[c + 1]), scale, neg_m));
          const float p2 = ex2(fmaf(u2f(s[c + 2]), scale, neg_m));
          const float p3 = ex2(fmaf(u2f(s[c + 3]), scale, neg_m));
          l0 += p0;
          l1 += p1;
          l2 += p2;
          l3 += p3;
          pk[j / 2] = pack_bf16x2(p0, p1);
          pk[j / 2 + 1] = pack_bf16x2(p2, p3);
        }
        tmem_st16(tS + c16 * 16, pk);
      }
      l += (l0 + l1) + (l2 + l3);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar(B_PFULL));
    }

    // ------------------------------------------------------------------------------------------ epilogue
    const float R = a.R ? a.R[lrow] : 1.f;
    const float inv = l > 0.f ? R / l : 0.f;
    const float* __restrict__ crow = a.C ? a.C + lrow * 128 : nullptr;
    const int grow = tile * 128 + row;
    const bool store = grow < a.seq_q, zero = grow >= a.q_valid;
    __nv_bfloat16* orow = a.o + b * a.os[0] + h * a.os[1] + (int64_t)grow * a.os[2];
    if (cnt > 0) {
      mbar_wait(bar(B_OFULL), 0);
      tc_fence_after();
    }
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t o[32];
      if (cnt > 0) {
        tmem_ld32(tO + c4 * 32, o);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = 0u;
      }
      if (dbg) {
#pragma unroll
        for (int j = 0; j < 32; ++j) a.dbg[16384 + row * 128 + c4 * 32 + j] = u2f(o[j]);
        if (c4 == 0) {
          a.dbg[32768 + row] = l;
          a.dbg[32768 + 128 + row] = m_ref;
        }
      }
      if (store) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 32 + j + 2 * e;
            const float c0 = crow ? crow[c] : 0.f, c1 = crow ? crow[c + 1] : 0.f;
            w[e] = zero ? 0u : pack_bf16x2(fmaf(u2f(o[j + 2 * e]), inv, c0), fmaf(u2f(o[j + 2 * e + 1]), inv, c1));
          }
          *reinterpret_cast<uint4*>(orow + c4 * 32 + j) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// [batch, heads, rows, 128] bf16 view with element strides st = (batch, head, row) -> 4-D tensor map whose box is
// one granule: 64 head_dim elements x 128 rows, 128-byte swizzle; rows past the end read as zeros (the
// reference zero-pads to a multiple of 128, rectified_wan21_attn.py:299-302).
int make_map(CUtensorMap* m, const __nv_bfloat16* base, int batch, int heads, int rows, const int64_t* st) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) RSA_FAIL(RSA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if ((uintptr_t)base % 16) RSA_FAIL(RSA_ERR_UNSUPPORTED, "q/k/v must be 16-byte aligned");
  cuuint64_t dims[4] = {128, (cuuint64_t)rows, (cuuint64_t)heads, (cuuint64_t)batch};
  int64_t sr = st[2], sh = st[1], sb = st[0];
  if (sr < 128) RSA_FAIL(RSA_ERR_UNSUPPORTED, "token stride must be >= head_dim");
  if (heads == 1 && sh == 0) sh = sr * rows;  // a size-1 dimension's stride is never used but must be valid
  if (batch == 1 && sb == 0) sb = sh * heads > sr * rows ? sh * heads : sr * rows;
  if (sh <= 0 || sb <= 0) RSA_FAIL(RSA_ERR_UNSUPPORTED, "head/batch strides must be positive");
  cuuint64_t strides[3] = {(cuuint64_t)sr * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) RSA_FAIL(RSA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return RSA_OK;
}

}  // namespace

int launch_attention_tc5(const AttnArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    RSA_CUDA_CHECK(cudaFuncSetAttribute(attn_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    RSA_CUDA_CHECK(cudaFuncSetAttribute(attn_tc5_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  if (a.nqt == 0) return RSA_OK;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_map(&tq, a.q, a.batch, a.heads, a.seq_q, a.qs)) != RSA_OK) return rc;
  if ((rc = make_map(&tk, a.k, a.batch, a.heads, a.seq_kv, a.ks)) != RSA_OK) return rc;
  if ((rc = make_map(&tv, a.v, a.batch, a.heads, a.seq_kv, a.vs)) != RSA_OK) return rc;
  dim3 grid(a.nqt, a.batch * a.heads);
  attn_tc5_kernel<<<grid, kThreads, kSmemBytes, s>>>(tq, tk, tv, a);
  RSA_CUDA_CHECK(cudaGetLastError());
  return RSA_OK;
}

}  // namespace rsa
