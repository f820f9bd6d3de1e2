// host_call.cu -- rsa_rectified_attention_host: the whole attention call on HOST buffers.
//
// The reference's inner surface takes device tensors (rectified_hunyuan_attn.py:393-417); a caller that keeps Q/K/V
// in host memory (offloaded activations, or the end-to-end leg of bench.py) would copy 3 tensors in, call, and copy
// one out: at C3a that is 2.13 GB + 0.71 GB over PCIe around 31 ms of compute, all serial.  Heads are independent
// end to end, so this entry point cuts the call into chunks of `heads_per_chunk` heads and runs a
// three-stage pipeline on three streams -- H2D of chunk c+1, the five kernels of chunk c, D2H of chunk c-1 -- over
// double-buffered device staging carved from the caller's scratch.  No allocation, no host synchronisation: the
// function only enqueues; completion is ordered on the caller's stream.
#include <mutex>

#include "rsa_common.cuh"

namespace rsa {
namespace {

constexpr int kMaxDevices = 64;

struct DeviceLanes {
  bool ready = false;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t start = nullptr, in_ready[2] = {}, computed[2] = {}, out_done[2] = {};
};
DeviceLanes g_lanes[kMaxDevices];
std::mutex g_lanes_mutex;

int lanes_for_current_device(DeviceLanes** out) {
  int dev = 0;
  RSA_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) RSA_FAIL(RSA_ERR_UNSUPPORTED, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_lanes_mutex);
  DeviceLanes& l = g_lanes[dev];
  if (!l.ready) {
    RSA_CUDA_CHECK(cudaStreamCreateWithFlags(&l.h2d, cudaStreamNonBlocking));
    RSA_CUDA_CHECK(cudaStreamCreateWithFlags(&l.d2h, cudaStreamNonBlocking));
    RSA_CUDA_CHECK(cudaEventCreateWithFlags(&l.start, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      RSA_CUDA_CHECK(cudaEventCreateWithFlags(&l.in_ready[i], cudaEventDisableTiming));
      RSA_CUDA_CHECK(cudaEventCreateWithFlags(&l.computed[i], cudaEventDisableTiming));
      RSA_CUDA_CHECK(cudaEventCreateWithFlags(&l.out_done[i], cudaEventDisableTiming));
    }
    l.ready = true;
  }
  *out = &l;
  return RSA_OK;
}

struct HostPlan {
  int hc;             // heads per chunk
  int n_chunks;
  int body_heads;     // heads [0, body_heads) go in chunks of hc; the rest one head per chunk (see make_host_plan)
  int body_chunks;
  bool out_token_major;  // host out has heads adjacent inside a token ([B,S,H,D]): stage as [B,S,hc,D]
  size_t tensor_bytes;   // one staged chunk of q / k / v / out
  size_t ws_bytes;       // attention workspace of one chunk
  size_t off_in[2][3], off_out[2], off_ws, total;
};

rsa_attn_desc chunk_desc(const rsa_attn_desc* d, int heads, bool out_token_major) {
  rsa_attn_desc c = *d;
  c.heads = heads;
  const int64_t S = d->seq, D = d->head_dim;
  const int64_t in_st[3] = {heads * S * D, S * D, D};
  for (int i = 0; i < 3; ++i) c.q_stride[i] = c.k_stride[i] = c.v_stride[i] = in_st[i];
  if (out_token_major) {
    c.o_stride[0] = S * heads * D, c.o_stride[1] = D, c.o_stride[2] = heads * D;
  } else {
    for (int i = 0; i < 3; ++i) c.o_stride[i] = in_st[i];
  }
  return c;
}

int make_host_plan(const rsa_attn_desc* d, int heads_per_chunk, HostPlan* p) {
  int rc = validate_desc(d);
  if (rc != RSA_OK) return rc;
  if (heads_per_chunk < 1) RSA_FAIL(RSA_ERR_ARG, "heads_per_chunk must be >= 1");
  p->hc = heads_per_chunk < d->heads ? heads_per_chunk : d->heads;
  // The input copies are the bottleneck of the pipeline (PCIe), so its length is (all copies in) + (kernels and copy
  // out of the LAST chunk): the last hc heads are issued one per chunk to keep that tail short.
  const int tail = (p->hc > 1 && d->heads >= 3 * p->hc) ? p->hc : 0;
  p->body_heads = d->heads - tail;
  p->body_chunks = (p->body_heads + p->hc - 1) / p->hc;
  p->n_chunks = p->body_chunks + tail;
  p->out_token_major = d->o_stride[1] == d->head_dim;
  p->tensor_bytes = align_up((size_t)d->batch * p->hc * d->seq * d->head_dim * 2, 256);
  const rsa_attn_desc c = chunk_desc(d, p->hc, p->out_token_major);
  p->ws_bytes = align_up(make_layout(&c).total, 256);
  size_t o = 0;
  for (int s = 0; s < 2; ++s) {
    for (int t = 0; t < 3; ++t) p->off_in[s][t] = o, o += p->tensor_bytes;
    p->off_out[s] = o, o += p->tensor_bytes;
  }
  p->off_ws = o, o += p->ws_bytes;
  p->total = o;
  return RSA_OK;
}

// host [B, H, S, D] view (element strides hs) heads [h0, h0+n) -> device [B, n, S, D] contiguous
int copy_in(const __nv_bfloat16* host, const int64_t* hs, char* dev, int batch, int h0, int n, int64_t S, int64_t D,
            cudaStream_t s) {
  const size_t head_bytes = (size_t)S * D * 2;
  for (int b = 0; b < batch; ++b) {
    const __nv_bfloat16* src0 = host + b * hs[0] + (int64_t)h0 * hs[1];
    char* dst0 = dev + (size_t)b * n * head_bytes;
    if (hs[2] == D && (n == 1 || hs[1] == S * D)) {
      RSA_CUDA_CHECK(cudaMemcpyAsync(dst0, src0, n * head_bytes, cudaMemcpyHostToDevice, s));
      continue;
    }
    for (int h = 0; h < n; ++h) {
      const __nv_bfloat16* src = src0 + (int64_t)h * hs[1];
      char* dst = dst0 + (size_t)h * head_bytes;
      if (hs[2] == D)
        RSA_CUDA_CHECK(cudaMemcpyAsync(dst, src, head_bytes, cudaMemcpyHostToDevice, s));
      else
        RSA_CUDA_CHECK(cudaMemcpy2DAsync(dst, D * 2, src, hs[2] * 2, D * 2, S, cudaMemcpyHostToDevice, s));
    }
  }
  return RSA_OK;
}

// device chunk ([B, S, n, D] if token_major else [B, n, S, D]) -> host view with element strides os
int copy_out(const char* dev, __nv_bfloat16* host, const int64_t* os, bool token_major, int batch, int h0, int n,
             int64_t S, int64_t D, cudaStream_t s) {
  for (int b = 0; b < batch; ++b) {
    const char* src0 = dev + (size_t)b * S * n * D * 2;
    __nv_bfloat16* dst0 = host + b * os[0] + (int64_t)h0 * os[1];
    if (token_major) {
      RSA_CUDA_CHECK(cudaMemcpy2DAsync(dst0, os[2] * 2, src0, n * D * 2, n * D * 2, S, cudaMemcpyDeviceToHost, s));
      continue;
    }
    for (int h = 0; h < n; ++h) {
      const char* src = src0 + (size_t)h * S * D * 2;
      __nv_bfloat16* dst = dst0 + (int64_t)h * os[1];
      if (os[2] == D)
        RSA_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)S * D * 2, cudaMemcpyDeviceToHost, s));
      else
        RSA_CUDA_CHECK(cudaMemcpy2DAsync(dst, os[2] * 2, src, D * 2, D * 2, S, cudaMemcpyDeviceToHost, s));
    }
  }
  return RSA_OK;
}

}  // namespace
}  // namespace rsa

using namespace rsa;

extern "C" size_t rsa_host_call_scratch_bytes(const rsa_attn_desc* d, int heads_per_chunk) {
  HostPlan p;
  if (make_host_plan(d, heads_per_chunk, &p) != RSA_OK) return 0;
  return p.total;
}

extern "C" int rsa_rectified_attention_host(const rsa_attn_desc* d, const void* q, const void* k, const void* v,
                                            void* out, int heads_per_chunk, void* device_scratch,
                                            size_t scratch_bytes, void* stream) {
  HostPlan p;
  int rc = make_host_plan(d, heads_per_chunk, &p);
  if (rc != RSA_OK) return rc;
  if (!q || !k || !v || !out) RSA_FAIL(RSA_ERR_ARG, "rsa_rectified_attention_host: null tensor");
  if (!device_scratch || (uintptr_t)device_scratch % 256) RSA_FAIL(RSA_ERR_WORKSPACE, "scratch must be a 256-byte aligned device pointer");
  if (scratch_bytes < p.total) RSA_FAIL(RSA_ERR_WORKSPACE, "scratch too small: %zu < %zu", scratch_bytes, p.total);
  DeviceLanes* lanes = nullptr;
  if ((rc = lanes_for_current_device(&lanes)) != RSA_OK) return rc;
  cudaStream_t cs = (cudaStream_t)stream;
  char* base = (char*)device_scratch;
  const void* host_in[3] = {q, k, v};
  const int64_t* host_st[3] = {d->q_stride, d->k_stride, d->v_stride};

  // whatever the caller's stream still does with the scratch must finish before the copy lanes touch it
  RSA_CUDA_CHECK(cudaEventRecord(lanes->start, cs));
  RSA_CUDA_CHECK(cudaStreamWaitEvent(lanes->h2d, lanes->start, 0));
  RSA_CUDA_CHECK(cudaStreamWaitEvent(lanes->d2h, lanes->start, 0));

  for (int c = 0; c < p.n_chunks; ++c) {
    const int slot = c & 1;
    const bool body = c < p.body_chunks;
    const int h0 = body ? c * p.hc : p.body_heads + (c - p.body_chunks);
    const int n = !body ? 1 : (p.body_heads - h0 < p.hc ? p.body_heads - h0 : p.hc);
    // H2D lane: the slot's input staging is free once chunk c-2 has been computed
    if (c >= 2) RSA_CUDA_CHECK(cudaStreamWaitEvent(lanes->h2d, lanes->computed[slot], 0));
    for (int t = 0; t < 3; ++t)
      if ((rc = copy_in((const __nv_bfloat16*)host_in[t], host_st[t], base + p.off_in[slot][t], d->batch, h0, n,
                        d->seq, d->head_dim, lanes->h2d)) != RSA_OK)
        return rc;
    RSA_CUDA_CHECK(cudaEventRecord(lanes->in_ready[slot], lanes->h2d));
    // compute lane (the caller's stream): needs the inputs, and the slot's output staging drained (chunk c-2)
    RSA_CUDA_CHECK(cudaStreamWaitEvent(cs, lanes->in_ready[slot], 0));
    if (c >= 2) RSA_CUDA_CHECK(cudaStreamWaitEvent(cs, lanes->out_done[slot], 0));
    const rsa_attn_desc cd = chunk_desc(d, n, p.out_token_major);
    if ((rc = rsa_rectified_attention(&cd, base + p.off_in[slot][0], base + p.off_in[slot][1],
                                      base + p.off_in[slot][2], base + p.off_out[slot], base + p.off_ws, p.ws_bytes,
                                      stream)) != RSA_OK)
      return rc;
    RSA_CUDA_CHECK(cudaEventRecord(lanes->computed[slot], cs));
    // D2H lane
    RSA_CUDA_CHECK(cudaStreamWaitEvent(lanes->d2h, lanes->computed[slot], 0));
    if ((rc = copy_out(base + p.off_out[slot], (__nv_bfloat16*)out, d->o_stride, p.out_token_major, d->batch, h0, n,
                       d->seq, d->head_dim, lanes->d2h)) != RSA_OK)
      return rc;
    RSA_CUDA_CHECK(cudaEventRecord(lanes->out_done[slot], lanes->d2h));
  }
  // the call is complete, on the caller's stream, when the last copies out have landed
  RSA_CUDA_CHECK(cudaStreamWaitEvent(cs, lanes->out_done[(p.n_chunks - 1) & 1], 0));
  if (p.n_chunks >= 2) RSA_CUDA_CHECK(cudaStreamWaitEvent(cs, lanes->out_done[p.n_chunks & 1], 0));
  return RSA_OK;
}
