// xcheck_api.cu -- C entry point of librsa_xcheck.so: the mma.sync (HMMA) cross-check implementation of kernel 4.
// TEST INFRASTRUCTURE, built into its own library under tests/xcheck/: the product library (librsa_b200.so) has exactly
// one attention kernel, the tcgen05 one, and no run-time switch.  The tests run this kernel over the kept-block lists,
// R and C that the PRODUCT's stages 2-3c left in a plan's workspace (read through the public rsa_attn_workspace_view),
// at sizes where the CPU oracle is too slow, and compare the two outputs.
#include <stdarg.h>
#include <stdio.h>

#include "rsa_common.cuh"

namespace rsa {
static thread_local char g_xerr[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_xerr, sizeof(g_xerr), fmt, ap);
  va_end(ap);
}
}  // namespace rsa

extern "C" {

typedef struct rsa_xcheck_args {
  const void *q, *k, *v;
  void* o;
  int32_t batch, heads;
  int64_t qs[3], ks[3], vs[3], os[3];  // element strides (batch, head, token)
  int32_t seq_q, seq_kv, kv_len, q_valid, vis_len, nq_vis, gap, nqt, nb;
  const uint16_t* kept_idx;  // [batch*heads, nqt, nb]
  const int32_t* kept_cnt;   // [batch*heads, nqt]
  const float* R;            // [batch*heads, nqt] or null
  const float* C;            // [batch*heads, nqt, 128] or null
  float scale_log2;
  int32_t q_round;
} rsa_xcheck_args;

const char* rsa_xcheck_last_error(void) { return rsa::g_xerr; }
size_t rsa_xcheck_args_size(void) { return sizeof(rsa_xcheck_args); }

int rsa_xcheck_attention(const rsa_xcheck_args* x, void* stream) {
  using namespace rsa;
  if (!x || !x->q || !x->k || !x->v || !x->o || !x->kept_idx || !x->kept_cnt) RSA_FAIL(RSA_ERR_ARG, "rsa_xcheck_attention: null pointer");
  AttnArgs a{};
  a.q = (const __nv_bfloat16*)x->q;
  a.k = (const __nv_bfloat16*)x->k;
  a.v = (const __nv_bfloat16*)x->v;
  a.o = (__nv_bfloat16*)x->o;
  a.batch = x->batch;
  a.heads = x->heads;
  for (int i = 0; i < 3; ++i) a.qs[i] = x->qs[i], a.ks[i] = x->ks[i], a.vs[i] = x->vs[i], a.os[i] = x->os[i];
  a.seq_q = x->seq_q, a.seq_kv = x->seq_kv, a.kv_len = x->kv_len, a.q_valid = x->q_valid;
  a.vis_len = x->vis_len, a.nq_vis = x->nq_vis, a.gap = x->gap, a.nqt = x->nqt, a.nb = x->nb;
  a.kept_idx = x->kept_idx, a.kept_cnt = x->kept_cnt, a.R = x->R, a.C = x->C;
  a.scale_log2 = x->scale_log2;
  a.q_round = x->q_round;
  a.head_dim = 128;
  return launch_attention_mma(a, (cudaStream_t)stream);
}

}  // extern "C"
