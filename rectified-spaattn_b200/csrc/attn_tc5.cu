// attn_tc5.cu -- placeholder until the tcgen05 kernel lands (next commit).
#include "rsa_common.cuh"
namespace rsa {
int launch_attention_tc5(const AttnArgs& a, cudaStream_t s) {
  (void)a; (void)s;
  RSA_FAIL(RSA_ERR_UNSUPPORTED, "tcgen05 attention kernel not built");
}
}  // namespace rsa
