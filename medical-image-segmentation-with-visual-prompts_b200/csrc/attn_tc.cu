// bf16 tcgen05 / TMEM fused window attention (placeholder until the kernel lands: reports
// "unsupported" so that impl=auto uses the fp32-math kernels).
#include "attn.cuh"

namespace pwa {
bool attn_tc_supported(const AttnParams&, int) { return false; }
int attn_tc_forward(const AttnParams&, cudaStream_t) {
  set_error("tcgen05 attention kernel not built");
  return PWA_ERR_UNSUPPORTED;
}
}  // namespace pwa
