// attn_tc.cu -- flash-style self-attention on tcgen05 (S = Q K^T and O += P V with TMEM accumulators).
#include "kernels.cuh"
#include "tc_common.cuh"

namespace xrd {

bool attention_tc_supported(const Tens& qkv, int heads) {
  (void)qkv; (void)heads;
  return false;   // until the tcgen05 kernel lands the CUDA-core flash kernel serves both modes
}

void attention_tc(Ctx& c, const Tens& qkv, int heads, Tens& out) {
  (void)c; (void)qkv; (void)heads; (void)out;
  fail(XRD_ERR_INVALID, "attention_tc: not available in this build");
}

}  // namespace xrd
