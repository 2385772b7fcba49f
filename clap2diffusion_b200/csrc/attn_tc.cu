// tcgen05 flash-attention forward (placeholder until the kernel lands: reports "unsupported" so that
// C2D_IMPL_AUTO resolves to the FFMA path and C2D_IMPL_TCGEN05 fails loudly).
#include "common.cuh"

namespace c2d {

bool attention_tc_supported(const AttnParams&, int) { return false; }

int attention_tc(const AttnParams&, int, cudaStream_t) {
  set_error("attention_tc: not built yet");
  return C2D_ERR_UNSUPPORTED;
}

}  // namespace c2d
