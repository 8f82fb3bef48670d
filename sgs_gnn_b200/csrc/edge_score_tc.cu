// K1 tensor-core forward (tcgen05 + TMEM).  Placeholder until the kernel lands.
#include "common.cuh"

namespace sgs {
size_t edge_score_tc_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }
int32_t edge_score_fwd_tc(const float*, int64_t, int64_t, const int32_t*, const int32_t*, const int32_t*, int64_t,
                          const float*, const float*, const float*, const float*, float, uint64_t, float*, void*,
                          size_t, int32_t, cudaStream_t) {
  set_error("sgs_edge_score_fwd: tensor-core path not built yet");
  return SGS_E_UNSUPPORTED;
}
}  // namespace sgs
