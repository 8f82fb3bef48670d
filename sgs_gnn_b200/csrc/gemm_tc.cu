// K4 tensor-core path (tcgen05 + TMA).  Placeholder until the kernel lands: reports
// SGS_E_UNSUPPORTED so callers fail loudly instead of silently changing precision.
#include "common.cuh"

namespace sgs {
int32_t gemm_tc(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t,
                int32_t, int32_t, cudaStream_t) {
  set_error("sgs_gemm: tensor-core path not built yet");
  return SGS_E_UNSUPPORTED;
}
}  // namespace sgs
