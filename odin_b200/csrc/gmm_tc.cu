// 3xTF32 tcgen05 path -- placeholder until the tensor-core kernels land.
#include "gmm.cuh"
namespace odin {
bool gmm_tc_supported(const odin_gmm*) { return false; }
int gmm_tc_refresh(odin_gmm*, cudaStream_t) { return ODIN_OK; }
int gmm_lse_tc(odin_gmm*, const float*, const uint8_t*, int64_t, float*, double*, cudaStream_t) {
  return set_error(ODIN_EINVAL, "tcgen05 path not built");
}
int gmm_stats_tc(odin_gmm*, const float*, const uint8_t*, int64_t, const float*, int, double*, cudaStream_t) {
  return set_error(ODIN_EINVAL, "tcgen05 path not built");
}
}  // namespace odin
