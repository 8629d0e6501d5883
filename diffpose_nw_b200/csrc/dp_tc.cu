// Tensor-core engine (tcgen05 / TMEM).  Placeholder until the kernel lands: reports "unsupported" so the
// fp32 engine serves every configuration.
#include "dp_internal.h"
namespace dp {
bool tc_supported(const Dims&) { return false; }
int tc_pack(dp_model*, cudaStream_t) { return DP_OK; }
void tc_free(dp_model*) {}
int tc_sample(dp_model*, const float*, int, float*, long, int, const dp_step*, const StepsArg*, int, const float*,
              const unsigned char*, cudaStream_t) {
  set_error("tensor-core engine not built");
  return DP_ERR_UNSUPPORTED;
}
}  // namespace dp
