// Generic tile networks (fc, rbm with hidden layers, conv): placeholder until
// the tile kernels land; every entry point fails loudly.
#include "internal.h"

namespace cgsvmc {
static int unsupported(const char* what) {
  set_error(std::string(what) + ": this ansatz is not built yet in the CUDA library");
  return CGSVMC_ERR_UNSUPPORTED;
}
int net_log_amp(const cgsvmc_ansatz*, const uint64_t*, int64_t, float*, cudaStream_t) { return unsupported("log_amp"); }
int net_mc_steps(const cgsvmc_ansatz*, uint64_t*, int64_t, int, uint64_t, uint64_t, uint64_t, unsigned long long*, float*, cudaStream_t) { return unsupported("mc_steps"); }
int net_mc_replay(const cgsvmc_ansatz*, uint64_t*, int64_t, const float*, const float*, int32_t*, int32_t*, float*, uint8_t*, cudaStream_t) { return unsupported("mc_step_replay"); }
int net_local_energy(const cgsvmc_ansatz*, const cgsvmc_ham*, const uint64_t*, int64_t, float*, float*, float*, float*, cudaStream_t) { return unsupported("local_energy"); }
int net_grad(cgsvmc_ansatz*, const uint64_t*, const float*, int64_t, int, float*, cudaStream_t) { return unsupported("weighted_grad_sum"); }
}  // namespace cgsvmc
