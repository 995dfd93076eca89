// Instantiations of the rbm2 kernels for walkers of 4 64-bit word(s).
#include "rbm2_impl.cuh"

namespace cgsvmc {
namespace rbm2 {

#define MC_CALL(NWV, LPWV, KJ4V) \
  launch_mc_variant<NWV, LPWV, KJ4V>(pl, img, packed, B, n_steps, seed, walker0, step0, accept_count, log_amp_out, st)
#define WALKER_CALL(NWV, LPWV, KJ4V) launch_walker_variant<NWV, LPWV, KJ4V>(pl, img, A, st)

int launch_mc_nw4(const Plan& pl, const float* img, uint64_t* packed, int64_t B, int n_steps,
                  uint64_t seed, uint64_t walker0, uint64_t step0, unsigned long long* accept_count,
                  float* log_amp_out, cudaStream_t st) {
  RBM2_VARIANT_SWITCH(4, MC_CALL)
}

int launch_walker_nw4(const Plan& pl, const float* img, const WalkerArgs& A, cudaStream_t st) {
  RBM2_VARIANT_SWITCH(4, WALKER_CALL)
}

}  // namespace rbm2
}  // namespace cgsvmc
