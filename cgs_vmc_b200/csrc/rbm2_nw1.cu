// Instantiations of the rbm2 kernels for walkers of 1 64-bit word(s).
#include "rbm2_impl.cuh"

namespace cgsvmc {
namespace rbm2 {

#define MC_CALL(NWV, LPWV, KJ4V) \
  launch_mc_variant<NWV, LPWV, KJ4V>(pl, img, packed, B, n_steps, seed, walker0, step0, accept_count, log_amp_out, st)
#define WALKER_CALL(NWV, LPWV, KJ4V) launch_walker_variant<NWV, LPWV, KJ4V>(pl, img, A, st)

int launch_mc_nw1(const Plan& pl, const float* img, uint64_t* packed, int64_t B, int n_steps,
                  uint64_t seed, uint64_t walker0, uint64_t step0, unsigned long long* accept_count,
                  float* log_amp_out, cudaStream_t st) {
  RBM2_VARIANT_SWITCH(1, MC_CALL)
}

int launch_walker_nw1(const Plan& pl, const float* img, const WalkerArgs& A, cudaStream_t st) {
  RBM2_VARIANT_SWITCH(1, WALKER_CALL)
}

}  // namespace rbm2
}  // namespace cgsvmc

#ifdef CGSVMC_RBM2_TIMING
// Development build only: copies the phase marks of the NW = 1 kernels
// ([kernel 0 = mc, 1 = walker][cta][mark][globaltimer ns, clock64]) to the host.
extern "C" int cgsvmc_debug_rbm2_marks(unsigned long long* host_out) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess)
    e = cudaMemcpyFromSymbol(host_out, cgsvmc::rbm2::g_phase_marks, sizeof(cgsvmc::rbm2::g_phase_marks));
  return e == cudaSuccess ? 0 : -2;
}
#endif

