// swb_chain.cu -- instantiates the CTA-chained pair-engine kernels (launch config 7, see swb_chain.cuh).
#include "swb_chain.cuh"
namespace swb {
const void* chain_kernel(int mode, int R) {
#define SWB_CHAIN_CASE(RR) case RR: return mode == 0 ? (const void*)sw_chain_kernel<RR, 0> : (const void*)sw_chain_kernel<RR, 1>;
  if (mode != 0 && mode != 1) return nullptr;
  switch (R) {
    SWB_CHAIN_CASE(1) SWB_CHAIN_CASE(2) SWB_CHAIN_CASE(3) SWB_CHAIN_CASE(4) SWB_CHAIN_CASE(6) SWB_CHAIN_CASE(8)
    default: return nullptr;
  }
#undef SWB_CHAIN_CASE
}
}  // namespace swb
