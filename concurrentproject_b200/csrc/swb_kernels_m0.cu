// swb_kernels_m0.cu -- instantiates the wavefront engine kernels of mode 0 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode0(int R, int config) { return engine_kernel_lookup<0>(R, config); }
}
