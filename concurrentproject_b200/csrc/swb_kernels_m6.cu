// swb_kernels_m6.cu -- instantiates the wavefront engine kernels of mode 6 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode6(int R, int config) { return engine_kernel_lookup<6>(R, config); }
}
