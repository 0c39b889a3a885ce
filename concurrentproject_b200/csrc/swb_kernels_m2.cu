// swb_kernels_m2.cu -- instantiates the wavefront engine kernels of mode 2 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode2(int R, int config) { return engine_kernel_lookup<2>(R, config); }
}
