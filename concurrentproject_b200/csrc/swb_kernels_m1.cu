// swb_kernels_m1.cu -- instantiates the wavefront engine kernels of mode 1 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode1(int R, int config) { return engine_kernel_lookup<1>(R, config); }
}
