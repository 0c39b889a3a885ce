// swb_kernels_m3.cu -- instantiates the wavefront engine kernels of mode 3 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode3(int R, int config) { return engine_kernel_lookup<3>(R, config); }
}
