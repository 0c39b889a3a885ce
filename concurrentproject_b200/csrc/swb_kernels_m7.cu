// swb_kernels_m7.cu -- instantiates the wavefront engine kernels of mode 7 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode7(int R, int config) { return engine_kernel_lookup<7>(R, config); }
}
