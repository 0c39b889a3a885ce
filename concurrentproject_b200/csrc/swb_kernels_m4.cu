// swb_kernels_m4.cu -- instantiates the wavefront engine kernels of mode 4 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode4(int R, int config) { return engine_kernel_lookup<4>(R, config); }
}
