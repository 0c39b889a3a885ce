// swb_kernels_m8.cu -- instantiates the wavefront engine kernels of mode 8 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode8(int R, int config) { return engine_kernel_lookup<8>(R, config); }
}
