// swb_kernels_m5.cu -- instantiates the wavefront engine kernels of mode 5 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode5(int R, int config) { return engine_kernel_lookup<5>(R, config); }
}
