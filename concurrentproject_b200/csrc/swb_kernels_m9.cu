// swb_kernels_m9.cu -- instantiates the wavefront engine kernels of mode 9 (see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode9(int R, int config) { return engine_kernel_lookup<9>(R, config); }
}
