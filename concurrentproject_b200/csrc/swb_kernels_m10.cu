// swb_kernels_m10.cu -- instantiates the traceback-direction kernels (modes 10 and 11, see swb_kernels.cuh).
#include "swb_kernels.cuh"
namespace swb {
const void* engine_kernel_mode10(int R, int config) { return engine_kernel_lookup<10>(R, config); }
const void* engine_kernel_mode11(int R, int config) { return engine_kernel_lookup<11>(R, config); }
}
