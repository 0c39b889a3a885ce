// lat.cu -- dependent-chain latencies of the instructions the wavefront step is made of (one warp, one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/lat bench/lat.cu && ./bench/lat
// Prints cycles per dependent op; used to model the step loop's recurrence ("what bounds a step").
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 4096

template <int KIND>
__global__ void chain(uint32_t seed, long long* out, uint32_t* sink) {
  __shared__ uint32_t sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 4 + 4) & 4095;   // byte offsets chain
  __syncthreads();
  uint32_t v = seed + threadIdx.x, a = seed * 3 + 1, b = seed ^ 0x5555;
  const int lane = threadIdx.x & 31, src = (lane + 31) & 31;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (KIND == 0) v = __shfl_sync(0xffffffffu, v, src);
    else if (KIND == 1) v = __viaddmax_s16x2(v, a, b);
    else if (KIND == 2) { v = __vadd2(v, a); v = __viaddmax_s16x2(v, a, b); }            // cross-pipe pair
    else if (KIND == 3) v = *reinterpret_cast<volatile uint32_t*>(reinterpret_cast<char*>(sm) + (v & 4092));
    else if (KIND == 4) { uint32_t d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(a), "r"(b)); v = d; }
    else if (KIND == 5) { v = __shfl_sync(0xffffffffu, v, src); uint32_t d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(a), "r"(b)); v = d; }
    else if (KIND == 6) v = __vimax3_s16x2_relu(v, a, b);
    else if (KIND == 7) v = __vadd2(v, a);
    else if (KIND == 8) { v = __shfl_up_sync(0xffffffffu, v, 1); }
    else if (KIND == 9) { v = __vmaxs2(v, a); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = v;
}

template <int KIND>
void run(const char* name, int ops) {
  long long* d; uint32_t* s; long long h = 0;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4 * 128);
  for (int r = 0; r < 3; ++r) { chain<KIND><<<1, 32>>>(12345u, d, s); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("{\"chain\": \"%s\", \"cycles_per_iter\": %.2f, \"dependent_ops_per_iter\": %d}\n", name, (double)h / N, ops);
  cudaFree(d); cudaFree(s);
}

int main() {
  run<0>("SHFL.IDX", 1);
  run<8>("SHFL.UP", 1);
  run<1>("VIADDMNMX.S16x2", 1);
  run<6>("VIMNMX3.S16x2.RELU", 1);
  run<9>("VIMNMX.S16x2", 1);
  run<7>("VIADD.16x2", 1);
  run<2>("VIADD.16x2 -> VIADDMNMX.S16x2", 2);
  run<4>("PRMT", 1);
  run<3>("LDS (dependent address)", 1);
  run<5>("SHFL.IDX -> PRMT", 2);
  return 0;
}
